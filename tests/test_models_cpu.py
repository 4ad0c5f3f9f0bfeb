"""CPU checks of the host-side mirror of the reference interface: module tree, state-dict keys, config
flags, the un-fused path against the oracle (bit-exact in fp32), and that fused flags do NOT silently fall
back on a machine without a GPU."""
import numpy as np
import pytest
import torch

from backpacks_flash_attn_b200.models.backpack import (BackpackConfig, BackpackLMHeadModel, flash_config)
from backpacks_flash_attn_b200.models.gpt import GPTLMHeadModel, create_mixer_cls
from backpacks_flash_attn_b200.utils.weights import name_seeded_, parameter_checksum
from oracle import backpack_oracle as O


def micro_cfg(**kw):
    return BackpackConfig(num_content_vectors=16, n_embd=384, n_head=6, n_layer=6, n_positions=512, vocab_size=50257,
                          reorder_and_upcast_attn=False, scale_attn_by_inverse_layer_idx=True,
                          pad_vocab_size_multiple=8, **kw)


@pytest.fixture(scope="module")
def micro_model():
    return name_seeded_(BackpackLMHeadModel(micro_cfg()).eval())


def test_state_dict_keys_match_reference_layout(micro_model):
    keys = set(micro_model.state_dict().keys())
    canonical = set(O.canonical_param_shapes(O.OracleConfig(**O.MICRO)))
    assert canonical <= keys
    aliases = keys - canonical
    assert aliases == {"lm_head.weight",
                       "transformer.embeddings.word_embeddings.weight",
                       "transformer.embeddings.position_embeddings.weight",
                       "transformer.content_model.embeddings.word_embeddings.weight",
                       "transformer.content_model.embeddings.position_embeddings.weight"}
    for k, shape in O.canonical_param_shapes(O.OracleConfig(**O.MICRO)).items():
        assert tuple(micro_model.state_dict()[k].shape) == shape, k
    assert sum(p.numel() for p in micro_model.parameters()) == 41_659_776
    assert micro_model.config.vocab_size == 50264     # padded in place (backpack.py:285-288)
    assert micro_model.lm_head.weight is micro_model.transformer.embeddings.word_embeddings.weight


def test_name_seeded_weights_equal_oracle_recipe(micro_model):
    w = O.name_seeded_weights(O.OracleConfig(**O.MICRO))
    sd = micro_model.state_dict()
    for k, v in w.items():
        assert torch.equal(sd[k], v), k
    assert parameter_checksum(micro_model).item() > 0


def test_unfused_path_reproduces_reference_golden(micro_model, golden_dir):
    """Flags off == the reference's pure-PyTorch path: fp32 golden vectors generated from the real reference."""
    g = np.load(f"{golden_dir}/micro_model.npz")
    ids = torch.from_numpy(g["ids"])
    with torch.no_grad():
        t = micro_model.transformer
        ctx_h = t.gpt2_model(ids)
        alpha = t.contextualization_attn(ctx_h)
        content = t.content_model(ids)
        logits = micro_model(ids).logits
    np.testing.assert_allclose(ctx_h.numpy(), g["ctx_h"], atol=2e-6)
    np.testing.assert_allclose(alpha[0, 3].numpy(), g["alpha_0_3"], atol=2e-6)
    np.testing.assert_allclose(content[1, 15].numpy(), g["content_1_15"], atol=2e-6)
    assert list(content.stride()) == g["content_strides"].tolist()
    np.testing.assert_allclose(logits[1, 127].numpy(), g["logits_last"], atol=1e-5)
    assert torch.equal(logits.argmax(-1), torch.from_numpy(g["argmax"]))


def test_softmax_scale_per_layer():
    cfg = micro_cfg()
    for i, want in enumerate([0.125, 0.0625, 0.0416667]):
        mha = create_mixer_cls(cfg, layer_idx=i)(cfg.hidden_size)
        assert abs(mha.inner_attn.softmax_scale - want) < 1e-6


def test_flash_config_selects_fused_modules():
    from backpacks_flash_attn_b200.modules.mha import FlashSelfAttention
    from backpacks_flash_attn_b200.ops.fused_dense import FusedDense, FusedDenseGeluDense
    cfg = flash_config(n_embd=128, n_head=2, n_layer=1, n_positions=64)
    m = BackpackLMHeadModel(cfg)
    blk = m.transformer.gpt2_model.layers[0]
    assert isinstance(blk.mixer.inner_attn, FlashSelfAttention)
    assert isinstance(blk.mixer.Wqkv, FusedDense) and isinstance(blk.mlp, FusedDenseGeluDense)
    assert blk.fused_dropout_add_ln and m.transformer.fused_sense_mix
    assert isinstance(m.transformer.content_model.final_mlp, FusedDenseGeluDense)
    assert m.transformer.content_model.final_mlp.fc2.out_features == 16 * 128
    # same keys as the un-fused variant: checkpoints move freely between them (demo_generate.py:28-37)
    plain = BackpackLMHeadModel(BackpackConfig(num_content_vectors=16, n_embd=128, n_head=2, n_layer=1,
                                               n_positions=64, pad_vocab_size_multiple=8))
    assert set(plain.state_dict()) == set(m.state_dict())


def test_fused_model_has_no_cpu_fallback():
    cfg = flash_config(n_embd=128, n_head=2, n_layer=1, n_positions=64)
    m = BackpackLMHeadModel(cfg).eval()
    ids = torch.zeros(1, 16, dtype=torch.long)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(ids)


def test_gpt_lm_head_model_unfused_runs():
    from transformers import GPT2Config
    cfg = GPT2Config(n_embd=64, n_head=2, n_layer=2, n_positions=32, vocab_size=100)
    m = GPTLMHeadModel(cfg).eval()
    with torch.no_grad():
        out = m(torch.randint(0, 100, (2, 16))).logits
    assert out.shape == (2, 16, 100)


# ---- generation (training/src/utils/generation.py, flash_attn/utils/generation.py): host logic on the eager path ----
def test_eager_incremental_decode_equals_prefix_rerun():
    """The KV-cache / sense-cache bookkeeping (offsets, cache writes, position ids) on the un-fused model in fp32:
    every incremental step must give the logits the reference's re-run loop gives (generation.py:62-72)."""
    from backpacks_flash_attn_b200.utils.generation import InferenceParams
    torch.manual_seed(0)
    cfg = BackpackConfig(num_content_vectors=4, n_embd=64, n_head=4, n_layer=2, vocab_size=101, n_positions=64,
                         resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    m = BackpackLMHeadModel(cfg).eval()
    ids = torch.randint(0, 101, (3, 7))
    inc = m.generate(ids, 20, return_dict_in_generate=True, output_scores=True)
    rerun = m.generate(ids, 20, return_dict_in_generate=True, output_scores=True, incremental=False)
    assert inc.sequences.shape == (3, 20) and torch.equal(inc.sequences[:, :7], ids)
    assert torch.equal(inc.sequences, rerun.sequences)
    assert max((a - b).abs().max().item() for a, b in zip(inc.scores, rerun.scores)) < 1e-5
    # the re-run loop is the reference's: token t is the arg-max of the full forward on the prefix
    with torch.no_grad():
        assert torch.equal(rerun.sequences[:, 9], m(rerun.sequences[:, :9]).logits[:, -1].argmax(-1))
    assert m.generate(ids, 9).shape == (3, 9) and m.sample(ids, 9).shape == (3, 9)
    # caches: one per layer + the Backpack's two, sized by InferenceParams; a batch slice at batch_size_offset
    params = InferenceParams(max_sequence_len=12, max_batch_size=5, batch_size_offset=2)
    with torch.no_grad():
        first = m(ids, inference_params=params).logits
        assert torch.allclose(first, m(ids).logits, atol=1e-6)
    kv = params.key_value_memory_dict
    assert set(kv) == {0, 1, "backpack.ctx_k", "backpack.ids"}
    assert kv[0].shape == (5, 12, 2, 4, 16) and kv["backpack.ctx_k"].shape == (5, 12, 4, 16)
    assert torch.equal(kv["backpack.ids"][2:5, :7], ids)
    params.sequence_len_offset = 7
    with torch.no_grad(), pytest.raises(RuntimeError, match="one position per call"):
        m(ids[:, :2], inference_params=params)
    params.sequence_len_offset = 12
    with torch.no_grad(), pytest.raises(RuntimeError, match="too small"):
        m(ids[:, :1], inference_params=params)


def test_gpt_eager_incremental_decode_equals_prefix_rerun():
    from transformers import GPT2Config
    torch.manual_seed(1)
    g = GPTLMHeadModel(GPT2Config(n_embd=64, n_head=4, n_layer=2, vocab_size=101, n_positions=64, resid_pdrop=0.0,
                                  embd_pdrop=0.0, attn_pdrop=0.0)).eval()
    ids = torch.randint(0, 101, (2, 5))
    inc = g.generate(ids, 16, return_dict_in_generate=True, output_scores=True)
    rerun = g.generate(ids, 16, return_dict_in_generate=True, output_scores=True, incremental=False)
    assert torch.equal(inc.sequences, rerun.sequences)
    assert max((a - b).abs().max().item() for a, b in zip(inc.scores, rerun.scores)) < 1e-5
    with pytest.raises(ValueError):
        g.generate(ids, 5)
