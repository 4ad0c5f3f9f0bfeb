"""End-to-end GPU parity of the Backpack forward (all fused kernels on) against the golden vectors that were
generated from the real reference, using the reference's model-level rule (tests/models/test_gpt.py:60,70):
our error against fp32 must stay below 3x the error of the reference's own same-precision eager path."""
import numpy as np
import pytest
import torch

from backpacks_flash_attn_b200 import _lib
from backpacks_flash_attn_b200.models.backpack import BackpackConfig, BackpackLMHeadModel, flash_config
from backpacks_flash_attn_b200.utils.weights import name_seeded_

pytestmark = pytest.mark.gpu


def _pair(dims):
    fused = name_seeded_(BackpackLMHeadModel(flash_config(**dims)).eval()).to("cuda", torch.bfloat16)
    eager = name_seeded_(BackpackLMHeadModel(BackpackConfig(
        num_content_vectors=16, vocab_size=50257, activation_function="gelu_new", reorder_and_upcast_attn=False,
        scale_attn_by_inverse_layer_idx=True, pad_vocab_size_multiple=8, **dims)).eval()).to("cuda", torch.bfloat16)
    return fused, eager


def _rule(ours, eager, ref, what):
    e_ours, e_eager = (ours.float() - ref).abs(), (eager.float() - ref).abs()
    print(f"{what}: ours max {e_ours.max():.3e} mean {e_ours.mean():.3e} | eager-bf16 max {e_eager.max():.3e} "
          f"mean {e_eager.mean():.3e} | max|ref| {ref.abs().max():.2f}")
    assert e_ours.max() <= 3 * e_eager.max() + 1e-3
    assert e_ours.mean() <= 3 * e_eager.mean() + 1e-4


def test_backpack_micro_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/micro_model.npz")
    fused, eager = _pair(dict(n_embd=384, n_head=6, n_layer=6, n_positions=512))
    ids = torch.from_numpy(g["ids"]).cuda()
    before = _lib.total_launches()
    with torch.inference_mode():
        hid, hid_e = fused.transformer(ids), eager.transformer(ids)
        logits, logits_e = fused(ids).logits, eager(ids).logits
        ctx_h, ctx_e = fused.transformer.gpt2_model(ids), eager.transformer.gpt2_model(ids)
        content, content_e = fused.transformer.content_model(ids), eager.transformer.content_model(ids)
    assert _lib.total_launches() > before          # the CUDA path is the one that ran
    _rule(ctx_h, ctx_e, torch.from_numpy(g["ctx_h"]).cuda(), "trunk hidden")
    _rule(content[1, 15], content_e[1, 15], torch.from_numpy(g["content_1_15"]).cuda(), "content[1,15]")
    _rule(hid, hid_e, torch.from_numpy(g["hid"]).cuda(), "backpack hidden")
    _rule(logits[1, 127], logits_e[1, 127], torch.from_numpy(g["logits_last"]).cuda(), "logits[1,127]")
    _rule(logits[:, :, :64], logits_e[:, :, :64], torch.from_numpy(g["logits_head"]).cuda(), "logits[:,:,:64]")
    agree = (logits.argmax(-1).cpu() == torch.from_numpy(g["argmax"])).float().mean().item()
    agree_e = (logits_e.argmax(-1).cpu() == torch.from_numpy(g["argmax"])).float().mean().item()
    print(f"greedy argmax agreement with fp32 reference: ours {agree:.3f}, eager bf16 {agree_e:.3f}")
    assert agree >= agree_e - 0.05
    assert list(content.shape) == [2, 16, 128, 384] and list(content.stride()) == g["content_strides"].tolist()


def test_backpack_small_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/small_model.npz")
    fused, eager = _pair(dict(n_embd=768, n_head=12, n_layer=12, n_positions=1024))
    ids = torch.from_numpy(g["ids"]).cuda()
    with torch.inference_mode():
        hid, hid_e = fused.transformer(ids), eager.transformer(ids)
        logits, logits_e = fused(ids).logits, eager(ids).logits
    _rule(hid, hid_e, torch.from_numpy(g["hid"]).cuda(), "small hidden")
    _rule(logits[0, 255, :256], logits_e[0, 255, :256], torch.from_numpy(g["logits_last_head"]).cuda(), "small logits")


def test_fused_and_eager_sense_mix_agree_inside_the_model():
    """transformer.sense_mix(h, C) (fused) vs torch.sum(contextualization_attn(h) @ C, 1) (reference composition,
    backpack.py:305-313), including an edited content tensor as the intervention wrappers build
    (intervened_models.py:78-101)."""
    fused, _ = _pair(dict(n_embd=384, n_head=6, n_layer=2, n_positions=512))
    ids = torch.randint(0, 50257, (3, 300), device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    with torch.inference_mode():
        t = fused.transformer
        h = t.gpt2_model(ids)
        content = t.content_model(ids)
        alpha = t.contextualization_attn(h)
        assert alpha.shape == (3, 16, 300, 300)
        want = torch.sum(alpha @ content, dim=1)
        got = t.sense_mix(h, content)
        assert torch.equal(got, t(ids))
        assert (got.float() - want.float()).abs().max() < 0.1
        edited = content.clone()
        edited[:, 3] *= 0.0            # knock one sense out
        got2 = t.sense_mix(h, edited)
        want2 = torch.sum(alpha @ edited, dim=1)
        assert (got2.float() - want2.float()).abs().max() < 0.1
        assert (got2.float() - got.float()).abs().max() > 1e-3


def _oracle_weights(model, dims):
    from oracle import backpack_oracle as O
    ocfg = O.OracleConfig(**dims)
    names = O.canonical_param_shapes(ocfg)
    # the model's own (bf16-rounded) weights, fp32 math: the oracle on identical inputs
    return ocfg, {k: v.detach().float() for k, v in model.state_dict().items() if k in names}


def test_sense_table_matches_the_oracle_content_vectors():
    """SURVEY.md §8f rank 1 against the ORACLE (not against this library): every row of the (vocab, nv, d) table must
    be oracle.content_vectors of that token (restating BackpackContentModule.forward, backpack.py:251-276), under the
    model-level rule (error < 3x the eager bf16 path's error, tests/models/test_gpt.py:60,70)."""
    from oracle import backpack_oracle as O
    dims = dict(n_embd=384, n_head=6, n_layer=2, n_positions=512)
    fused, eager = _pair(dims)
    ocfg, w = _oracle_weights(fused, dims)
    ids = torch.randint(0, 50257, (3, 200), device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    ids[0, :3] = torch.tensor([0, 50256, 50263], device="cuda")    # first, last real, last padded vocabulary row
    with torch.inference_mode():
        table = fused.transformer.build_sense_table(chunk=4096)
        assert table.shape == (50264, 16, 384)
        ref = O.content_vectors(ids, w, ocfg, fused_ln=True)                 # (b, nv, s, d) fp32
        content_e = eager.transformer.content_model(ids)
        got = table[ids].transpose(1, 2)
        _rule(got, content_e, ref, "sense table rows vs oracle.content_vectors")
        got_content = fused.transformer.content(ids)                          # the analysis-script accessor
        assert got_content.shape == ref.shape and torch.equal(got_content, got)
        assert list(got_content.stride()) == list(fused.transformer.content_model(ids).stride())


def test_sense_table_model_forward_matches_the_oracle():
    """The table-mode forward (`config.use_sense_table`: gather inside the sense-mix kernel, content model skipped)
    against the fp32 oracle of the whole model, next to the eager bf16 path; and against the non-table forward."""
    from oracle import backpack_oracle as O
    from backpacks_flash_attn_b200.models.backpack import serving_config
    dims = dict(n_embd=384, n_head=6, n_layer=2, n_positions=512)
    fused, eager = _pair(dims)
    served = name_seeded_(BackpackLMHeadModel(serving_config(**dims)).eval()).to("cuda", torch.bfloat16)
    assert served.transformer.use_sense_table and "transformer.sense_table" not in served.state_dict()
    ocfg, w = _oracle_weights(fused, dims)
    ids = torch.randint(0, 50257, (2, 300), device="cuda", generator=torch.Generator("cuda").manual_seed(8))
    with torch.inference_mode():
        before = dict(_lib.launch_counts)
        hid_t = served.transformer(ids)                # builds the table on first use
        assert _lib.launch_counts.get("bp_sense_mix_table_fwd", 0) == before.get("bp_sense_mix_table_fwd", 0) + 1
        hid, hid_e = fused.transformer(ids), eager.transformer(ids)
        ref = O.backpack_hidden(ids, w, ocfg, fused_ln=True)
        _rule(hid_t, hid_e, ref, "table-mode hidden vs oracle")
        _rule(served(ids).logits[:, -1], eager(ids).logits[:, -1], O.backpack_logits(ids, w, ocfg, fused_ln=True)[:, -1],
              "table-mode last-position logits vs oracle")
        # same kernels on the same rows: only a library GEMM picking another tile shape for another M can differ
        assert (hid_t.float() - hid.float()).abs().max() <= 0.06
        # training mode never uses the table
        served.train()
        assert served.transformer.current_sense_table() is None
        served.eval()
        assert served.transformer.current_sense_table() is not None


def test_sense_table_follows_weight_edits_and_device_moves():
    """the table must not serve stale vectors after load_state_dict / in-place weight edits (what the
    intervention scripts do to the content weights) and must follow .to()."""
    from backpacks_flash_attn_b200.models.backpack import serving_config
    dims = dict(n_embd=128, n_head=2, n_layer=1, n_positions=128)
    model = name_seeded_(BackpackLMHeadModel(serving_config(**dims)).eval()).to("cuda", torch.bfloat16)
    ids = torch.randint(0, 50257, (2, 128), device="cuda", generator=torch.Generator("cuda").manual_seed(9))
    with torch.inference_mode():
        a = model.transformer(ids).clone()
        t0 = model.transformer.sense_table
        assert model.transformer(ids) is not None and model.transformer.sense_table is t0     # cached
    with torch.no_grad():
        model.transformer.content_model.final_mlp.fc2.weight.mul_(0.5)                       # in-place edit
        model.transformer.content_model.final_mlp.fc2.bias.mul_(0.5)
    with torch.inference_mode():
        b = model.transformer(ids)
        assert model.transformer.sense_table is not t0                                       # rebuilt
        assert (b.float() - 0.5 * a.float()).abs().max() < 0.05 * a.float().abs().max()      # C halves => output halves
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["transformer.content_model.final_mlp.fc2.weight"] *= 2.0
    sd["transformer.content_model.final_mlp.fc2.bias"] *= 2.0
    model.load_state_dict(sd)
    with torch.inference_mode():
        c = model.transformer(ids)
        assert (c.float() - a.float()).abs().max() < 0.03 * a.float().abs().max() + 0.05
    model.transformer.drop_sense_table()
    assert model.transformer.sense_table is None


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_non_default_device():
    """Launches follow the tensors' device (the reference guards with CUDAGuard, fmha_api.cpp:267)."""
    fused, _ = _pair(dict(n_embd=128, n_head=2, n_layer=1, n_positions=128))
    ids = torch.randint(0, 50257, (2, 128), generator=torch.Generator().manual_seed(1))
    with torch.inference_mode():
        a = fused(ids.cuda(0)).logits
        b = fused.to("cuda:1")(ids.to("cuda:1")).logits
    assert torch.equal(a.cpu(), b.cpu())


def test_graphed_forward_replays_the_same_kernels():
    """utils/graph.GraphedForward: the captured CUDA graph reproduces the eager forward bit for bit, for new ids."""
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
    from backpacks_flash_attn_b200.utils.graph import GraphedForward
    from backpacks_flash_attn_b200.utils.weights import name_seeded_
    cfg = flash_config(n_embd=128, n_head=2, n_layer=2, n_positions=512)
    model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(7)
    ids_a = torch.randint(0, 50257, (2, 384), generator=g).cuda()
    ids_b = torch.randint(0, 50257, (2, 384), generator=g).cuda()
    with torch.inference_mode():
        fwd = GraphedForward(model, ids_a)
        for ids in (ids_a, ids_b, ids_a):
            out = fwd(ids).clone()
            assert torch.equal(out, model(ids).logits)
        with pytest.raises(RuntimeError, match="captured for ids of shape"):
            fwd(ids_a[:, :128])


def test_graphed_forward_with_sense_table():
    """The serving configuration (table gathered inside the kernel) under a CUDA graph: build the table first,
    capture, replay for new ids."""
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, serving_config
    from backpacks_flash_attn_b200.utils.graph import GraphedForward
    cfg = serving_config(n_embd=128, n_head=2, n_layer=2, n_positions=512)
    model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(11)
    ids_a = torch.randint(0, 50257, (2, 384), generator=g).cuda()
    ids_b = torch.randint(0, 50257, (2, 384), generator=g).cuda()
    with torch.inference_mode():
        model.transformer.build_sense_table()
        fwd = GraphedForward(model, ids_a)
        for ids in (ids_a, ids_b):
            assert torch.equal(fwd(ids).clone(), model(ids).logits)
