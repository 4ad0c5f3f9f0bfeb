"""Incremental decoding on the GPU.

Operator level: bp_decode_attn_fwd / bp_sense_mix_decode_fwd against the LAST ROW of the oracle's exact fp32 operators
(attention_fp32_ref non-causal on the cache = mha.py:437-440; sense_mix_fp32_ref row s-1), with the reference's 2x rule
against the same-precision eager composition.  Model level: a decode step must reproduce the logits of the full-prefix
forward at the last position (which is what the reference's generation loop computes, generation.py:34-44), and the
incremental greedy loop must emit the tokens of the re-run loop.
"""
import math

import pytest
import torch

from backpacks_flash_attn_b200 import _lib
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config, serving_config
from backpacks_flash_attn_b200.models.gpt import GPTLMHeadModel
from backpacks_flash_attn_b200.ops.decode import decode_attention, sense_mix_decode
from backpacks_flash_attn_b200.utils.generation import InferenceParams, greedy_decode
from backpacks_flash_attn_b200.utils.weights import name_seeded_
from oracle import backpack_oracle as oracle

pytestmark = pytest.mark.gpu


def _randn(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


# ------------------------------------------------------------------------------------------------------------
# decode attention
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("dh", [64, 128])
@pytest.mark.parametrize("batch,nheads,seqlen,max_seqlen", [(1, 12, 1, 16), (3, 12, 17, 64), (2, 4, 129, 129),
                                                            (5, 12, 1023, 1024), (64, 12, 1024, 1024), (1, 16, 4096, 4096)])
def test_decode_attention_matches_the_oracle(dtype, dh, batch, nheads, seqlen, max_seqlen):
    cache = _randn((batch, max_seqlen, 2, nheads, dh), dtype, 1)
    q = _randn((batch, 1, nheads, dh), dtype, 2)
    before = _lib.total_launches()
    out = decode_attention(q, cache, seqlen)
    assert _lib.total_launches() == before + 1
    k, v = cache[:, :seqlen, 0], cache[:, :seqlen, 1]
    ref, _ = oracle.attention_fp32_ref(q, k, v, causal=False)
    # same-precision eager composition of the reference (mha.py:226-268 CrossAttention, no mask)
    scale = 1.0 / math.sqrt(dh)
    p = torch.softmax(torch.einsum("bthd,bshd->bhts", q, k * scale), dim=-1, dtype=dtype)
    eager = torch.einsum("bhts,bshd->bthd", p, v)
    err, err_e = (out.float() - ref).abs().max().item(), (eager.float() - ref).abs().max().item()
    print(f"decode attn {dtype} dh{dh} b{batch} s{seqlen}: ours {err:.3e} eager {err_e:.3e}")
    assert err <= 2 * err_e + 1e-5          # tests/test_flash_attn.py:416 rule
    assert out.shape == q.shape and out.dtype == dtype


def test_decode_attention_per_sequence_lengths_and_strided_cache():
    dtype, b, h, dh, max_s = torch.bfloat16, 6, 12, 64, 300
    big = _randn((b + 2, max_s, 2, h, dh), dtype, 3)
    cache = big[1:1 + b]                              # a batch slice of a larger cache (batch_size_offset)
    q = _randn((b, h, dh), dtype, 4)
    lens = torch.tensor([1, 7, 64, 65, 299, 300], dtype=torch.int32, device="cuda")
    out = decode_attention(q, cache, 0, seqlens_k=lens)
    for i, n in enumerate(lens.tolist()):
        ref, _ = oracle.attention_fp32_ref(q[i:i + 1, None], cache[i:i + 1, :n, 0], cache[i:i + 1, :n, 1])
        assert (out[i].float() - ref[0, 0]).abs().max().item() < 2e-2
        # another batch size may split the keys over another number of CTAs: same result up to fp32 summation order
        one = decode_attention(q[i:i + 1], cache[i:i + 1], n)
        assert (one[0].float() - out[i].float()).abs().max().item() < 1e-2
    assert torch.equal(out, decode_attention(q, cache, 0, seqlens_k=lens))      # deterministic per launch configuration


def test_decode_attention_is_deterministic_and_rejects_bad_input():
    dtype = torch.float16
    cache = _randn((4, 512, 2, 8, 64), dtype, 5)
    q = _randn((4, 1, 8, 64), dtype, 6)
    a, b = decode_attention(q, cache, 511), decode_attention(q, cache, 511)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        decode_attention(q, cache, 513)
    with pytest.raises(RuntimeError):
        decode_attention(_randn((4, 2, 8, 64), dtype, 7), cache, 8)
    with pytest.raises(RuntimeError):
        decode_attention(q.float(), cache, 8)
    with pytest.raises(RuntimeError, match="head dim"):
        decode_attention(_randn((4, 1, 8, 32), dtype, 7), _randn((4, 16, 2, 8, 32), dtype, 8), 8)
    with pytest.raises(RuntimeError):
        decode_attention(q.cpu(), cache, 8)


# ------------------------------------------------------------------------------------------------------------
# sense-mix of the last position
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("batch,seqlen,nv,dk,d,vocab", [(1, 1, 16, 48, 768, 97), (2, 5, 16, 48, 768, 1000),
                                                        (3, 257, 16, 24, 384, 5000), (2, 1024, 16, 48, 768, 50264),
                                                        (64, 512, 16, 48, 768, 50264), (2, 300, 64, 16, 768, 3000),
                                                        (2, 129, 4, 192, 776, 500), (1, 4096, 16, 48, 768, 50264)])
def test_sense_mix_decode_matches_the_oracle(dtype, batch, seqlen, nv, dk, d, vocab):
    max_s = seqlen + 3
    table = _randn((vocab, nv, d), dtype, 11)
    ids = torch.randint(0, vocab, (batch, max_s), device="cuda", generator=torch.Generator("cuda").manual_seed(12))
    qk = _randn((batch, max_s, 2, nv, dk), dtype, 13, scale=1.5)
    k_cache = qk[:, :, 1].contiguous()
    q_last = qk[:, seqlen - 1, 0]
    before = _lib.total_launches()
    out = sense_mix_decode(q_last, k_cache, ids, table, seqlen)
    assert _lib.total_launches() == before + 1
    content = table[ids[:, :seqlen]].transpose(1, 2)                 # (b, nv, s, d)
    ref, _ = oracle.sense_mix_fp32_ref(qk[:, :seqlen], content)
    eager = oracle.sense_mix_eager(qk[:, :seqlen], content)
    err, err_e = (out.float() - ref[:, -1]).abs().max().item(), (eager[:, -1].float() - ref[:, -1]).abs().max().item()
    print(f"sense-mix decode {dtype} b{batch} s{seqlen} nv{nv} dk{dk} d{d}: ours {err:.3e} eager {err_e:.3e}")
    assert err <= 2 * err_e + 1e-5
    assert out.shape == (batch, d) and out.dtype == dtype
    assert torch.equal(out, sense_mix_decode(q_last, k_cache, ids, table, seqlen))     # deterministic


def test_sense_mix_decode_per_sequence_lengths_and_id_clamping():
    dtype, b, nv, dk, d, vocab, max_s = torch.bfloat16, 4, 16, 48, 768, 211, 400
    table = _randn((vocab, nv, d), dtype, 21)
    ids = torch.randint(0, vocab, (b, max_s), device="cuda", generator=torch.Generator("cuda").manual_seed(22))
    k_cache = _randn((b, max_s, nv, dk), dtype, 23)
    q = _randn((b, nv, dk), dtype, 24)
    lens = torch.tensor([1, 256, 257, 400], dtype=torch.int32, device="cuda")
    out = sense_mix_decode(q, k_cache, ids, table, 0, seqlens=lens)
    for i, n in enumerate(lens.tolist()):
        one = sense_mix_decode(q[i:i + 1], k_cache[i:i + 1], ids[i:i + 1], table, n)
        assert (one[0].float() - out[i].float()).abs().max().item() < 2e-2
    assert torch.equal(out, sense_mix_decode(q, k_cache, ids, table, 0, seqlens=lens))
    # out-of-range ids are clamped like in bp_sense_mix_table_fwd (never an out-of-bounds read)
    bad = ids.clone()
    bad[:, 3] = vocab + 1000
    bad[:, 5] = -7
    clamped = ids.clone()
    clamped[:, 3] = vocab - 1
    clamped[:, 5] = 0
    assert torch.equal(sense_mix_decode(q, k_cache, bad, table, 64), sense_mix_decode(q, k_cache, clamped, table, 64))


# ------------------------------------------------------------------------------------------------------------
# model level
# ------------------------------------------------------------------------------------------------------------
def _small(n_layer=4, **kw):
    cfg = serving_config(n_embd=768, n_head=12, n_layer=n_layer, n_positions=1024, **kw)
    return name_seeded_(BackpackLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)


def test_backpack_decode_step_matches_the_full_prefix_forward():
    model = _small()
    b, prompt, total = 3, 37, 48
    ids = torch.randint(0, 50257, (b, total), device="cuda", generator=torch.Generator("cuda").manual_seed(31))
    with torch.inference_mode():
        full = model(ids).logits                                      # (b, total, vocab): what re-running would give
        params = InferenceParams(max_sequence_len=total, max_batch_size=b)
        step = [model(ids[:, :prompt], inference_params=params, num_last_tokens=1).logits[:, -1]]
        params.sequence_len_offset = prompt
        before = _lib.total_launches()
        for t in range(prompt, total):
            pos = torch.full((b, 1), t, dtype=torch.long, device="cuda")
            step.append(model(ids[:, t:t + 1], position_ids=pos, inference_params=params).logits[:, -1])
            params.sequence_len_offset += 1
        assert _lib.total_launches() > before
    # bf16 runs of the same function through different kernels: compare on the scale of the logits
    for i, got in enumerate(step):
        want = full[:, prompt - 1 + i]
        err = (got.float() - want.float()).abs().max().item()
        ref_scale = want.float().abs().max().item()
        assert err <= 0.02 * ref_scale + 0.05, (i, err, ref_scale)
    agree = torch.stack([s.argmax(-1) for s in step], 1) == full[:, prompt - 1:].argmax(-1)
    assert agree.float().mean().item() >= 0.9


def test_backpack_decode_against_the_oracle_fp32_model():
    """The decode path against the oracle's fp32 model run on the whole prefix (a tiny config the CPU finishes in
    seconds), under the model-level rule next to the eager bf16 model (tests/models/test_gpt.py:60,70)."""
    from backpacks_flash_attn_b200.models.backpack import BackpackConfig
    dims = dict(n_embd=128, n_head=2, n_layer=2, n_positions=64, vocab_size=512, num_content_vectors=4)
    model = name_seeded_(BackpackLMHeadModel(serving_config(**dims)).eval()).to("cuda", torch.bfloat16)
    eager = name_seeded_(BackpackLMHeadModel(BackpackConfig(
        activation_function="gelu_new", reorder_and_upcast_attn=False, scale_attn_by_inverse_layer_idx=True,
        pad_vocab_size_multiple=8, **dims)).eval()).to("cuda", torch.bfloat16)
    ocfg = oracle.OracleConfig(**dims)
    names = oracle.canonical_param_shapes(ocfg)
    w = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if k in names}
    b, prompt, total = 2, 9, 20
    ids = torch.randint(0, 512, (b, total), generator=torch.Generator().manual_seed(5))
    ref = oracle.backpack_logits(ids, w, ocfg, fused_ln=True)[:, prompt - 1:]
    with torch.inference_mode():
        ref_e = eager(ids.cuda()).logits[:, prompt - 1:].float().cpu()
        params = InferenceParams(max_sequence_len=total, max_batch_size=b)
        got = [model(ids[:, :prompt].cuda(), inference_params=params, num_last_tokens=1).logits[:, -1]]
        params.sequence_len_offset = prompt
        for t in range(prompt, total):
            pos = torch.full((b, 1), t, dtype=torch.long, device="cuda")
            got.append(model(ids[:, t:t + 1].cuda(), position_ids=pos, inference_params=params).logits[:, -1])
            params.sequence_len_offset += 1
    got = torch.stack(got, 1).float().cpu()
    err, err_e = (got - ref).abs(), (ref_e - ref).abs()
    print(f"decode vs oracle fp32: ours max {err.max():.3e} mean {err.mean():.3e} | eager bf16 max {err_e.max():.3e} "
          f"mean {err_e.mean():.3e}")
    assert err.max() <= 3 * err_e.max() + 1e-3 and err.mean() <= 3 * err_e.mean() + 1e-4


def test_greedy_decode_incremental_equals_rerun():
    model = _small(n_layer=2)
    ids = torch.randint(0, 50257, (2, 11), device="cuda", generator=torch.Generator("cuda").manual_seed(41))
    inc = greedy_decode(ids, model, 40, cuda_graph=False)
    rerun = greedy_decode(ids, model, 40, incremental=False)
    assert inc.sequences.shape == (2, 40) and torch.equal(inc.sequences[:, :11], ids)
    # random-init logits are nearly flat, so a bf16-level difference can flip an argmax and the sequences diverge from
    # there on: compare token by token until the first flip, and require the flipped pair to be a near-tie
    for i in range(2):
        neq = (inc.sequences[i] != rerun.sequences[i]).nonzero()
        if len(neq):
            t = neq[0].item() - 11
            s = rerun.scores[t][i].float()
            a, c = inc.sequences[i, 11 + t], rerun.sequences[i, 11 + t]
            assert (s[c] - s[a]).abs().item() < 0.05 * s.abs().max().item() + 0.05
    assert len(inc.scores) == 29 and inc.scores[0].shape == (2, model.lm_head.weight.shape[0])


def test_graphed_decode_equals_the_eager_incremental_loop():
    """The CUDA-graph step (device-side offsets, index_copy_ cache writes, lengths read by the kernels) must produce the
    logits of the host-driven incremental loop: same kernels on the same data, so bit-identical up to the split choice
    of the decode kernels (they see `seqlens` instead of a host length)."""
    model = _small(n_layer=3)
    assert model.graphed_decode_ok()
    ids = torch.randint(0, 50257, (4, 19), device="cuda", generator=torch.Generator("cuda").manual_seed(61))
    before = dict(_lib.launch_counts)
    graphed = greedy_decode(ids, model, 60)                      # default: graph
    n_dec = _lib.launch_counts.get("bp_sense_mix_decode_fwd", 0) - before.get("bp_sense_mix_decode_fwd", 0)
    assert n_dec == 2                                            # one eager warm-up step + one capture; 40 replays
    eager = greedy_decode(ids, model, 60, cuda_graph=False)
    assert graphed.sequences.shape == (4, 60) and len(graphed.scores) == 41
    worst = max((a.float() - b.float()).abs().max().item() for a, b in zip(graphed.scores, eager.scores))
    for i in range(4):
        neq = (graphed.sequences[i] != eager.sequences[i]).nonzero()
        t = (neq[0].item() - 19) if len(neq) else 41
        # identical up to the first token flip (a near-tie); logits agree to bf16 noise before it
        for k in range(min(t + 1, 41)):
            a, c = graphed.scores[k][i].float(), eager.scores[k][i].float()
            assert (a - c).abs().max().item() <= 0.02 * c.abs().max().item() + 0.02
    print(f"graphed vs eager decode: worst logit difference {worst:.3e}")
    # no scores requested: none kept
    assert model.generate(ids, 24, return_dict_in_generate=True).scores is None
    assert torch.equal(model.generate(ids, 24), graphed.sequences[:, :24])      # same kernels, same data


def test_gpt_graphed_decode():
    from transformers import GPT2Config
    cfg = GPT2Config(n_embd=256, n_head=4, n_layer=2, vocab_size=1000, n_positions=128, activation_function="gelu_new",
                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    cfg.use_flash_attn = True
    cfg.fused_bias_fc = True
    cfg.fused_dense_gelu_dense = True
    cfg.fused_dropout_add_ln = True
    model = name_seeded_(GPTLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)
    ids = torch.randint(0, 1000, (3, 9), device="cuda", generator=torch.Generator("cuda").manual_seed(71))
    graphed = greedy_decode(ids, model, 40)
    eager = greedy_decode(ids, model, 40, cuda_graph=False)
    with torch.inference_mode():
        full = model(graphed.sequences).logits                     # teacher-forced on the graphed output
    for k in range(31):
        want = full[:, 8 + k].float()
        assert (graphed.scores[k].float() - want).abs().max().item() <= 0.02 * want.abs().max().item() + 0.02
    assert graphed.sequences.shape == eager.sequences.shape == (3, 40)


def test_gpt_decode_with_rotary_matches_the_full_prefix_forward():
    from transformers import GPT2Config
    cfg = GPT2Config(n_embd=256, n_head=4, n_layer=2, vocab_size=1000, n_positions=0, rotary_emb_fraction=0.5,
                     activation_function="gelu_new", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    cfg.use_flash_attn = True
    cfg.fused_bias_fc = True
    cfg.fused_dense_gelu_dense = True
    cfg.fused_dropout_add_ln = True
    model = name_seeded_(GPTLMHeadModel(cfg).eval()).to("cuda", torch.float16)
    b, prompt, total = 2, 13, 24
    ids = torch.randint(0, 1000, (b, total), device="cuda", generator=torch.Generator("cuda").manual_seed(51))
    with torch.inference_mode():
        full = model(ids).logits
        params = InferenceParams(max_sequence_len=total, max_batch_size=b)
        got = [model(ids[:, :prompt], inference_params=params, num_last_tokens=1).logits[:, -1]]
        params.sequence_len_offset = prompt
        for t in range(prompt, total):
            got.append(model(ids[:, t:t + 1], inference_params=params).logits[:, -1])
            params.sequence_len_offset += 1
    got = torch.stack(got, 1).float()
    want = full[:, prompt - 1:].float()
    assert (got - want).abs().max().item() <= 0.02 * want.abs().max().item() + 0.02


def test_decode_requires_single_position_after_the_prompt():
    model = _small(n_layer=1)
    ids = torch.randint(0, 50257, (1, 8), device="cuda")
    with torch.inference_mode():
        params = InferenceParams(max_sequence_len=16, max_batch_size=1)
        model(ids, inference_params=params)
        params.sequence_len_offset = 8
        with pytest.raises(RuntimeError, match="one position per call"):
            model(ids[:, :2], inference_params=params)
        params.sequence_len_offset = 16
        with pytest.raises(RuntimeError, match="too small"):
            model(ids[:, :1], inference_params=params)
