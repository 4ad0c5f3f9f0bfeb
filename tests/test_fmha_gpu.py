"""GPU parity of bp_fmha_fwd (through the reference-shaped Python operators) against the oracle.

Criterion = the reference's own (tests/test_flash_attn.py:426-428):
    max|ours - fp32| <= 2 * max|same-precision eager PyTorch - fp32|      (+ 1e-5 slack)
on identical 16-bit-rounded inputs, plus LSE within 1e-3 (the fp32 side output), determinism
(tests/test_flash_attn.py:727-793) and the strided packed-qkv layout (flash_attn_interface.py:59).
"""
import math

import pytest
import torch

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from backpacks_flash_attn_b200 import flash_attn_interface as F
    return F


def _make_qkv(b, s, h, d, dtype, seed=0, style="randn"):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if style == "linear":   # the reference test's input style (tests/test_flash_attn.py:369-381)
        x = torch.randn(b, s, h * d, device="cuda", generator=g)
        w = torch.randn(3 * h * d, h * d, device="cuda", generator=g) * (h * d) ** -0.5
        qkv = (x @ w.t()).reshape(b, s, 3, h, d)
    else:
        qkv = torch.randn(b, s, 3, h, d, device="cuda", generator=g)
    return qkv.to(dtype)


def _check(qkv, causal, scale=None):
    F = _ops()
    b, s, _, h, d = qkv.shape
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    q, k, v = (qkv[:, :, i].reshape(b * s, h, d) for i in range(3))
    out, lse = F.flash_attn_unpadded_with_lse(q, k, v, cu, cu, s, s, softmax_scale=scale, causal=causal)
    out2 = F.flash_attn_unpadded_qkvpacked_func(qkv.reshape(b * s, 3, h, d), cu, s, 0.0, softmax_scale=scale,
                                                causal=causal)
    assert torch.equal(out, out2)
    ref, lse_ref = O.attention_fp32_ref(*qkv.unbind(2), scale, causal)
    eager = O.self_attention_eager(qkv, scale, causal)
    err = O.max_abs(out.reshape(b, s, h, d), ref)
    err_eager = O.max_abs(eager, ref)
    assert err <= 2 * err_eager + 1e-5, f"max err {err:.3e} vs eager {err_eager:.3e}"
    assert O.mean_abs(out.reshape(b, s, h, d), ref) <= 2 * O.mean_abs(eager, ref) + 1e-6
    assert O.max_abs(lse[:, :, :s], lse_ref) < 1e-3
    return err, err_eager


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("d", [64, 128, 32, 40, 80, 16])
@pytest.mark.parametrize("s", [97, 128, 200, 256, 257, 384, 512, 768, 1024, 1025, 2048])
def test_fmha_matches_oracle(s, d, causal, dtype):
    h = 4 if d <= 64 else 2
    qkv = _make_qkv(2, s, h, d, dtype, seed=s * 131 + d, style="linear" if s % 2 else "randn")
    _check(qkv, causal)


@pytest.mark.parametrize("layer_idx", [0, 5, 11])
def test_fmha_layer_scaled_softmax(layer_idx):
    """Backpack configs divide the scale by (layer_idx+1) (gpt.py:46-50)."""
    qkv = _make_qkv(2, 512, 12, 64, torch.bfloat16, seed=layer_idx)
    _check(qkv, True, scale=64 ** -0.5 / (layer_idx + 1))


def test_fmha_config2_shape_and_modules():
    """BASELINE config 2: b32 h12 s1024 d64 bf16 causal, through FlashAttention() and the packed func."""
    from backpacks_flash_attn_b200.flash_attention import FlashAttention
    qkv = _make_qkv(32, 1024, 12, 64, torch.bfloat16, seed=0)
    err, err_eager = _check(qkv, True)
    out, _ = FlashAttention()(qkv, causal=True)
    ref, _ = O.attention_fp32_ref(*qkv.unbind(2), None, True)
    assert O.max_abs(out, ref) <= 2 * err_eager + 1e-5
    hist = O.bf16_ulp_histogram(out, ref)
    total = sum(hist.values())
    # P is rounded to bf16 before the PV product (as in the reference, fmha_fprop_kernel_1xN.h:508-512), so
    # outputs that are small by cancellation sit several of their own (tiny) ulps away; the bulk is <= 1 ulp.
    print("bf16 ulp histogram vs rounded fp32 oracle:", hist)
    assert (hist[0] + hist[1]) / total > 0.8, hist


def test_fmha_varlen_matches_per_sequence():
    F = _ops()
    torch.manual_seed(3)
    lens = [5, 128, 300, 1, 257, 64]
    h, d = 3, 64
    total = sum(lens)
    qkv = torch.randn(total, 3, h, d, device="cuda").bfloat16()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, max(lens), 0.0, causal=True)
    start = 0
    for n in lens:
        piece = qkv[start:start + n].unsqueeze(0)
        ref, _ = O.attention_fp32_ref(*piece.unbind(2), None, True)
        eager = O.self_attention_eager(piece, None, True)
        assert O.max_abs(out[start:start + n], ref[0]) <= 2 * O.max_abs(eager, ref) + 1e-5
        start += n


def test_fmha_cross_lengths_noncausal():
    """seqlen_q != seqlen_k through the kv-packed entry point (flash_attn_interface.py:270-303)."""
    F = _ops()
    torch.manual_seed(4)
    b, sq, sk, h, d = 2, 130, 333, 2, 64
    q = torch.randn(b * sq, h, d, device="cuda").bfloat16()
    kv = torch.randn(b * sk, 2, h, d, device="cuda").bfloat16()
    cq = torch.arange(0, (b + 1) * sq, sq, dtype=torch.int32, device="cuda")
    ck = torch.arange(0, (b + 1) * sk, sk, dtype=torch.int32, device="cuda")
    out = F.flash_attn_unpadded_kvpacked_func(q, kv, cq, ck, sq, sk, 0.0, causal=False)
    qf, kf, vf = q.float().view(b, sq, h, d), kv[:, 0].float().view(b, sk, h, d), kv[:, 1].float().view(b, sk, h, d)
    p = torch.softmax(torch.einsum("bthd,bshd->bhts", qf, kf) / math.sqrt(d), -1)
    ref = torch.einsum("bhts,bshd->bthd", p, vf)
    assert O.max_abs(out.view(b, sq, h, d), ref) < 2e-2


def test_fmha_deterministic():
    F = _ops()
    qkv = _make_qkv(4, 1024, 12, 64, torch.bfloat16, seed=9).reshape(4 * 1024, 3, 12, 64)
    cu = torch.arange(0, 5 * 1024, 1024, dtype=torch.int32, device="cuda")
    first = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True)
    for _ in range(10):
        assert torch.equal(first, F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True))


def test_fmha_fp32_output_mode_bound():
    """T2 of SURVEY.md §8c: LSE (fp32 side output) within 1e-3 of the oracle; bf16 O within 1 ulp."""
    qkv = _make_qkv(2, 1024, 12, 64, torch.bfloat16, seed=21)
    F = _ops()
    cu = torch.arange(0, 3 * 1024, 1024, dtype=torch.int32, device="cuda")
    q, k, v = (qkv[:, :, i].reshape(2 * 1024, 12, 64) for i in range(3))
    out, lse = F.flash_attn_unpadded_with_lse(q, k, v, cu, cu, 1024, 1024, causal=True)
    ref, lse_ref = O.attention_fp32_ref(*qkv.unbind(2), None, True)
    assert O.max_abs(lse, lse_ref) < 1e-3
    assert O.mean_abs(out.view_as(ref), ref) < 5e-4


def test_fmha_rejects_bad_arguments():
    F = _ops()
    qkv = torch.zeros(64, 3, 2, 64, device="cuda", dtype=torch.float32)
    cu = torch.tensor([0, 64], dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="fp16 and bf16"):
        F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 64, 0.0)
    with pytest.raises(RuntimeError, match="dropout"):
        F.flash_attn_unpadded_qkvpacked_func(qkv.bfloat16(), cu, 64, 0.1)
    with pytest.raises(RuntimeError, match="int32"):
        F.flash_attn_unpadded_qkvpacked_func(qkv.bfloat16(), cu.long(), 64, 0.0)
    bad = torch.zeros(64, 3, 2, 36, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        F.flash_attn_unpadded_qkvpacked_func(bad, cu, 64, 0.0)


@pytest.mark.parametrize("b,h,s", [(149, 1, 256), (37, 8, 384), (3, 99, 130)])
def test_fmha_scheduler_chunk_boundaries(b, h, s):
    """More (batch, head) pairs than one scheduling chunk (148), a ragged last chunk, odd tile counts: every work
    item must be handed out exactly once by the ticket scheduler."""
    qkv = _make_qkv(b, s, h, 64, torch.bfloat16, seed=b * 7 + h)
    _check(qkv, True)


def test_fmha_concurrent_streams_do_not_share_scheduler_state():
    F = _ops()
    qkv = _make_qkv(8, 1024, 12, 64, torch.bfloat16, seed=11)
    cu = torch.arange(0, 9 * 1024, 1024, dtype=torch.int32, device="cuda")
    ref = F.flash_attn_unpadded_qkvpacked_func(qkv.reshape(-1, 3, 12, 64), cu, 1024, 0.0, causal=True)
    streams = [torch.cuda.Stream() for _ in range(3)]
    torch.cuda.synchronize()
    outs = []
    for rep in range(4):
        for st in streams:
            with torch.cuda.stream(st):
                outs.append(F.flash_attn_unpadded_qkvpacked_func(qkv.reshape(-1, 3, 12, 64), cu, 1024, 0.0, causal=True))
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, ref)
