"""GPU parity of bp_fmha_fwd (through the reference-shaped Python operators) against the oracle.

Criterion = the reference's own (tests/test_flash_attn.py:426-428):
    max|ours - fp32| <= 2 * max|same-precision eager PyTorch - fp32|      (+ 1e-5 slack)
on identical 16-bit-rounded inputs, plus LSE within 1e-3 (the fp32 side output), determinism
(tests/test_flash_attn.py:727-793) and the strided packed-qkv layout (flash_attn_interface.py:59).
"""
import ctypes
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


def _per_batch_check(qkv, causal):
    """The 2x rule of _check with the oracle evaluated one batch element at a time (bounds the s x s fp32 scores
    of the long-sequence cells)."""
    F = _ops()
    b, s, _, h, d = qkv.shape
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    q, k, v = (qkv[:, :, i].reshape(b * s, h, d) for i in range(3))
    out, lse = F.flash_attn_unpadded_with_lse(q, k, v, cu, cu, s, s, causal=causal)
    out = out.reshape(b, s, h, d)
    for i in range(b):
        piece = qkv[i:i + 1]
        ref, lse_ref = O.attention_fp32_ref(*piece.unbind(2), None, causal)
        eager = O.self_attention_eager(piece, None, causal)
        err, err_eager = O.max_abs(out[i:i + 1], ref), O.max_abs(eager, ref)
        assert err <= 2 * err_eager + 1e-5, f"batch {i}: max err {err:.3e} vs eager {err_eager:.3e}"
        assert O.mean_abs(out[i:i + 1], ref) <= 2 * O.mean_abs(eager, ref) + 1e-6
        assert O.max_abs(lse[i:i + 1, :, :s], lse_ref) < 1e-3


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("s,b", [(4096, 8), (2048, 16), (512, 64)])
def test_fmha_config5_cells(s, b, d):
    """Every cell of BASELINE config 5 the sweep in profiles/ times (seq x head dim at 32768 tokens, h = 768 / d);
    s = 1024 is test_fmha_config2_shape_and_modules."""
    qkv = _make_qkv(b, s, 768 // d, d, torch.bfloat16, seed=s + d)
    _per_batch_check(qkv, True)


@pytest.mark.parametrize("dtype,bound", [(torch.float16, 1e-3), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("d,s", [(64, 1024), (128, 1024), (64, 333), (40, 257)])
def test_fmha_fp32_output_mode(d, s, dtype, bound):
    """T2 of SURVEY.md §8c: the kernel's fp32-output test mode (O stored before the final 16-bit rounding) against
    the fp32 oracle on identical 16-bit-rounded inputs.

    What remains once the output rounding is gone is the rounding of P to the 16-bit MMA operand type before P.V --
    which the reference's kernel does too (csrc/flash_attn/src/fmha_fprop_kernel_1xN.h:508-512).  With fp16
    operands (11-bit mantissa) that is < 1e-3 max-abs, the north-star bar, and it is asserted here; with bf16
    operands (8-bit mantissa) a single probability of ~0.7 already carries up to 2^-9 * 0.7 * |v| ~ 4e-3 for
    |v| ~ 3, so the honest bound is 6e-3 max-abs, with the MEAN error and the LSE below 1e-3."""
    F = _ops()
    h = 4 if d <= 64 else 2
    qkv = _make_qkv(2, s, h, d, dtype, seed=7 * s + d)
    cu = torch.arange(0, 3 * s, s, dtype=torch.int32, device="cuda")
    q, k, v = (qkv[:, :, i].reshape(2 * s, h, d) for i in range(3))
    for causal in (True, False):
        out32, lse = F.flash_attn_unpadded_with_lse(q, k, v, cu, cu, s, s, causal=causal, out_fp32=True)
        assert out32.dtype == torch.float32
        ref, lse_ref = O.attention_fp32_ref(*qkv.unbind(2), None, causal)
        err = O.max_abs(out32.view_as(ref), ref)
        print(f"fp32-output mode d{d} s{s} {dtype} causal={causal}: max|err| {err:.2e} mean {O.mean_abs(out32.view_as(ref), ref):.2e}")
        assert err < bound, err
        assert O.mean_abs(out32.view_as(ref), ref) < 2e-4
        assert O.max_abs(lse[:, :, :s], lse_ref) < 1e-3
        # the production output is exactly the rounding of what the test mode stores
        out16, _ = F.flash_attn_unpadded_with_lse(q, k, v, cu, cu, s, s, causal=causal)
        assert torch.equal(out16, out32.to(dtype))


def test_fmha_key_padding_mask_and_flashmha():
    """FlashAttention(key_padding_mask=...) and FlashMHA against the reference's semantics
    (flash_attn/flash_attention.py:52-71: unpad -> varlen kernel -> pad with zeros; :74-101)."""
    from backpacks_flash_attn_b200.flash_attention import FlashAttention, FlashMHA
    torch.manual_seed(5)
    b, s, h, d = 4, 300, 3, 64
    lens = torch.tensor([300, 1, 129, 257], device="cuda")
    mask = torch.arange(s, device="cuda")[None, :] < lens[:, None]
    qkv = torch.randn(b, s, 3, h, d, device="cuda").bfloat16()
    for causal in (False, True):
        out, w = FlashAttention()(qkv, key_padding_mask=mask, causal=causal)
        assert w is None and out.shape == (b, s, h, d)
        for i, n in enumerate(lens.tolist()):
            piece = qkv[i:i + 1, :n]
            ref, _ = O.attention_fp32_ref(*piece.unbind(2), None, causal)
            eager = O.self_attention_eager(piece, None, causal)
            assert O.max_abs(out[i:i + 1, :n], ref) <= 2 * O.max_abs(eager, ref) + 1e-5
            assert out[i, n:].abs().max().item() == 0 if n < s else True      # padded slots are zero, as pad_input leaves them
    # unpadded entry of the module (cu_seqlens given) == packed function
    F = _ops()
    cu = torch.tensor([0, 300, 301, 430, 687], dtype=torch.int32, device="cuda")
    packed = torch.cat([qkv[i, :n] for i, n in enumerate(lens.tolist())])
    out_u, _ = FlashAttention()(packed, cu_seqlens=cu, max_s=300, causal=True)
    assert torch.equal(out_u, F.flash_attn_unpadded_qkvpacked_func(packed, cu, 300, 0.0, causal=True))
    # FlashMHA = Wqkv -> attention -> out_proj with the reference's parameter names
    mha = FlashMHA(h * d, h, causal=True, device="cuda", dtype=torch.bfloat16).eval()
    assert sorted(mha.state_dict()) == ["Wqkv.bias", "Wqkv.weight", "out_proj.bias", "out_proj.weight"]
    x = torch.randn(b, s, h * d, device="cuda").bfloat16()
    with torch.no_grad():
        y, _ = mha(x, key_padding_mask=mask)
        qkv2 = mha.Wqkv(x).reshape(b, s, 3, h, d)
        for i, n in enumerate(lens.tolist()):
            ref, _ = O.attention_fp32_ref(*qkv2[i:i + 1, :n].unbind(2), None, True)
            want = torch.nn.functional.linear(ref.reshape(1, n, h * d), mha.out_proj.weight.float(), mha.out_proj.bias.float())
            assert O.max_abs(y[i:i + 1, :n], want) < 3e-2


def test_fmha_rejects_bad_arguments():
    F = _ops()
    qkv = torch.zeros(64, 3, 2, 64, device="cuda", dtype=torch.float32)
    cu = torch.tensor([0, 64], dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="fp16 and bf16"):
        F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 64, 0.0)
    with pytest.raises(RuntimeError, match="dropout"):
        F.flash_attn_unpadded_qkvpacked_func(qkv.bfloat16(), cu, 64, 1.0)
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


def test_fmha_many_streams_get_their_own_scheduler_counters():
    """70 streams > the 64 per-stream slots of the first scheduler: counters are per launch now, so nothing is shared."""
    F = _ops()
    qkv = _make_qkv(2, 512, 12, 64, torch.bfloat16, seed=12).reshape(-1, 3, 12, 64)
    cu = torch.arange(0, 3 * 512, 512, dtype=torch.int32, device="cuda")
    ref = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 512, 0.0, causal=True)
    streams = [torch.cuda.Stream() for _ in range(70)]
    torch.cuda.synchronize()
    outs = []
    for st in streams:
        with torch.cuda.stream(st):
            outs.append(F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 512, 0.0, causal=True))
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, ref)


def test_fmha_poisoned_scheduler_state_cannot_leak_into_a_launch():
    """An aborted launch used to leave non-zero ticket counters behind; every launch now zeroes its own counter."""
    from backpacks_flash_attn_b200 import _lib
    F = _ops()
    qkv = _make_qkv(4, 1024, 12, 64, torch.bfloat16, seed=13).reshape(-1, 3, 12, 64)
    cu = torch.arange(0, 5 * 1024, 1024, dtype=torch.int32, device="cuda")
    ref = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True)
    lib = _lib.load()
    assert lib.bp_debug_poison_fmha_sched(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    for _ in range(3):
        assert torch.equal(F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True), ref)


def test_fmha_graph_replays_on_two_streams_and_next_to_eager():
    """Two captured graphs replayed concurrently on different streams, next to eager launches on a third: each
    captured launch owns its counter (a memset node re-arms it on every replay)."""
    F = _ops()
    qkv = _make_qkv(8, 1024, 12, 64, torch.bfloat16, seed=14).reshape(-1, 3, 12, 64)
    cu = torch.arange(0, 9 * 1024, 1024, dtype=torch.int32, device="cuda")
    ref = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True)
    graphs, outs = [], []
    for _ in range(2):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            o = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True)
        graphs.append(g)
        outs.append(o)
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(5):
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            graphs[0].replay()
        with torch.cuda.stream(s2):
            graphs[1].replay()
        with torch.cuda.stream(s3):
            eager = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, 1024, 0.0, causal=True)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref) and torch.equal(eager, ref)
