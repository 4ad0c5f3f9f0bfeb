"""GPU parity of the sense-mix backward: the hand-derived backward of ops/sense_mix.py (batched
library GEMMs around ONE own element-wise pass, bp_sense_softmax_bwd) against fp32 autograd through the oracle's exact
operator (oracle.sense_mix_fp32_ref), next to autograd through the reference's own same-precision composition
(oracle.sense_mix_eager = backpack.py:116-122, 313).  Rule: the attention-gradient rule of the reference's tests
(tests/test_flash_attn.py: backward error <= a small multiple of the same-precision PyTorch error) with factor 3.

bp_sense_softmax_bwd itself is also checked directly against fp32 softmax / softmax-backward of its definition."""
import math

import pytest
import torch

from backpacks_flash_attn_b200 import _lib
from backpacks_flash_attn_b200.ops.sense_mix import (sense_mix, _sense_mix_backward, _sense_mix_backward_eager)
from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(b, s, nv, d, dtype, seed):
    g = torch.Generator("cuda").manual_seed(seed)
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda", generator=g).to(dtype)
    # the reference's layout: content model output (b, s, nv, d) handed over as a transposed view (backpack.py:276)
    content = (torch.randn(b, s, nv, d, device="cuda", generator=g) * 0.5).to(dtype)
    dout = torch.randn(b, s, d, device="cuda", generator=g).to(dtype)
    return qk, content, dout


def _grads(fn, qk, content, dout):
    q = qk.detach().clone().requires_grad_()
    c = content.detach().clone().requires_grad_()
    out = fn(q, c.transpose(1, 2))
    out.backward(dout.to(out.dtype))
    return out.detach(), q.grad, c.grad


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,s,nv,d", [(2, 256, 16, 768), (3, 200, 4, 768), (1, 1024, 16, 768), (2, 136, 8, 256),
                                      (1, 2048, 4, 256), (2, 64, 64, 768), (1, 4096, 4, 256), (1, 2304, 8, 256)])
def test_sense_mix_backward_matches_the_oracle(dtype, b, s, nv, d):
    qk, content, dout = _inputs(b, s, nv, d, dtype, seed=s + nv)
    before = _lib.launch_counts.get("bp_sense_softmax_bwd", 0)
    out, dq, dc = _grads(lambda q, c: sense_mix(q, c), qk, content, dout)
    assert _lib.launch_counts.get("bp_sense_softmax_bwd", 0) > before, "the fused backward pass did not run"
    _, dq_e, dc_e = _grads(lambda q, c: O.sense_mix_eager(q, c), qk, content, dout)
    _, dq_r, dc_r = _grads(lambda q, c: O.sense_mix_fp32_ref(q, c)[0], qk.float(), content.float(), dout.float())
    for name, ours, eager, ref in (("dqk", dq, dq_e, dq_r), ("dcontent", dc, dc_e, dc_r)):
        e_ours = (ours.float() - ref).abs()
        e_eager = (eager.float() - ref).abs()
        scale = ref.abs().max().item()
        print(f"{name}: ours max {e_ours.max():.3e} mean {e_ours.mean():.3e} | eager max {e_eager.max():.3e} "
              f"mean {e_eager.mean():.3e} | |ref| {scale:.3e}")
        assert torch.isfinite(ours.float()).all()
        assert e_ours.max() <= 3 * e_eager.max() + 1e-3 * scale, name
        assert e_ours.mean() <= 3 * e_eager.mean() + 1e-4 * scale, name
    # the gradient of the content tensor comes back in the content tensor's own (transposed-view) layout
    assert dc.shape == content.shape


def test_sense_mix_backward_is_deterministic_and_chunking_is_invisible():
    qk, content, dout = _inputs(5, 256, 16, 768, torch.bfloat16, seed=3)
    c = content.transpose(1, 2)
    scale = 48 ** -0.5
    a = _sense_mix_backward(qk, c, dout, scale, True, True)
    b_ = _sense_mix_backward(qk, c, dout, scale, True, True)
    assert torch.equal(a[0], b_[0]) and torch.equal(a[1], b_[1])
    # 2 batch elements per chunk (with a short last chunk): same values up to the library GEMM's batch-size heuristics
    small = _sense_mix_backward(qk, c, dout, scale, True, True, chunk_bytes=2 * 2 * 16 * 256 * 256 * 2)
    assert (small[0].float() - a[0].float()).abs().max() <= 2e-2 * a[0].float().abs().max()
    assert (small[1].float() - a[1].float()).abs().max() <= 2e-2 * a[1].float().abs().max()
    # only one of the two gradients requested
    only_q = _sense_mix_backward(qk, c, dout, scale, True, False)
    only_c = _sense_mix_backward(qk, c, dout, scale, False, True)
    assert only_q[1] is None and only_c[0] is None
    assert torch.equal(only_q[0], a[0]) and torch.equal(only_c[1], a[1])


def test_expanded_and_sliced_content_layouts():
    """A content tensor that is not a permutation of a dense one (broadcast over the batch, or a slice of a wider tensor)
    gets a plain contiguous gradient: no aliasing through stride-0 or padded strides."""
    g = torch.Generator("cuda").manual_seed(8)
    b, s, nv, d = 3, 128, 4, 256
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda", generator=g).bfloat16()
    dout = torch.randn(b, s, d, device="cuda", generator=g).bfloat16()
    base = (torch.randn(1, nv, s, d, device="cuda", generator=g) * 0.5).bfloat16()
    wide = (torch.randn(b, nv, s, 2 * d, device="cuda", generator=g) * 0.5).bfloat16()
    for src, make in ((base, lambda t: t.expand(b, nv, s, d)), (wide, lambda t: t[..., :d])):
        leaf = src.clone().requires_grad_()
        leaf_e = src.clone().requires_grad_()
        sense_mix(qk, make(leaf)).backward(dout)
        O.sense_mix_eager(qk, make(leaf_e)).backward(dout)
        assert leaf.grad.shape == src.shape
        err = (leaf.grad.float() - leaf_e.grad.float()).abs().max().item()
        assert err <= 3e-2 * leaf_e.grad.float().abs().max().item() + 1e-3, err


def test_unsupported_sequence_lengths_fall_back_to_autograd_through_the_eager_composition():
    qk, content, dout = _inputs(2, 203, 4, 256, torch.bfloat16, seed=4)      # 203 is not a multiple of 8
    before = _lib.launch_counts.get("bp_sense_softmax_bwd", 0)
    _, dq, dc = _grads(lambda q, c: sense_mix(q, c), qk, content, dout)
    assert _lib.launch_counts.get("bp_sense_softmax_bwd", 0) == before
    ref = _sense_mix_backward_eager(qk, content.transpose(1, 2), dout, 64 ** -0.5, True, True)
    assert torch.equal(dq, ref[0]) and torch.equal(dc, ref[1].transpose(1, 2))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mats,s", [(3, 8), (5, 200), (2, 256), (3, 520), (2, 1024), (1, 2048), (1, 2056), (1, 4096)])
def test_softmax_backward_kernel_directly(dtype, mats, s):
    g = torch.Generator("cuda").manual_seed(s)
    scores = (torch.randn(mats, s, s, device="cuda", generator=g) * 4).to(dtype)
    dalpha = torch.randn(mats, s, s, device="cuda", generator=g).to(dtype)
    scale = 0.17
    lib = _lib.load()
    S, dA = scores.clone(), dalpha.clone()
    # poison the strictly-upper triangle: it must never be read
    upper = torch.ones(s, s, dtype=torch.bool, device="cuda").triu(1)
    S.masked_fill_(upper, float("nan"))
    dA.masked_fill_(upper, float("nan"))
    _lib.check(lib.bp_sense_softmax_bwd(S.data_ptr(), dA.data_ptr(), mats * s, s, scale, _lib.dtype_code(dtype),
                                        _lib.stream_ptr(S.device)), "bp_sense_softmax_bwd")
    x = (scores.float() * scale).masked_fill(upper, float("-inf"))
    p = torch.softmax(x, -1)
    pr = p.to(dtype).float()                                   # the kernel's gradient uses the rounded probabilities
    da = dalpha.float().masked_fill(upper, 0.0)
    ds = scale * pr * (da - (pr * da).sum(-1, keepdim=True))
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert torch.isfinite(S.float()).all() and torch.isfinite(dA.float()).all()
    assert (S.float() - p).abs().max() <= ulp + 1e-6                                     # p <= 1: one rounding
    assert (dA.float() - ds).abs().max() <= 2 * ulp * ds.abs().max() + 1e-6
    assert (S.float()[:, upper] == 0).all() and (dA.float()[:, upper] == 0).all()


def test_softmax_backward_kernel_rejects_bad_arguments():
    lib = _lib.load()
    t = torch.zeros(16, 16, device="cuda", dtype=torch.bfloat16)
    st = _lib.stream_ptr(t.device)
    dt = _lib.dtype_code(torch.bfloat16)
    assert lib.bp_sense_softmax_bwd(t.data_ptr(), t.data_ptr(), 16, 12, 1.0, dt, st) == -2      # seqlen % 8
    assert lib.bp_sense_softmax_bwd(t.data_ptr(), t.data_ptr(), 8200, 8200, 1.0, dt, st) == -2  # seqlen > 8192
    assert lib.bp_sense_softmax_bwd(t.data_ptr(), t.data_ptr(), 17, 16, 1.0, dt, st) == -1      # rows % seqlen
    assert lib.bp_sense_softmax_bwd(0, t.data_ptr(), 16, 16, 1.0, dt, st) == -1
    assert lib.bp_sense_softmax_bwd(t.data_ptr(), t.data_ptr(), 16, 16, 1.0, 2, st) == -1       # fp32 storage
    assert math.isfinite(float(t.sum()))
