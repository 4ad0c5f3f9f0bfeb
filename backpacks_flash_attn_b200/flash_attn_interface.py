"""FlashAttention operators with the reference's Python signatures (flash_attn/flash_attn_interface.py:242-380),
backed by bp_fmha_fwd / bp_fmha_bwd instead of flash_attn_cuda.fwd / .bwd.

Differentiable like the reference's autograd functions (flash_attn_interface.py:50-240): when an input requires
grad the call goes through `_AttnFn`, whose backward is bp_fmha_bwd (deterministic).  `dropout_p > 0` runs the dropout
variants of both kernels: the keep mask is a counter-based function of (seed, batch, head, query, key) -- see
`attention_dropout_mask` -- drawn from a seed taken from PyTorch's CPU generator, so `torch.manual_seed` makes a run
reproducible and the backward regenerates the mask of its forward.  p is quantised to round(256 p) / 256.
"""
from __future__ import annotations

import torch

from . import _lib


_M32 = 0xFFFFFFFF


def _mix32(x):
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def effective_dropout_p(dropout_p: float) -> float:
    """The probability the kernels actually use: an 8-bit threshold, round(256 p) / 256 clamped to [1, 255] / 256."""
    if not dropout_p > 0.0:
        return 0.0
    return min(255, max(1, int(dropout_p * 256.0 + 0.5))) / 256.0


def attention_dropout_mask(seed: int, batch: int, nheads: int, seqlen_q: int, seqlen_k: int, dropout_p: float,
                           device="cpu") -> torch.Tensor:
    """The keep mask (batch, nheads, seqlen_q, seqlen_k) bool the dropout kernels apply for `seed` -- a restatement of
    csrc/bp_common.cuh (mix32 / drop_base / drop_row_word / drop_col_word / drop_keep) in int64 tensor arithmetic.
    Plays the role of the S_dmask the reference's forward can return (flash_attn_interface.py:50-68) for tests."""
    thr = int(effective_dropout_p(dropout_p) * 256)
    bh = torch.arange(batch * nheads, dtype=torch.int64, device=device)
    base = _mix32((seed & _M32) ^ _mix32(((seed >> 32) & _M32) + bh))                       # (bh,)
    q = torch.arange(seqlen_q, dtype=torch.int64, device=device)
    k = torch.arange(seqlen_k, dtype=torch.int64, device=device)
    rw = _mix32(base[:, None] + q[None, :] * 0x9E3779B1)                                    # (bh, sq)
    cw = _mix32((~base & _M32)[:, None] + k[None, :] * 0x85EBCA77)                          # (bh, sk)
    z = ((rw[:, :, None] ^ cw[:, None, :]) * 0x2C1B3C6D) & _M32
    return (z >= (thr << 24)).view(batch, nheads, seqlen_q, seqlen_k)


def _new_seed() -> int:
    """63 random bits from PyTorch's CPU generator (no device synchronisation; follows torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _check_common(dropout_p, return_attn_probs, *tensors):
    if not 0.0 <= dropout_p < 1.0:
        raise RuntimeError("dropout_p must be in [0, 1)")
    if return_attn_probs:
        raise RuntimeError("return_attn_probs is not supported (the S matrix is never materialised)")


def _flash_attn_forward(q, k, v, out, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                        softmax_scale, causal, out_fp32=False, dropout_p=0.0, seed=0):
    """q, k, v: (total, nheads, headdim) views with unit last stride (fmha_api.cpp:72-80).
    Returns (out, softmax_lse (batch, nheads, max_seqlen_q rounded up to 16) fp32).
    `out_fp32` is the test mode of SURVEY.md §8c (T2): `out` is an fp32 tensor of q's shape and receives O before
    the final 16-bit rounding (a debug switch of the library, not part of the C ABI)."""
    _lib.require_cuda(q, k, v, out, cu_seqlens_q, cu_seqlens_k)
    if q.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("FlashAttention only support fp16 and bf16 data type")      # fmha_api.cpp:215-217
    if k.dtype != q.dtype or v.dtype != q.dtype or out.dtype != (torch.float32 if out_fp32 else q.dtype):
        raise RuntimeError("query, key, value and out must have the same dtype")       # fmha_api.cpp:218-221
    if cu_seqlens_q.dtype != torch.int32 or cu_seqlens_k.dtype != torch.int32:
        raise RuntimeError("cu_seqlens must have dtype int32")                         # fmha_api.cpp:222-223
    for t in (q, k, v, out):
        if t.dim() != 3 or t.stride(-1) != 1:
            raise RuntimeError("q, k, v, out must be (total, nheads, headdim) with contiguous last dimension")
    if not (cu_seqlens_q.is_contiguous() and cu_seqlens_k.is_contiguous()):
        raise RuntimeError("cu_seqlens must be contiguous")
    total_q, nheads, d = q.shape
    total_k = k.shape[0]
    if k.shape != (total_k, nheads, d) or v.shape != (total_k, nheads, d) or out.shape != q.shape:
        raise RuntimeError("q/k/v/out shape mismatch")
    batch = cu_seqlens_q.numel() - 1
    if batch <= 0 or cu_seqlens_k.numel() != batch + 1:
        raise RuntimeError("cu_seqlens_q and cu_seqlens_k must both have batch+1 entries")
    if d % 8 != 0 or d > 128:
        raise RuntimeError("head dimension must be a multiple of 8 and at most 128")   # fmha_api.cpp:245
    lse_stride = (max_seqlen_q + 15) // 16 * 16                                        # fmha_api.cpp:254,276
    lse = torch.empty((batch, nheads, lse_stride), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        if out_fp32:
            lib.bp_debug_set_fmha_out_f32(1)
        try:
            args = (q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                    cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr(),
                    batch, nheads, d, total_q, total_k, max_seqlen_q, max_seqlen_k,
                    q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1),
                    out.stride(0), out.stride(1), lse_stride, float(softmax_scale), int(bool(causal)),
                    _lib.dtype_code(q.dtype))
            if dropout_p > 0.0:
                ws_bytes = lib.bp_fmha_fwd_dropout_workspace_bytes(batch, nheads, max_seqlen_k)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
                st = lib.bp_fmha_fwd_dropout(*args, float(dropout_p), int(seed), ws.data_ptr(), ws_bytes,
                                             _lib.stream_ptr(q.device))
            else:
                st = lib.bp_fmha_fwd(*args, _lib.stream_ptr(q.device))
        finally:
            if out_fp32:
                lib.bp_debug_set_fmha_out_f32(0)
    _lib.check(st, "bp_fmha_fwd_dropout" if dropout_p > 0.0 else "bp_fmha_fwd")
    return out, lse


def _flash_attn_backward(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k,
                         max_seqlen_q, max_seqlen_k, softmax_scale, causal, dropout_p=0.0, seed=0):
    """Gradients of _flash_attn_forward into the caller-allocated dq, dk, dv (strided views of a packed gradient are
    written in place, as flash_attn_interface.py:77-83 does).  Mirrors flash_attn_interface.py:31-47; returns
    (dq, dk, dv)."""
    _lib.require_cuda(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k)
    dout = dout.contiguous()   # e.g. the stride-0 gradient of out.sum(); the reference does the same (line 41)
    if q.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("FlashAttention only support fp16 and bf16 data type")
    if any(t.dtype != q.dtype for t in (dout, k, v, out, dq, dk, dv)) or softmax_lse.dtype != torch.float32:
        raise RuntimeError("dout, q, k, v, out, dq, dk, dv must share a 16-bit dtype and softmax_lse must be fp32")
    for t in (dout, q, k, v, out, dq, dk, dv):
        if t.dim() != 3 or t.stride(-1) != 1:
            raise RuntimeError("tensors must be (total, nheads, headdim) with contiguous last dimension")
    total_q, nheads, d = q.shape
    total_k = k.shape[0]
    if (dout.shape != q.shape or out.shape != q.shape or dq.shape != q.shape or k.shape != (total_k, nheads, d)
            or v.shape != k.shape or dk.shape != k.shape or dv.shape != k.shape):
        raise RuntimeError("shape mismatch between q/k/v/out/dout and their gradients")
    batch = cu_seqlens_q.numel() - 1
    if softmax_lse.dim() != 3 or softmax_lse.shape[:2] != (batch, nheads) or not softmax_lse.is_contiguous():
        raise RuntimeError("softmax_lse must be the contiguous (batch, nheads, lse_stride) tensor of the forward")
    lib = _lib.load()
    drop = dropout_p > 0.0
    ws_bytes = (lib.bp_fmha_bwd_dropout_workspace_bytes(batch, nheads, max_seqlen_q, max_seqlen_k) if drop
                else lib.bp_fmha_bwd_workspace_bytes(batch, nheads, max_seqlen_q))
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    import ctypes
    strides = (ctypes.c_int64 * 16)(*[x for t in (dout, q, k, v, out, dq, dk, dv) for x in (t.stride(0), t.stride(1))])
    args = (dout.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), softmax_lse.data_ptr(),
            dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr(),
            batch, nheads, d, total_q, total_k, max_seqlen_q, max_seqlen_k, ctypes.addressof(strides),
            softmax_lse.shape[2], float(softmax_scale), int(bool(causal)), _lib.dtype_code(q.dtype))
    with torch.cuda.device(q.device):
        if drop:
            st = lib.bp_fmha_bwd_dropout(*args, float(dropout_p), int(seed), workspace.data_ptr(), ws_bytes,
                                         _lib.stream_ptr(q.device))
        else:
            st = lib.bp_fmha_bwd(*args, workspace.data_ptr(), ws_bytes, _lib.stream_ptr(q.device))
    _lib.check(st, "bp_fmha_bwd")
    return dq, dk, dv


class _AttnFn(torch.autograd.Function):
    """One autograd node for the three packings of the reference (FlashAttnQKVPackedFunc / FlashAttnKVPackedFunc /
    FlashAttnFunc, flash_attn_interface.py:50-240).  `packing` = "qkv" (a = qkv), "kv" (a = q, b = kv) or "none"
    (a, b, c = q, k, v); the gradient of a packed input is allocated once and filled through strided views."""

    @staticmethod
    def forward(ctx, packing, a, b, c, cu_q, cu_k, max_q, max_k, softmax_scale, causal, dropout_p=0.0, seed=0):
        if packing == "qkv":
            q, k, v = a[:, 0], a[:, 1], a[:, 2]
        elif packing == "kv":
            q, k, v = a, b[:, 0], b[:, 1]
        else:
            q, k, v = a, b, c
        out, lse = _flash_attn_forward(q, k, v, torch.empty_like(q), cu_q, cu_k, max_q, max_k, softmax_scale, causal,
                                       dropout_p=dropout_p, seed=seed)
        ctx.dropout_p, ctx.seed = dropout_p, seed
        saved = {"qkv": (a,), "kv": (a, b)}.get(packing, (a, b, c))
        ctx.save_for_backward(*saved, out, lse, cu_q, cu_k)
        ctx.packing, ctx.max_q, ctx.max_k, ctx.softmax_scale, ctx.causal = packing, max_q, max_k, softmax_scale, causal
        return out

    @staticmethod
    def backward(ctx, dout):
        *inputs, out, lse, cu_q, cu_k = ctx.saved_tensors
        if ctx.packing == "qkv":
            qkv, = inputs
            dqkv = torch.empty_like(qkv)
            q, k, v, grads = qkv[:, 0], qkv[:, 1], qkv[:, 2], (dqkv, None, None)
            dq, dk, dv = dqkv[:, 0], dqkv[:, 1], dqkv[:, 2]
        elif ctx.packing == "kv":
            q, kv = inputs
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            k, v, dk, dv, grads = kv[:, 0], kv[:, 1], dkv[:, 0], dkv[:, 1], (dq, dkv, None)
        else:
            q, k, v = inputs
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            grads = (dq, dk, dv)
        _flash_attn_backward(dout, q, k, v, out, lse, dq, dk, dv, cu_q, cu_k, ctx.max_q, ctx.max_k,
                             ctx.softmax_scale, ctx.causal, dropout_p=ctx.dropout_p, seed=ctx.seed)
        return (None, *grads, None, None, None, None, None, None, None, None)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def flash_attn_unpadded_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p, softmax_scale=None,
                                       causal=False, return_attn_probs=False):
    """qkv: (total, 3, nheads, headdim); cu_seqlens: (batch+1,) int32.  Returns (total, nheads, headdim).
    Mirrors flash_attn_interface.py:242-267."""
    _check_common(dropout_p, return_attn_probs, qkv)
    if softmax_scale is None:
        softmax_scale = qkv.shape[-1] ** (-0.5)
    seed = _new_seed() if dropout_p > 0.0 else 0
    if _needs_grad(qkv):
        return _AttnFn.apply("qkv", qkv, None, None, cu_seqlens, cu_seqlens, max_seqlen, max_seqlen, softmax_scale,
                             causal, dropout_p, seed)
    out = torch.empty_like(qkv[:, 0])
    _flash_attn_forward(qkv[:, 0], qkv[:, 1], qkv[:, 2], out, cu_seqlens, cu_seqlens, max_seqlen, max_seqlen,
                        softmax_scale, causal, dropout_p=dropout_p, seed=seed)
    return out


def flash_attn_unpadded_kvpacked_func(q, kv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                      dropout_p, softmax_scale=None, causal=False, return_attn_probs=False):
    """q: (total_q, nheads, headdim); kv: (total_k, 2, nheads, headdim).  flash_attn_interface.py:270-303."""
    _check_common(dropout_p, return_attn_probs, q, kv)
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    seed = _new_seed() if dropout_p > 0.0 else 0
    if _needs_grad(q, kv):
        return _AttnFn.apply("kv", q, kv, None, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale,
                             causal, dropout_p, seed)
    out = torch.empty_like(q)
    _flash_attn_forward(q, kv[:, 0], kv[:, 1], out, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                        softmax_scale, causal, dropout_p=dropout_p, seed=seed)
    return out


def flash_attn_unpadded_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p,
                             softmax_scale=None, causal=False, return_attn_probs=False):
    """q: (total_q, nheads, headdim); k, v: (total_k, nheads, headdim).  flash_attn_interface.py:306-340."""
    _check_common(dropout_p, return_attn_probs, q, k, v)
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    seed = _new_seed() if dropout_p > 0.0 else 0
    if _needs_grad(q, k, v):
        return _AttnFn.apply("none", q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale,
                             causal, dropout_p, seed)
    out = torch.empty_like(q)
    _flash_attn_forward(q, k, v, out, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                        softmax_scale, causal, dropout_p=dropout_p, seed=seed)
    return out


def flash_attn_unpadded_with_lse(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                 softmax_scale=None, causal=False, out_fp32=False):
    """Like flash_attn_unpadded_func but also returns softmax_lse, as _flash_attn_forward does in the
    reference (flash_attn_interface.py:13-28)."""
    _check_common(0.0, False, q, k, v)
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    out = torch.empty(q.shape, dtype=torch.float32 if out_fp32 else q.dtype, device=q.device)
    return _flash_attn_forward(q, k, v, out, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                               softmax_scale, causal, out_fp32=out_fp32)


def flash_attn_func(qkv, cu_seqlens, dropout_p, max_s, softmax_scale=None, causal=False,
                    return_attn_probs=False):
    """Back-compat alias with the old argument order (flash_attn_interface.py:374-380)."""
    return flash_attn_unpadded_qkvpacked_func(qkv, cu_seqlens, max_s, dropout_p, softmax_scale, causal,
                                              return_attn_probs)
