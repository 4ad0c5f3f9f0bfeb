"""`dropout_add_layer_norm` with the reference's signature (flash_attn/ops/layer_norm.py:207-252), running
bp_ln_residual_fwd, and differentiable through bp_ln_residual_bwd like the reference's DropoutAddLayerNormFn
(layer_norm.py:104-160).  dropout_p > 0 (training) is applied to x0 INSIDE the kernels, like the reference's
(ln_fwd_kernels.cuh:96-131): bp_ln_residual_fwd_dropout / bp_ln_residual_bwd_dropout regenerate a counter-based keep
mask from a 64-bit seed, so no mask tensor is stored; `layer_norm_dropout_mask` restates it (tests,
`return_dropout_mask=True`).  rowscale / layerscale are rejected."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..flash_attn_interface import _M32, _mix32, _new_seed, effective_dropout_p


def layer_norm_dropout_mask(seed: int, rows: int, cols: int, dropout_p: float, device="cpu") -> torch.Tensor:
    """The keep mask (rows, cols) bool that bp_ln_residual_fwd_dropout applies to x0 for `seed` -- a restatement of
    csrc/bp_common.cuh (drop_base with the "LN" tag / drop_row_word / drop_col_word / drop_keep) in int64 tensor
    arithmetic.  Plays the role of the dmask the reference's forward returns (layer_norm.py:120-131)."""
    thr = int(effective_dropout_p(dropout_p) * 256)
    base = _mix32((seed & _M32) ^ _mix32(((seed >> 32) & _M32) + 0x4C4E))     # drop_base(seed, "LN"): plain integers
    r = torch.arange(rows, dtype=torch.int64, device=device)
    c = torch.arange(cols, dtype=torch.int64, device=device)
    rw = _mix32(base + r * 0x9E3779B1)
    cw = _mix32((~base & _M32) + c * 0x85EBCA77)
    z = ((rw[:, None] ^ cw[None, :]) * 0x2C1B3C6D) & _M32
    return z >= (thr << 24)


def _ln_residual_forward(x0, x1, gamma, beta, epsilon, residual_in_fp32, want_residual, want_stats=False,
                         dropout_p=0.0, seed=0):
    _lib.require_cuda(x0, x1, gamma, beta)
    cols = x0.shape[-1]
    x0m = x0.reshape(-1, cols)
    if not x0m.is_contiguous():
        x0m = x0m.contiguous()
    x1m = None
    if x1 is not None:
        if x1.shape != x0.shape:
            raise RuntimeError("x1 must have the same shape as x0")
        x1m = x1.reshape(-1, cols)
        if not x1m.is_contiguous():
            x1m = x1m.contiguous()
    # residual dtype rule of the reference (ln_api.cpp:101-104): x1's dtype if given, else fp32 when
    # residual_in_fp32, else the input dtype.
    rdtype = x1.dtype if x1 is not None else (torch.float32 if residual_in_fp32 else x0.dtype)
    gamma, beta = gamma.contiguous(), beta.contiguous()
    if gamma.dtype != beta.dtype or gamma.shape != (cols,) or beta.shape != (cols,):
        raise RuntimeError("gamma and beta must both be (hidden,) with the same dtype")
    rows = x0m.shape[0]
    z = torch.empty_like(x0m)
    # the residual stream only has to be written when it differs from x0 or the caller wants it back
    need_x = want_residual and (x1 is not None or rdtype != x0.dtype or dropout_p > 0.0)
    x_out = torch.empty((rows, cols), dtype=rdtype, device=x0.device) if need_x else None
    mu = torch.empty(rows, dtype=torch.float32, device=x0.device) if want_stats else None
    rs = torch.empty(rows, dtype=torch.float32, device=x0.device) if want_stats else None
    args = (x0m.data_ptr(), _lib.ptr(x1m), gamma.data_ptr(), beta.data_ptr(), z.data_ptr(), _lib.ptr(x_out),
            _lib.ptr(mu), _lib.ptr(rs), rows, cols, float(epsilon), _lib.dtype_code(x0.dtype), _lib.dtype_code(rdtype),
            _lib.dtype_code(gamma.dtype))
    with torch.cuda.device(x0.device):
        if dropout_p > 0.0:
            _lib.check(_lib.load().bp_ln_residual_fwd_dropout(*args, float(dropout_p), int(seed), _lib.stream_ptr(x0.device)),
                       "bp_ln_residual_fwd_dropout")
        else:
            _lib.check(_lib.load().bp_ln_residual_fwd(*args, _lib.stream_ptr(x0.device)), "bp_ln_residual_fwd")
    z = z.reshape(x0.shape)
    res = None if not want_residual else (x_out.reshape(x0.shape) if need_x else x0)
    return (z, res, mu, rs) if want_stats else (z, res)


def layer_norm_from_residual(x, weight, bias, epsilon):
    """z = LayerNorm(x) for an already-summed residual stream x (fp32 in, weight dtype out): bp_ln_fwd.  Used by
    Block when the branch GEMM has added its output into the residual in its epilogue."""
    _lib.require_cuda(x, weight, bias)
    cols = x.shape[-1]
    if not x.is_contiguous():
        x = x.contiguous()
    weight, bias = weight.contiguous(), bias.contiguous()
    if weight.dtype != bias.dtype or weight.shape != (cols,) or bias.shape != (cols,):
        raise RuntimeError("gamma and beta must both be (hidden,) with the same dtype")
    if torch.is_grad_enabled() and any(t.requires_grad for t in (x, weight, bias)):
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    z = torch.empty(x.shape, dtype=weight.dtype, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.load().bp_ln_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), z.data_ptr(), None, None,
                                   x.numel() // cols, cols, float(epsilon), _lib.dtype_code(x.dtype),
                                   _lib.dtype_code(z.dtype), _lib.dtype_code(weight.dtype), _lib.stream_ptr(x.device))
    _lib.check(st, "bp_ln_fwd")
    return z


class _DropoutAddLayerNormFn(torch.autograd.Function):
    """The autograd node of DropoutAddLayerNormFn (layer_norm.py:104-160).  The forward keeps the pre-norm sum x (the
    residual stream it writes anyway), gamma, the row statistics mu / rsigma and -- with dropout -- the seed of the mask."""

    @staticmethod
    def forward(ctx, x0, x1, gamma, beta, epsilon, residual_in_fp32, prenorm, dropout_p=0.0, seed=0):
        z, x, mu, rs = _ln_residual_forward(x0, x1, gamma, beta, epsilon, residual_in_fp32, True, want_stats=True,
                                            dropout_p=dropout_p, seed=seed)
        ctx.save_for_backward(x, gamma, mu, rs)
        ctx.has_x1, ctx.prenorm, ctx.epsilon, ctx.x0_dtype = x1 is not None, prenorm, float(epsilon), x0.dtype
        ctx.dropout_p, ctx.seed = float(dropout_p), int(seed)
        ctx.set_materialize_grads(False)
        if not prenorm:
            return z
        # when no separate residual stream was written x is x0 itself; hand back a distinct tensor object
        return z, (x if x is not x0 else x0.view_as(x0))

    @staticmethod
    def backward(ctx, dz, dx=None):
        x, gamma, mu, rs = ctx.saved_tensors
        cols = x.shape[-1]
        rows = x.numel() // cols
        if dz is None:
            dz = torch.zeros(x.shape, dtype=ctx.x0_dtype, device=x.device)
        dz = dz.contiguous()
        if dx is not None:
            dx = dx.contiguous()
            if dx.dtype != x.dtype:
                dx = dx.to(x.dtype)
        dx0 = torch.empty(x.shape, dtype=ctx.x0_dtype, device=x.device)
        dx1 = torch.empty_like(x) if ctx.has_x1 else None
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        lib = _lib.load()
        ws_bytes = lib.bp_ln_bwd_workspace_bytes(cols)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        args = (dz.data_ptr(), _lib.ptr(dx), x.data_ptr(), gamma.data_ptr(), mu.data_ptr(), rs.data_ptr(),
                dx0.data_ptr(), _lib.ptr(dx1),
                dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), ws_bytes, rows, cols, ctx.epsilon,
                _lib.dtype_code(ctx.x0_dtype), _lib.dtype_code(x.dtype), _lib.dtype_code(gamma.dtype))
        with torch.cuda.device(x.device):
            if ctx.dropout_p > 0.0:
                _lib.check(lib.bp_ln_residual_bwd_dropout(*args, ctx.dropout_p, ctx.seed, _lib.stream_ptr(x.device)),
                           "bp_ln_residual_bwd_dropout")
            else:
                _lib.check(lib.bp_ln_residual_bwd(*args, _lib.stream_ptr(x.device)), "bp_ln_residual_bwd")
        return dx0, dx1, dgamma, dbeta, None, None, None, None, None


def dropout_add_layer_norm(x0, x1, weight, bias, dropout_p, epsilon, rowscale=None, layerscale=None,
                           prenorm=False, residual_in_fp32=False, return_dropout_mask=False, seed=None):
    """z = LayerNorm(dropout(x0) + x1) (and the fp32/16-bit residual dropout(x0) + x1 when prenorm=True).
    residual_in_fp32 only matters when x1 is None (layer_norm.py:209-212).  As in the reference the caller passes
    dropout_p = 0 in eval mode (DropoutAddLayerNorm.forward, layer_norm.py:248).  The effective probability is
    `effective_dropout_p(dropout_p)` (an 8-bit threshold, like the attention dropout); `seed` fixes the mask (default: 63
    bits from PyTorch's CPU generator); with `return_dropout_mask` the keep mask is appended to the outputs."""
    if rowscale is not None or layerscale is not None:
        raise RuntimeError("rowscale / layerscale are not supported (out of scope)")
    dropout_p = float(dropout_p)
    if not 0.0 <= dropout_p < 1.0:
        raise RuntimeError(f"dropout_p must be in [0, 1), got {dropout_p}")
    if dropout_p > 0.0 and seed is None:
        seed = _new_seed()
    seed = int(seed or 0)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x0, x1, weight, bias)):
        out = _DropoutAddLayerNormFn.apply(x0, x1, weight.contiguous(), bias.contiguous(), epsilon, residual_in_fp32,
                                           prenorm, dropout_p, seed)
        out = out if prenorm else (out,)
    else:
        z, res = _ln_residual_forward(x0, x1, weight, bias, epsilon, residual_in_fp32, prenorm, dropout_p=dropout_p,
                                      seed=seed)
        out = (z, res) if prenorm else (z,)
    if return_dropout_mask:
        cols = x0.shape[-1]
        mask = (layer_norm_dropout_mask(seed, x0.numel() // cols, cols, dropout_p, device=x0.device).view(x0.shape)
                if dropout_p > 0.0 else torch.ones(x0.shape, dtype=torch.bool, device=x0.device))
        out = (*out, mask)
    return out if len(out) > 1 else out[0]


class DropoutAddLayerNorm(nn.Module):
    """Module form (layer_norm.py:232-252); parameters `weight`, `bias`."""

    def __init__(self, hidden_size, prenorm=False, p=0.0, eps=1e-5, residual_in_fp32=False, device=None,
                 dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.prenorm = prenorm
        self.p = p
        self.epsilon = eps
        self.residual_in_fp32 = residual_in_fp32
        self.weight = nn.Parameter(torch.ones(hidden_size, **factory_kwargs))
        self.bias = nn.Parameter(torch.zeros(hidden_size, **factory_kwargs))

    def forward(self, x0, x1=None):
        return dropout_add_layer_norm(x0, x1, self.weight, self.bias, self.p if self.training else 0.0,
                                      self.epsilon, prenorm=self.prenorm, residual_in_fp32=self.residual_in_fp32)
