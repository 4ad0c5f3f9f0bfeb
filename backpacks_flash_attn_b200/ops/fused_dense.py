"""`FusedDense` / `FusedDenseGeluDense` with the reference's interface (flash_attn/ops/fused_dense.py:116-129,
357-402), differentiable like FusedDenseFunc / FusedDenseGeluDenseFunc (fused_dense.py:39-80, 179-300).

* `FusedDense.forward` is a GEMM + bias (F.linear in the reference, fused_dense.py:52,112).  On the inference
  path (CUDA, fp16/bf16, no autograd) `linear()` picks between this library's tcgen05 GEMM
  (bp_linear_bias_act_fwd with BP_ACT_NONE) and the cuBLAS call F.linear makes, per shape, by measurement -- see
  `_backend` below.  Under autograd, on the CPU or for shapes the kernel does not take (n or k not a multiple of
  8) it is F.linear, exactly as in the reference.
* `FusedDenseGeluDense`: fc1 + bias + tanh-GELU is ONE kernel, bp_linear_bias_act_fwd -- the replacement of
  fused_dense_lib.linear_gelu_forward (csrc/fused_dense_lib/fused_dense.cpp:88-142) -- followed by the fc2
  GEMM (fused_dense.py:225).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


def _dense_bias(bias: torch.Tensor | None, n: int) -> torch.Tensor | None:
    """The C side reads `bias` as n dense elements: reject a wrongly sized bias, densify a strided one (e.g. a slice
    of a fused parameter)."""
    if bias is None:
        return None
    if bias.shape != (n,):
        raise RuntimeError(f"bias must have shape ({n},), got {tuple(bias.shape)}")
    return bias if bias.is_contiguous() else bias.contiguous()


def linear_bias_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None,
                    activation: str = "gelu_tanh", save_pre_act: bool = True) -> torch.Tensor:
    """act(x @ weight.T + bias) through the hand-written tcgen05 GEMM (weight in nn.Linear layout).  Under autograd
    with an activation, `save_pre_act` makes the forward kernel also store the pre-activation for the backward (the
    reference's checkpoint_lvl 0 / 1); otherwise the backward recomputes it with one more GEMM (checkpoint_lvl 2)."""
    _lib.require_cuda(x, weight, bias)
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype:
        raise RuntimeError("linear_bias_act needs fp16/bf16 activations and weights of the same dtype")
    if bias is not None and bias.dtype != x.dtype:
        raise RuntimeError("bias must have the activation dtype")
    n, k = weight.shape
    if x.shape[-1] != k:
        raise RuntimeError("shape mismatch between x and weight")
    bias = _dense_bias(bias, n)
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad
                                    or (bias is not None and bias.requires_grad)):
        return _LinearBiasActFn.apply(x, weight, bias, activation, save_pre_act)
    x2 = x.reshape(-1, k)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    weight = weight.contiguous()
    out = torch.empty((x2.shape[0], n), dtype=x.dtype, device=x.device)
    act = {"none": _lib.BP_ACT_NONE, "gelu_tanh": _lib.BP_ACT_GELU_TANH}[activation]
    with torch.cuda.device(x.device):
        st = _lib.load().bp_linear_bias_act_fwd(x2.data_ptr(), weight.data_ptr(), _lib.ptr(bias), out.data_ptr(),
                                                x2.shape[0], n, k, act, _lib.dtype_code(x.dtype),
                                                _lib.stream_ptr(x.device))
    _lib.check(st, "bp_linear_bias_act_fwd")
    return out.reshape(*x.shape[:-1], n)


def _linear_bias_act_aux(x2: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, activation: str):
    """(act(x W^T + b), x W^T + b) from ONE pass of the GEMM (bp_linear_bias_act_aux_fwd: the reference's
    linear_gelu_forward with save_gelu_in, fused_dense.py:220-222).  x2 (m, k) contiguous, m >= 256."""
    n, k = weight.shape
    out = torch.empty((x2.shape[0], n), dtype=x2.dtype, device=x2.device)
    pre = torch.empty_like(out)
    act = {"none": _lib.BP_ACT_NONE, "gelu_tanh": _lib.BP_ACT_GELU_TANH}[activation]
    with torch.cuda.device(x2.device):
        st = _lib.load().bp_linear_bias_act_aux_fwd(x2.data_ptr(), weight.data_ptr(), _lib.ptr(bias), out.data_ptr(),
                                                    pre.data_ptr(), x2.shape[0], n, k, act, _lib.dtype_code(x2.dtype),
                                                    _lib.stream_ptr(x2.device))
    _lib.check(st, "bp_linear_bias_act_aux_fwd")
    return out, pre


def bias_act_backward(dact: torch.Tensor, pre: torch.Tensor | None, activation: str, want_dbias: bool):
    """The epilogue half of the dense backward (bp_bias_act_bwd): returns (dpre, dbias) for "gelu_tanh"
    (dpre = dact * gelu'(pre)) or (dact, dbias) for "none"; dbias = column sums, None unless wanted."""
    _lib.require_cuda(dact, pre)
    n = dact.shape[-1]
    d2 = dact.reshape(-1, n)
    if not d2.is_contiguous():
        d2 = d2.contiguous()
    gelu = activation == "gelu_tanh"
    if not gelu and not want_dbias:
        return dact, None
    p2 = dpre = None
    if gelu:
        p2 = pre.reshape(-1, n)
        if not p2.is_contiguous():
            p2 = p2.contiguous()
        dpre = torch.empty_like(d2)
    dbias = torch.empty(n, dtype=dact.dtype, device=dact.device) if want_dbias else None
    lib = _lib.load()
    ws_bytes = lib.bp_bias_act_bwd_workspace_bytes(n) if want_dbias else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dact.device) if want_dbias else None
    with torch.cuda.device(dact.device):
        st = lib.bp_bias_act_bwd(d2.data_ptr(), _lib.ptr(p2), _lib.ptr(dpre), _lib.ptr(dbias), _lib.ptr(ws), ws_bytes,
                                 d2.shape[0], n, _lib.BP_ACT_GELU_TANH if gelu else _lib.BP_ACT_NONE,
                                 _lib.dtype_code(dact.dtype), _lib.stream_ptr(dact.device))
    _lib.check(st, "bp_bias_act_bwd")
    return (dpre.reshape(dact.shape) if gelu else dact), dbias


def _dgrad(dy2: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """dX = dY W for W in nn.Linear layout (n, k): this library's GEMM on a transposed copy of the (small) weight when
    the shape is one it takes, cuBLAS otherwise (the reference's backward GEMMs are all cuBLASLt)."""
    n, k = weight.shape
    if _choice(k, n, dy2.shape[0]) == "own" and n % 8 == 0 and k % 8 == 0 and dy2.shape[0] >= 512:
        with torch.no_grad():
            return linear_bias_act(dy2, weight.t().contiguous(), None, "none")
    return dy2 @ weight


class _LinearBiasActFn(torch.autograd.Function):
    """act(x W^T + b) with this library's forward GEMM (FusedDenseFunc, fused_dense.py:39-80, and the fc1 half of
    FusedDenseGeluDenseFunc, :179-300).  With an activation the forward kernel also stores the pre-activation through a
    second output map (the reference's save_gelu_in / checkpoint_lvl 0-1, fused_dense.py:220-222, 262-266) unless
    `save_pre_act` is False or there are fewer than 256 rows, in which case the backward recomputes it by one more GEMM
    (checkpoint_lvl = 2).  dgelu + bias gradient are one pass of bp_bias_act_bwd; dgrad on this library's GEMM, wgrad on
    cuBLAS."""

    @staticmethod
    def forward(ctx, x, weight, bias, activation, save_pre_act=True):
        n, k = weight.shape
        pre = None
        with torch.no_grad():
            if activation != "none" and save_pre_act and x.numel() // k >= 256:
                x2 = x.reshape(-1, k)
                x2 = x2 if x2.is_contiguous() else x2.contiguous()
                out, pre = _linear_bias_act_aux(x2, weight.contiguous(), bias, activation)
                out = out.reshape(*x.shape[:-1], n)
            else:
                out = linear_bias_act(x, weight, bias, activation)
        ctx.save_for_backward(x, weight, bias, pre)
        ctx.activation = activation
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, bias, pre = ctx.saved_tensors
        n, k = weight.shape
        x2 = x.reshape(-1, k)
        d2 = dout.reshape(-1, n)
        want_db = bias is not None and ctx.needs_input_grad[2]
        if ctx.activation == "gelu_tanh" and pre is None:
            with torch.no_grad():
                pre = linear_bias_act(x2, weight, bias, "none")
        dpre, dbias = bias_act_backward(d2, pre, ctx.activation, want_db)
        dx = _dgrad(dpre, weight).reshape(x.shape) if ctx.needs_input_grad[0] else None
        dw = dpre.t() @ x2 if ctx.needs_input_grad[1] else None
        return dx, dw, dbias, None, None


def linear_bias_residual_(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None,
                          residual: torch.Tensor) -> torch.Tensor:
    """residual += x @ weight.T + bias, in place, on the fp32 residual stream (bp_linear_bias_residual_fwd): the
    out_proj / fc2 GEMM of a pre-norm block with the "add" of dropout_add_layer_norm (block.py:84-88,101-105)
    in its epilogue.  Returns `residual`."""
    _lib.require_cuda(x, weight, bias, residual)
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype:
        raise RuntimeError("linear_bias_residual_ needs fp16/bf16 activations and weights of the same dtype")
    if bias is not None and bias.dtype != x.dtype:
        raise RuntimeError("bias must have the activation dtype")
    n, k = weight.shape
    bias = _dense_bias(bias, n)
    if x.shape[-1] != k or residual.shape[-1] != n or residual.numel() // n != x.numel() // k:
        raise RuntimeError("shape mismatch between x, weight and residual")
    if residual.dtype != torch.float32 or not residual.is_contiguous():
        raise RuntimeError("the residual stream must be a contiguous fp32 tensor")
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or residual.requires_grad):
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    x2 = x.reshape(-1, k)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    weight = weight.contiguous()
    with torch.cuda.device(x.device):
        st = _lib.load().bp_linear_bias_residual_fwd(x2.data_ptr(), weight.data_ptr(), _lib.ptr(bias),
                                                     residual.data_ptr(), x2.shape[0], n, k,
                                                     _lib.dtype_code(x.dtype), _lib.stream_ptr(x.device))
    _lib.check(st, "bp_linear_bias_residual_fwd")
    return residual


def can_fuse_residual(x: torch.Tensor, weight: torch.Tensor, residual: torch.Tensor | None) -> bool:
    """The epilogue-add GEMM needs a CUDA fp32 residual, 16-bit operands and at least one 256-row tile."""
    return (residual is not None and residual.is_cuda and residual.dtype == torch.float32 and residual.is_contiguous()
            and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x.dtype
            and x.numel() // x.shape[-1] >= 256 and not torch.is_grad_enabled())


def _own_gemm_ok(x, weight, bias) -> bool:
    """Conditions of bp_linear_bias_act_fwd; anything else is the reference's F.linear."""
    if not (x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x.dtype):
        return False
    if bias is not None and (bias.dtype != x.dtype or bias.shape != (weight.shape[0],)):
        return False
    return not (weight.shape[0] % 8 or weight.shape[1] % 8 or x.numel() == 0)


# Which GEMM serves the plain linears on the inference path: "own" (bp_linear_bias_act_fwd), "library" (F.linear ->
# cuBLAS, what the reference calls), a per-shape choice {(n, k): "own" | "library"}, or "auto" (default):
#   "auto" = own GEMM for every linear with at least 512 rows and n <= 16384 -- all of the model but the tied LM head
#   -- and the library for the LM head and for the skinny GEMMs of incremental decoding (a handful of rows against the
#   whole weight matrix is a bandwidth-bound GEMV-like problem cuBLAS has split-K kernels for).
# Measured on B200 (profiles/): the step runs under the 1000 W power cap, so what counts is throughput at that
# operating point (benchmarks/gemm_sustained.py: own >= cuBLAS at every shape but fc2, -3 %) and, in the end, the
# whole forward replayed as one CUDA graph (benchmarks/linear_policy_ab.py): all-library 24.9 ms; own for one shape
# only: Wqkv 24.8, out_proj 25.0, fc2 25.0, contextualisation Wqkv 24.9, content projection 24.9, LM head 25.5.
# set_linear_backend() overrides (tests, bench.py).
_backend = "auto"
_timing_hook = None      # bench.py: callable(tag, n, k) -> context manager bracketing the library GEMM with CUDA events


def set_linear_backend(backend) -> None:
    global _backend
    if not (backend in ("auto", "own", "library") or isinstance(backend, dict)):
        raise ValueError('linear backend must be "auto", "own", "library" or a {(n, k): backend} dict')
    _backend = backend


def get_linear_backend():
    return _backend


def _choice(n: int, k: int, m: int) -> str:
    if isinstance(_backend, dict):
        return _backend.get((n, k), "library")
    if _backend == "auto":
        return "own" if (m >= 512 and n <= 16384) else "library"
    return _backend


def linear(x, weight, bias=None):
    """x @ weight.T + bias: this library's GEMM on the inference path, F.linear otherwise (see the module docstring)."""
    n, k = weight.shape
    if _choice(n, k, x.numel() // max(k, 1)) == "own" and _own_gemm_ok(x, weight, bias):
        return linear_bias_act(x, weight, bias, "none")
    if _timing_hook is not None and x.is_cuda:
        with _timing_hook("F.linear", n, k):
            return F.linear(x, weight, bias)
    return F.linear(x, weight, bias)


def fused_dense_func(x, weight, bias=None, return_residual=False, process_group=None):
    if process_group is not None:
        raise RuntimeError("tensor parallelism is out of scope for this path (batch sharding only)")
    out = linear(x, weight, bias)
    return out if not return_residual else (out, x)


class FusedDense(nn.Linear):

    def __init__(self, in_features: int, out_features: int, bias: bool = True, return_residual: bool = False,
                 device=None, dtype=None) -> None:
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        self.return_residual = return_residual

    def forward(self, x, process_group=None):
        return fused_dense_func(x, self.weight, self.bias, return_residual=self.return_residual,
                                process_group=process_group)


def fused_dense_gelu_dense_func(x, weight1, weight2, bias1=None, bias2=None, save_pre_act=False,
                                return_residual=False, checkpoint_lvl=0, heuristic=0, process_group=None):
    """fc2(gelu_tanh(fc1(x))) (fused_dense.py:332-354).  checkpoint_lvl 0 / 1: the fc1 kernel stores the pre-activation
    for the backward; 2: the backward recomputes it (fused_dense.py:262-266).  `save_pre_act` (an inference-path flag of
    the reference) and `heuristic` (its cuBLASLt algorithm choice) are accepted for signature compatibility."""
    if process_group is not None:
        raise RuntimeError("tensor parallelism is out of scope for this path (batch sharding only)")
    hidden = linear_bias_act(x, weight1, bias1, "gelu_tanh", save_pre_act=checkpoint_lvl != 2)
    out = linear(hidden, weight2, bias2)
    return out if not return_residual else (out, x)


class FusedDenseGeluDense(nn.Module):

    def __init__(self, in_features, hidden_features, out_features=None, bias1=True, bias2=True,
                 return_residual=False, checkpoint_lvl=0, heuristic=0, device=None, dtype=None):
        if checkpoint_lvl not in (0, 1, 2):
            raise ValueError("checkpoint_lvl must be 0, 1 or 2")
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        if out_features is None:
            out_features = in_features
        self.return_residual = return_residual
        self.checkpoint_lvl = checkpoint_lvl
        self.heuristic = heuristic
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias1, **factory_kwargs)
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias2, **factory_kwargs)

    def forward_into_residual(self, x, residual):
        """residual += fc2(gelu_tanh(fc1(x))) with the add in the fc2 epilogue; returns residual."""
        hidden = linear_bias_act(x, self.fc1.weight, self.fc1.bias, "gelu_tanh")
        return linear_bias_residual_(hidden, self.fc2.weight, self.fc2.bias, residual)

    def forward(self, x, process_group=None):
        return fused_dense_gelu_dense_func(x, self.fc1.weight, self.fc2.weight, self.fc1.bias, self.fc2.bias,
                                           save_pre_act=self.training, return_residual=self.return_residual,
                                           checkpoint_lvl=self.checkpoint_lvl, heuristic=self.heuristic,
                                           process_group=process_group)
