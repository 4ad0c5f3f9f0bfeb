"""ctypes binding of libbackpack_b200.so (the C ABI declared in include/backpack_b200.h).

There is no fallback: if the shared library is missing, or a call returns a non-zero status, this module
raises.  Build the library with ``python -m backpacks_flash_attn_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_TAG = os.environ.get("BP_LIB_TAG", "")   # debug builds only (see build.py: BP_BUILD_TAG)
LIB_PATH = os.path.join(_PKG, "libbackpack_b200" + (f"_{_TAG}" if _TAG else "") + ".so")

BP_DTYPE_F16, BP_DTYPE_BF16, BP_DTYPE_F32 = 0, 1, 2
BP_ACT_NONE, BP_ACT_GELU_TANH = 0, 1

_DTYPES = {torch.float16: BP_DTYPE_F16, torch.bfloat16: BP_DTYPE_BF16, torch.float32: BP_DTYPE_F32}

# name -> argtypes; every function returns int (bp_status_t) unless listed in _RESTYPES
SIGNATURES = {
    "bp_abi_version": [],
    "bp_last_error": [],
    "bp_check_device": [],
    "bp_fmha_fwd": [c_void_p] * 7 + [c_int32] * 7 + [c_int64] * 8 + [c_int32, c_float, c_int32, c_int32, c_void_p],
    "bp_fmha_fwd_dropout_workspace_bytes": [c_int32, c_int32, c_int32],
    "bp_fmha_fwd_dropout": [c_void_p] * 7 + [c_int32] * 7 + [c_int64] * 8 + [c_int32, c_float, c_int32, c_int32, c_float,
                                                                            c_uint64, c_void_p, c_int64, c_void_p],
    "bp_fmha_bwd_dropout_workspace_bytes": [c_int32, c_int32, c_int32, c_int32],
    "bp_fmha_bwd_dropout": [c_void_p] * 11 + [c_int32] * 7 + [c_void_p, c_int32, c_float, c_int32, c_int32, c_float,
                                                               c_uint64, c_void_p, c_int64, c_void_p],
    "bp_fmha_bwd_workspace_bytes": [c_int32, c_int32, c_int32],
    "bp_fmha_bwd": [c_void_p] * 11 + [c_int32] * 7 + [c_void_p, c_int32, c_float, c_int32, c_int32, c_void_p, c_int64,
                                                       c_void_p],
    "bp_sense_lse_fwd": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p],
    "bp_sense_mix_fwd": [c_void_p] * 4 + [c_int32] * 5 + [c_int64] * 3 + [c_float, c_int32, c_void_p],
    "bp_sense_mix_table_fwd": [c_void_p] * 5 + [c_int32] * 6 + [c_float, c_int32, c_void_p],
    "bp_ln_residual_fwd": [c_void_p] * 8 + [c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_void_p],
    "bp_linear_bias_act_fwd": [c_void_p] * 4 + [c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "bp_linear_bias_act_aux_fwd": [c_void_p] * 5 + [c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "bp_lm_head_stats_fwd": [c_void_p] * 7 + [c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "bp_decode_attn_fwd": [c_void_p] * 4 + [c_int32] * 4 + [c_int64] * 3 + [c_float, c_int32, c_void_p],
    "bp_sense_mix_decode_fwd": [c_void_p] * 6 + [c_int32] * 6 + [c_int64] * 2 + [c_float, c_int32, c_void_p],
    "bp_linear_bias_residual_fwd": [c_void_p] * 4 + [c_int64, c_int32, c_int32, c_int32, c_void_p],
    "bp_ln_fwd": [c_void_p] * 6 + [c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_void_p],
    "bp_ln_residual_fwd_dropout": [c_void_p] * 8 + [c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_float, c_uint64,
                                   c_void_p],
    "bp_ln_residual_bwd_dropout": [c_void_p] * 11 + [c_int64, c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_float,
                                                     c_uint64, c_void_p],
    "bp_ln_bwd_workspace_bytes": [c_int32],
    "bp_ln_residual_bwd": [c_void_p] * 11 + [c_int64, c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_void_p],
    "bp_bias_act_bwd_workspace_bytes": [c_int32],
    "bp_bias_act_bwd": [c_void_p] * 5 + [c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p],
    "bp_xentropy_fwd": [c_void_p] * 4 + [c_int64, c_int32, c_int64, c_float, c_int64, c_int32, c_int32, c_void_p],
    "bp_xentropy_bwd": [c_void_p] * 5 + [c_int64, c_int32, c_int64, c_int64, c_float, c_int64, c_int32, c_int32, c_void_p],
    "bp_sense_softmax_bwd": [c_void_p, c_void_p, c_int64, c_int32, c_float, c_int32, c_void_p],
    "bp_rotary_qk_inplace": [c_void_p] * 5 + [c_int32] * 6 + [c_void_p],
}
_RESTYPES = {"bp_last_error": c_char_p, "bp_fmha_bwd_workspace_bytes": c_int64,
             "bp_fmha_fwd_dropout_workspace_bytes": c_int64, "bp_fmha_bwd_dropout_workspace_bytes": c_int64,
             "bp_ln_bwd_workspace_bytes": c_int64, "bp_bias_act_bwd_workspace_bytes": c_int64}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library is required (no CPU/PyTorch fallback exists). "
                "Build it with `python -m backpacks_flash_attn_b200.build`.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == header/library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if lib.bp_abi_version() != 1:
            raise RuntimeError("libbackpack_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().bp_last_error().decode()


# Every successful C-ABI call launches exactly one kernel of this library; bench.py reads these counters
# for its `gpu_launches` claim.
launch_counts: dict[str, int] = {}


def check(status: int, what: str, launched: bool = True) -> None:
    """Raise on a non-zero status; count the call as one kernel launch unless `launched` is False (queries)."""
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")
    if launched:
        launch_counts[what] = launch_counts.get(what, 0) + 1


def total_launches() -> int:
    return sum(launch_counts.values())


class KernelTimer:
    """Optional CUDA-event bracket around one C-ABI entry point (used by bench.py to time every kernel of this library
    live inside an eager pass, on the launching stream).  `key_fn(args)` groups the launches of one entry point
    (e.g. the GEMM by its (n, k, activation)).  Usage:
        with KernelTimer("bp_fmha_fwd") as t: ...run steps...;  t.mean_ms(), t.by_key()"""

    def __init__(self, name: str, key_fn=None):
        self.name = name
        self.key_fn = key_fn
        self.events: list = []
        self._orig = None

    def __enter__(self):
        lib = load()
        self._orig = getattr(lib, self.name)
        orig, events, key_fn = self._orig, self.events, self.key_fn

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            st = orig(*args)
            b.record()
            events.append((a, b, key_fn(args) if key_fn else None))
            return st

        setattr(lib, self.name, timed)
        return self

    def __exit__(self, *exc):
        setattr(load(), self.name, self._orig)
        return False

    def mean_ms(self) -> float:
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b, _ in self.events) / max(1, len(self.events))

    def by_key(self) -> dict:
        """{key: (launches, total ms)}"""
        torch.cuda.synchronize()
        out: dict = {}
        for a, b, k in self.events:
            n, t = out.get(k, (0, 0.0))
            out[k] = (n + 1, t + a.elapsed_time(b))
        return out


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {dtype}") from None


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("backpacks_flash_attn_b200 operators need CUDA tensors on a B200 "
                               "(there is no CPU path; the CPU oracle lives under oracle/ for tests only)")
