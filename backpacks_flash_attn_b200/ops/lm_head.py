"""LM head without logits: softmax statistics fused into the GEMM epilogue (bp_lm_head_stats_fwd).

The reference projects every position to the vocabulary (`lm_head`, training/src/models/backpack.py:349) and then
either takes a cross-entropy over those logits (flash_attn/losses/cross_entropy.py:19-129 ->
csrc/xentropy/xentropy_kernel.cu:430-760; training/src/metrics/perplexity.py) or, when generating, keeps
`logits[:, -1]` only (training/src/utils/generation.py:34-44).  At Backpack-Small / batch 64 / seq 1024 the logits are
6.6 GB per forward.  `lm_head_stats` returns what those consumers need -- per-row log-sum-exp, arg-max, max logit and the
target's logit -- straight from the GEMM's fp32 accumulators; nothing of size (rows, vocab) is written.
"""
from __future__ import annotations

import torch

from .. import _lib


def lm_head_stats(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor | None = None,
                  n_valid: int | None = None) -> dict:
    """hidden: (..., d) fp16/bf16; weight: (vocab, d) (the tied embedding matrix); targets: (...) int64 or None.
    Returns {"lse", "argmax", "max_logit"[, "target_logit"]} with hidden's leading shape (fp32 / int32)."""
    _lib.require_cuda(hidden, weight, targets)
    if hidden.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != hidden.dtype:
        raise RuntimeError("lm_head_stats needs fp16/bf16 hidden states and weights of the same dtype")
    n, k = weight.shape
    if hidden.shape[-1] != k:
        raise RuntimeError("shape mismatch between hidden and weight")
    if torch.is_grad_enabled() and (hidden.requires_grad or weight.requires_grad):
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    lead = hidden.shape[:-1]
    x = hidden.reshape(-1, k)
    if not x.is_contiguous():
        x = x.contiguous()
    weight = weight.contiguous()
    m = x.shape[0]
    if targets is not None:
        if targets.dtype != torch.int64 or targets.shape != lead:
            raise RuntimeError(f"targets must be int64 of shape {tuple(lead)}")
        targets = targets.reshape(-1).contiguous()
    n_valid = n if n_valid is None else int(n_valid)
    dev = hidden.device
    lse = torch.empty(m, dtype=torch.float32, device=dev)
    amax = torch.empty(m, dtype=torch.int32, device=dev)
    mlog = torch.empty(m, dtype=torch.float32, device=dev)
    tlog = torch.empty(m, dtype=torch.float32, device=dev) if targets is not None else None
    with torch.cuda.device(dev):
        st = _lib.load().bp_lm_head_stats_fwd(x.data_ptr(), weight.data_ptr(), _lib.ptr(targets), lse.data_ptr(),
                                              amax.data_ptr(), mlog.data_ptr(), _lib.ptr(tlog), m, n, k, n_valid,
                                              _lib.dtype_code(hidden.dtype), _lib.stream_ptr(dev))
    _lib.check(st, "bp_lm_head_stats_fwd")
    out = {"lse": lse.reshape(lead), "argmax": amax.reshape(lead), "max_logit": mlog.reshape(lead)}
    if tlog is not None:
        out["target_logit"] = tlog.reshape(lead)
    return out


def lm_head_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor, ignore_index: int = -100,
                          reduction: str = "mean") -> torch.Tensor:
    """Cross-entropy of softmax(hidden @ weight.T) against `targets` with torch.nn.CrossEntropyLoss's `ignore_index` /
    `reduction` semantics (what flash_attn/losses/cross_entropy.py wraps), without materialising the logits."""
    if reduction not in ("mean", "sum", "none"):
        raise ValueError("reduction must be 'mean', 'sum' or 'none'")
    keep = targets != ignore_index
    st = lm_head_stats(hidden, weight, torch.where(keep, targets, torch.zeros_like(targets)))
    loss = torch.where(keep, st["lse"] - st["target_logit"], torch.zeros_like(st["lse"]))
    if reduction == "none":
        return loss
    if reduction == "sum":
        return loss.sum()
    return loss.sum() / keep.sum().clamp(min=1)
