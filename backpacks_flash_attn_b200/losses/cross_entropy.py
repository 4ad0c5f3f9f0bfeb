"""`CrossEntropyLoss` with the reference's interface (flash_attn/losses/cross_entropy.py:19-129), running
bp_xentropy_fwd / bp_xentropy_bwd instead of xentropy_cuda_lib: one pass over the logits forward, one pass backward,
optionally in place over the logits (`inplace_backward=True`: the (tokens, vocab) tensor is not duplicated).
Tensor-parallel vocabularies (`process_group`) are out of scope for this path (batch sharding only)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib


class SoftmaxCrossEntropyLossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, labels, smoothing=0.0, ignored_index=-100, inplace_backward=False, process_group=None):
        """logits: (batch, vocab_size) fp16 / bf16 / fp32 with unit last stride; labels: (batch,) int64.
        Returns the per-row losses in fp32 (the kernel accumulates in fp32, as the reference's does)."""
        if process_group is not None:
            raise RuntimeError("tensor-parallel cross entropy is out of scope for this path (batch sharding only)")
        _lib.require_cuda(logits, labels)
        if logits.dim() != 2 or labels.shape != (logits.shape[0],):
            raise RuntimeError("logits must be (batch, vocab) and labels (batch,)")
        if labels.dtype != torch.int64:
            raise RuntimeError("labels must be int64")
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        labels = labels.contiguous()
        batch, vocab = logits.shape
        losses = torch.empty(batch, dtype=torch.float32, device=logits.device)
        lse = torch.empty(batch, dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            st = _lib.load().bp_xentropy_fwd(logits.data_ptr(), labels.data_ptr(), losses.data_ptr(), lse.data_ptr(),
                                             batch, vocab, logits.stride(0), float(smoothing), int(ignored_index), -1,
                                             _lib.dtype_code(logits.dtype), _lib.stream_ptr(logits.device))
        _lib.check(st, "bp_xentropy_fwd")
        ctx.save_for_backward(logits, lse, labels)
        ctx.smoothing, ctx.ignored_index, ctx.inplace_backward = float(smoothing), int(ignored_index), inplace_backward
        return losses

    @staticmethod
    def backward(ctx, grad_loss):
        logits, lse, labels = ctx.saved_tensors
        grad_loss = grad_loss.contiguous().float()
        grad_logits = logits if ctx.inplace_backward else torch.empty_like(logits)
        batch, vocab = logits.shape
        with torch.cuda.device(logits.device):
            st = _lib.load().bp_xentropy_bwd(grad_loss.data_ptr(), logits.data_ptr(), lse.data_ptr(), labels.data_ptr(),
                                             grad_logits.data_ptr(), batch, vocab, logits.stride(0), grad_logits.stride(0),
                                             ctx.smoothing, ctx.ignored_index, -1, _lib.dtype_code(logits.dtype),
                                             _lib.stream_ptr(logits.device))
        _lib.check(st, "bp_xentropy_bwd")
        return grad_logits, None, None, None, None, None


class CrossEntropyLoss(nn.Module):
    """flash_attn/losses/cross_entropy.py:112-129.  reduction: 'mean' (over the non-ignored targets) or 'none'."""

    def __init__(self, ignore_index=-100, reduction="mean", label_smoothing=0.0, inplace_backward=False,
                 process_group=None):
        super().__init__()
        if reduction not in ("mean", "none"):
            raise NotImplementedError("Only support reduction = 'mean' or 'none'")
        if process_group is not None:
            raise RuntimeError("tensor-parallel cross entropy is out of scope for this path (batch sharding only)")
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.label_smoothing = label_smoothing
        self.inplace_backward = inplace_backward
        self.process_group = None

    def forward(self, input, target):
        if not (input.is_cuda and target.is_cuda):
            raise RuntimeError("CrossEntropyLoss needs CUDA tensors (there is no CPU path)")
        loss = SoftmaxCrossEntropyLossFn.apply(input, target, self.label_smoothing, self.ignore_index,
                                               self.inplace_backward, None)
        if self.reduction == "mean":
            return loss.sum() / (target != self.ignore_index).sum()
        return loss


CrossEntropyLossApex = CrossEntropyLoss   # the name the reference's test imports (tests/losses/test_cross_entropy.py:9)
