"""Fused Backpack sense-mix operator.

    out[b, i, :] = sum_l sum_{j<=i} softmax_j(scale * q_li . k_lj) * content[b, l, j, :]

replaces `torch.sum(contextualization @ content, dim=1)` together with the softmax half of
`ContextSelfAttn.forward` (training/src/models/backpack.py:116-122, 313).  alpha (b, nv, s, s) is never
materialised.  `content` may be ANY (b, nv, s, d) tensor with unit last stride -- in particular the
transposed view the reference's content model returns (backpack.py:276) and the edited sense tensors
the intervention wrappers build (training/src/models/intervened_models.py:78-101).
"""
from __future__ import annotations

import torch

from .. import _lib


def sense_mix(qk: torch.Tensor, content: torch.Tensor, softmax_scale: float | None = None,
              return_lse: bool = False):
    """qk: (batch, seqlen, 2, nv, dk) -- ContextSelfAttn.Wqkv output viewed as at backpack.py:112-116;
    content: (batch, nv, seqlen, d).  Returns (batch, seqlen, d) [and lse (batch, nv, seqlen) fp32]."""
    _lib.require_cuda(qk, content)
    if qk.dim() != 5 or qk.shape[2] != 2:
        raise RuntimeError("qk must be (batch, seqlen, 2, nv, dk)")
    if qk.dtype not in (torch.float16, torch.bfloat16) or content.dtype != qk.dtype:
        raise RuntimeError("sense_mix needs fp16/bf16 qk and content of the same dtype")
    b, s, _, nv, dk = qk.shape
    if content.dim() != 4 or content.shape[:3] != (b, nv, s):
        raise RuntimeError(f"content must be (batch, nv, seqlen, d) = ({b}, {nv}, {s}, d), got {tuple(content.shape)}")
    if torch.is_grad_enabled() and (qk.requires_grad or content.requires_grad):
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    d = content.shape[3]
    if dk % 8 != 0:
        # TMA needs 16-byte row strides.  Zero-padding the sense key width (e.g. d/nv = 12 at k = 64 senses of a
        # 768-wide model) leaves every q.k dot product unchanged; the softmax scale keeps using the true width.
        if softmax_scale is None:
            softmax_scale = dk ** -0.5
        qk = torch.nn.functional.pad(qk, (0, (-dk) % 8))
        dk = qk.shape[-1]
    if not qk.is_contiguous():
        qk = qk.contiguous()
    if content.stride(3) != 1 or any(st % 8 for st in content.stride()[:3]) or content.data_ptr() % 16:
        content = content.contiguous()
    scale = float(softmax_scale) if softmax_scale is not None else dk ** -0.5
    lse = torch.empty((b, nv, s), dtype=torch.float32, device=qk.device)
    out = torch.empty((b, s, d), dtype=qk.dtype, device=qk.device)
    lib = _lib.load()
    dt = _lib.dtype_code(qk.dtype)
    with torch.cuda.device(qk.device):
        stream = _lib.stream_ptr(qk.device)
        _lib.check(lib.bp_sense_lse_fwd(qk.data_ptr(), lse.data_ptr(), b, s, nv, dk, scale, dt, stream),
                   "bp_sense_lse_fwd")
        _lib.check(lib.bp_sense_mix_fwd(qk.data_ptr(), content.data_ptr(), lse.data_ptr(), out.data_ptr(),
                                        b, s, nv, dk, d, content.stride(0), content.stride(1), content.stride(2),
                                        scale, dt, stream), "bp_sense_mix_fwd")
    return (out, lse) if return_lse else out
