"""Rotary position embedding with the reference's interface (flash_attn/layers/rotary.py:78-201),
applied in place to q and k of a packed qkv by bp_rotary_qk_inplace."""
from __future__ import annotations

import torch

from .. import _lib


def apply_rotary_emb_qkv_(qkv, cos, sin, cos_k=None, sin_k=None):
    """qkv: (batch, seqlen, 3, nheads, headdim), modified in place and returned; cos, sin (and the optional
    XPos tables for k): (seqlen, rotary_dim / 2) in qkv's dtype (rotary.py:81-105)."""
    _lib.require_cuda(qkv, cos, sin, cos_k, sin_k)
    if qkv.dim() != 5 or qkv.shape[2] != 3:
        raise RuntimeError("qkv must be (batch, seqlen, 3, nheads, headdim)")
    if not qkv.is_contiguous():
        raise RuntimeError("qkv must be contiguous")
    b, s, _, h, d = qkv.shape
    rotary_seqlen, half = cos.shape
    rotary_dim = half * 2
    if rotary_dim > d:
        raise RuntimeError("rotary_dim must be <= headdim")                      # rotary.py:91
    if s > rotary_seqlen:
        raise RuntimeError("cos/sin tables shorter than the sequence")          # rotary.py:92
    if sin.shape != cos.shape or cos.dtype != qkv.dtype or sin.dtype != qkv.dtype:
        raise RuntimeError("cos and sin must have the same shape and qkv's dtype")
    if torch.is_grad_enabled() and qkv.requires_grad:
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    cos, sin = cos[:s].contiguous(), sin[:s].contiguous()
    if cos_k is not None:
        cos_k, sin_k = cos_k[:s].contiguous(), sin_k[:s].contiguous()
    with torch.cuda.device(qkv.device):
        st = _lib.load().bp_rotary_qk_inplace(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), _lib.ptr(cos_k),
                                              _lib.ptr(sin_k), b, s, h, d, rotary_dim, _lib.dtype_code(qkv.dtype),
                                              _lib.stream_ptr(qkv.device))
    _lib.check(st, "bp_rotary_qk_inplace")
    return qkv


class RotaryEmbedding(torch.nn.Module):
    """RoFormer rotary embedding with optional XPos scaling; caches cos/sin tables in the activation dtype
    exactly as the reference does (rotary.py:125-201): angles and the XPos scale are computed in fp32 and
    rounded once."""

    def __init__(self, dim: int, base=10000, scale_base=0, device=None):
        super().__init__()
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
        self.register_buffer("inv_freq", inv_freq)
        self.scale_base = scale_base
        scale = ((torch.arange(0, dim, 2, device=device, dtype=torch.float32) + 0.4 * dim) / (1.4 * dim)
                 if scale_base > 0 else None)
        self.register_buffer("scale", scale)
        self._seq_len_cached = 0
        self._cos_cached = self._sin_cached = self._cos_k_cached = self._sin_k_cached = None

    def _update_cos_sin_cache(self, x, seqlen_offset=0):
        seqlen = x.shape[1] + seqlen_offset
        if (seqlen > self._seq_len_cached or self._cos_cached is None or self._cos_cached.device != x.device
                or self._cos_cached.dtype != x.dtype):
            self._seq_len_cached = seqlen
            t = torch.arange(seqlen, device=x.device, dtype=self.inv_freq.dtype)
            freqs = torch.outer(t, self.inv_freq.to(device=t.device))
            if self.scale is None:
                self._cos_cached = torch.cos(freqs).to(x.dtype)
                self._sin_cached = torch.sin(freqs).to(x.dtype)
            else:
                power = (torch.arange(seqlen, dtype=self.scale.dtype, device=self.scale.device)
                         - seqlen // 2) / self.scale_base
                scale = self.scale.to(device=power.device) ** power.unsqueeze(1)
                self._cos_cached = (torch.cos(freqs) * scale).to(x.dtype)
                self._sin_cached = (torch.sin(freqs) * scale).to(x.dtype)
                self._cos_k_cached = (torch.cos(freqs) / scale).to(x.dtype)
                self._sin_k_cached = (torch.sin(freqs) / scale).to(x.dtype)

    def forward(self, qkv: torch.Tensor, seqlen_offset: int = 0):
        self._update_cos_sin_cache(qkv, seqlen_offset)
        if self.scale is None:
            return apply_rotary_emb_qkv_(qkv, self._cos_cached[seqlen_offset:], self._sin_cached[seqlen_offset:])
        return apply_rotary_emb_qkv_(qkv, self._cos_cached[seqlen_offset:], self._sin_cached[seqlen_offset:],
                                     self._cos_k_cached[seqlen_offset:], self._sin_k_cached[seqlen_offset:])
