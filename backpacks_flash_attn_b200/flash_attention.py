"""`FlashAttention` / `FlashMHA` modules with the reference's constructor and forward signatures
(flash_attn/flash_attention.py:11-101), running on bp_fmha_fwd.

The key_padding_mask branch of the reference (flash_attention.py:52-63) goes through bert_padding's
unpad/pad; here the mask is turned into cu_seqlens directly and the kernel's varlen path is used, which
covers right-padded batches (the only kind the reference's own tests generate,
tests/test_flash_attn.py:25-40).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .flash_attn_interface import flash_attn_unpadded_qkvpacked_func


class FlashAttention(nn.Module):
    """Scaled dot-product attention with softmax; qkv (B, S, 3, H, D) or unpadded (nnz, 3, H, D)."""

    def __init__(self, softmax_scale=None, attention_dropout=0.0, device=None, dtype=None):
        super().__init__()
        self.softmax_scale = softmax_scale
        self.dropout_p = attention_dropout

    def forward(self, qkv, key_padding_mask=None, causal=False, cu_seqlens=None, max_s=None,
                need_weights=False):
        if need_weights:
            raise RuntimeError("need_weights is not supported")
        if qkv.dtype not in (torch.float16, torch.bfloat16) or not qkv.is_cuda:
            raise RuntimeError("FlashAttention needs fp16/bf16 CUDA tensors")
        p = self.dropout_p if self.training else 0.0
        if cu_seqlens is not None:
            if max_s is None:
                raise RuntimeError("max_s is required together with cu_seqlens")
            out = flash_attn_unpadded_qkvpacked_func(qkv, cu_seqlens, max_s, p, softmax_scale=self.softmax_scale,
                                                     causal=causal)
            return out, None
        b, s = qkv.shape[:2]
        if key_padding_mask is None:
            cu = torch.arange(0, (b + 1) * s, step=s, dtype=torch.int32, device=qkv.device)
            out = flash_attn_unpadded_qkvpacked_func(qkv.reshape(b * s, *qkv.shape[2:]), cu, s, p,
                                                     softmax_scale=self.softmax_scale, causal=causal)
            return out.reshape(b, s, *out.shape[1:]), None
        # right-padded batch: gather the kept tokens, run varlen, scatter back (zeros at padded slots)
        keep = key_padding_mask.bool()
        lens = keep.sum(dim=1, dtype=torch.int32)
        idx = torch.nonzero(keep.flatten(), as_tuple=False).flatten()
        cu = torch.zeros(b + 1, dtype=torch.int32, device=qkv.device)
        cu[1:] = torch.cumsum(lens, 0)
        packed = qkv.reshape(b * s, *qkv.shape[2:]).index_select(0, idx)
        out_unpad = flash_attn_unpadded_qkvpacked_func(packed, cu, int(lens.max()), p,
                                                       softmax_scale=self.softmax_scale, causal=causal)
        out = torch.zeros(b * s, *out_unpad.shape[1:], dtype=out_unpad.dtype, device=qkv.device)
        out.index_copy_(0, idx, out_unpad)
        return out.reshape(b, s, *out.shape[1:]), None


class FlashMHA(nn.Module):
    """Wqkv -> FlashAttention -> out_proj (flash_attention.py:74-101).  State-dict keys: Wqkv.*, out_proj.*"""

    def __init__(self, embed_dim, num_heads, bias=True, batch_first=True, attention_dropout=0.0, causal=False,
                 device=None, dtype=None, **kwargs) -> None:
        if not batch_first:
            raise RuntimeError("FlashMHA only supports batch_first=True")
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.embed_dim = embed_dim
        self.causal = causal
        self.num_heads = num_heads
        if embed_dim % num_heads != 0:
            raise RuntimeError("embed_dim must be divisible by num_heads")
        self.head_dim = embed_dim // num_heads
        if self.head_dim % 8 != 0 or self.head_dim > 128:
            raise RuntimeError("Only support head_dim <= 128 and divisible by 8")
        self.Wqkv = nn.Linear(embed_dim, 3 * embed_dim, bias=bias, **factory_kwargs)
        self.inner_attn = FlashAttention(attention_dropout=attention_dropout, **factory_kwargs)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias, **factory_kwargs)

    def forward(self, x, key_padding_mask=None, need_weights=False):
        b, s, _ = x.shape
        qkv = self.Wqkv(x).reshape(b, s, 3, self.num_heads, self.head_dim)
        ctx, w = self.inner_attn(qkv, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                 causal=self.causal)
        return self.out_proj(ctx.reshape(b, s, self.embed_dim)), w
