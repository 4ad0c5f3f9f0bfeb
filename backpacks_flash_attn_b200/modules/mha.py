"""Self-attention modules with the reference's interface (flash_attn/modules/mha.py):
`FlashSelfAttention` (:34-100) on bp_fmha_fwd, the eager `SelfAttention` (:179-224) that
`use_flash_attn=False` selects, and `MHA` (:287-467, self-attention branch).

`MHA.forward(..., inference_params)` is the KV-cache path of mha.py:356-380, 432-440: the prompt pass fills the cache,
every further single-token step attends to the whole cache (bp_decode_attn_fwd with use_flash_attn, the eager
non-causal cross-attention of mha.py:226-268 otherwise).

Out of scope here, as in SURVEY.md §2.1 #4: cross-attention between different sequences, depth-wise conv and the
tensor-parallel ParallelMHA.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..flash_attn_interface import flash_attn_unpadded_qkvpacked_func
from ..layers.rotary import RotaryEmbedding
from ..ops.decode import decode_attention
from ..ops.fused_dense import FusedDense, linear_bias_residual_


class FlashSelfAttention(nn.Module):
    """Scaled dot-product attention on the fused kernel.  qkv: (B, S, 3, H, D), or (total, 3, H, D) together
    with cu_seqlens / max_seqlen."""

    def __init__(self, causal=False, softmax_scale=None, attention_dropout=0.0, triton=False):
        super().__init__()
        if triton:
            raise RuntimeError("the Triton path is out of scope")
        self.causal = causal
        self.softmax_scale = softmax_scale
        self.dropout_p = attention_dropout
        self.triton = False

    def forward(self, qkv, causal=None, cu_seqlens=None, max_seqlen=None):
        if qkv.dtype not in (torch.float16, torch.bfloat16) or not qkv.is_cuda:
            raise RuntimeError("FlashSelfAttention needs fp16/bf16 CUDA tensors")
        causal = self.causal if causal is None else causal
        p = self.dropout_p if self.training else 0.0
        if cu_seqlens is not None:
            if cu_seqlens.dtype != torch.int32 or not isinstance(max_seqlen, int):
                raise RuntimeError("cu_seqlens must be int32 and max_seqlen an int")
            return flash_attn_unpadded_qkvpacked_func(qkv, cu_seqlens, max_seqlen, p,
                                                      softmax_scale=self.softmax_scale, causal=causal)
        b, s = qkv.shape[:2]
        cu = torch.arange(0, (b + 1) * s, step=s, dtype=torch.int32, device=qkv.device)
        out = flash_attn_unpadded_qkvpacked_func(qkv.reshape(b * s, *qkv.shape[2:]), cu, s, p,
                                                 softmax_scale=self.softmax_scale, causal=causal)
        return out.reshape(b, s, *out.shape[1:])


class SelfAttention(nn.Module):
    """The eager attention `use_flash_attn=False` selects (mha.py:195-224); materialises (B, H, S, S)."""

    def __init__(self, causal=False, softmax_scale=None, attention_dropout=0.0):
        super().__init__()
        self.causal = causal
        self.softmax_scale = softmax_scale
        self.dropout_p = attention_dropout

    def forward(self, qkv, causal=None, key_padding_mask=None):
        b, s = qkv.shape[:2]
        causal = self.causal if causal is None else causal
        q, k, v = qkv.unbind(dim=2)
        scale = self.softmax_scale or 1.0 / math.sqrt(q.shape[-1])
        scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
        if key_padding_mask is not None:
            pad = torch.full((b, s), -10000.0, dtype=scores.dtype, device=scores.device)
            pad.masked_fill_(key_padding_mask, 0.0)
            scores = scores + pad[:, None, None, :]
        if causal:
            mask = torch.triu(torch.full((s, s), -10000.0, device=scores.device), 1)
            scores = scores + mask.to(dtype=scores.dtype)
        attention = torch.softmax(scores, dim=-1, dtype=v.dtype)
        attention = torch.nn.functional.dropout(attention, self.dropout_p if self.training else 0.0)
        return torch.einsum("bhts,bshd->bthd", attention, v)


class MHA(nn.Module):
    """Multi-head self-attention: Wqkv -> (rotary) -> inner attention -> out_proj.
    State-dict keys: Wqkv.{weight,bias}, out_proj.{weight,bias} (unchanged from the reference)."""

    def __init__(self, embed_dim, num_heads, cross_attn=False, bias=True, dropout=0.0, softmax_scale=None,
                 causal=False, layer_idx=None, dwconv=False, rotary_emb_dim=0, rotary_emb_scale_base=0,
                 fused_bias_fc=False, use_flash_attn=False, return_residual=False, checkpointing=False,
                 device=None, dtype=None) -> None:
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        if cross_attn or dwconv or checkpointing:
            raise RuntimeError("cross_attn / dwconv / checkpointing are out of scope for the Backpack path")
        self.embed_dim = embed_dim
        self.cross_attn = False
        self.causal = causal
        self.layer_idx = layer_idx
        self.dwconv = False
        self.rotary_emb_dim = rotary_emb_dim
        self.use_flash_attn = use_flash_attn
        self.return_residual = return_residual
        self.checkpointing = False
        self.num_heads = num_heads
        if embed_dim % num_heads != 0:
            raise RuntimeError("embed_dim must be divisible by num_heads")
        self.head_dim = embed_dim // num_heads
        if self.rotary_emb_dim > 0:
            self.rotary_emb = RotaryEmbedding(self.rotary_emb_dim, scale_base=rotary_emb_scale_base, device=device)
        linear_cls = nn.Linear if not fused_bias_fc else FusedDense
        if return_residual:
            raise RuntimeError("return_residual (post-norm backward fusion) is out of scope")
        self.Wqkv = linear_cls(embed_dim, 3 * embed_dim, bias=bias, **factory_kwargs)
        inner_attn_cls = FlashSelfAttention if use_flash_attn else SelfAttention
        self.inner_attn = inner_attn_cls(causal=causal, softmax_scale=softmax_scale, attention_dropout=dropout)
        self.out_proj = linear_cls(embed_dim, embed_dim, **factory_kwargs)

    def _update_kv_cache(self, kv, inference_params):
        """kv: (batch, seqlen, 2, nheads, head_dim) of the positions being processed.  Writes them at
        `sequence_len_offset` into this layer's preallocated cache (mha.py:356-380) and returns the cache rows of
        this batch slice (all max_sequence_len positions; the caller knows how many are valid)."""
        if self.layer_idx is None:
            raise RuntimeError("generation requires layer_idx in the constructor")
        cache = inference_params.key_value_memory_dict.get(self.layer_idx)
        if cache is None:
            cache = torch.empty(inference_params.max_batch_size, inference_params.max_sequence_len, 2, self.num_heads,
                                self.head_dim, dtype=kv.dtype, device=kv.device)
            inference_params.key_value_memory_dict[self.layer_idx] = cache
        batch_start = inference_params.batch_size_offset
        batch_end = batch_start + kv.shape[0]
        sequence_start = inference_params.sequence_len_offset
        sequence_end = sequence_start + kv.shape[1]
        if batch_end > cache.shape[0] or sequence_end > cache.shape[1]:
            raise RuntimeError(f"KV cache of shape {tuple(cache.shape)} is too small for batch rows "
                               f"{batch_start}:{batch_end}, positions {sequence_start}:{sequence_end}")
        cache[batch_start:batch_end, sequence_start:sequence_end] = kv
        return cache[batch_start:batch_end]

    def _forward_cached_device_offsets(self, x, inference_params):
        """Decode step whose write position and context lengths are device tensors (`InferenceParams.cache_position`,
        `cache_lengths`): nothing depends on host integers, so the step can be captured in a CUDA graph."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise RuntimeError("a device-offset decode step takes (batch, 1, hidden) input")
        if not self.use_flash_attn or self.rotary_emb_dim > 0:
            raise RuntimeError("device-offset decoding needs use_flash_attn=True and no rotary embedding "
                               "(the rotary tables are sliced on the host)")
        cache = inference_params.key_value_memory_dict.get(self.layer_idx)
        if cache is None:
            raise RuntimeError("device-offset decoding starts after the prompt pass has allocated the KV cache")
        b0 = inference_params.batch_size_offset
        b1 = b0 + x.shape[0]
        if b1 > cache.shape[0]:
            raise RuntimeError(f"KV cache of shape {tuple(cache.shape)} is too small for batch rows {b0}:{b1}")
        qkv = self.Wqkv(x)
        qkv = qkv.reshape(*qkv.shape[:-1], 3, self.num_heads, self.head_dim)
        cache = cache[b0:b1]
        cache.index_copy_(1, inference_params.cache_position, qkv[:, :, 1:])
        scale = self.inner_attn.softmax_scale or 1.0 / math.sqrt(self.head_dim)
        return decode_attention(qkv[:, :, 0], cache, 0, softmax_scale=scale,
                                seqlens_k=inference_params.cache_lengths[b0:b1])

    def _forward_cached(self, x, inference_params):
        """Prompt pass (offset 0): self-attention as usual, K/V stored.  Decode step (one new position, offset > 0):
        the new query against all offset + 1 cached keys, non-causal (mha.py:437-440)."""
        if x.dim() != 3:
            raise RuntimeError("generation needs (batch, seqlen, hidden) input")
        offset = inference_params.sequence_len_offset
        qkv = self.Wqkv(x)
        qkv = qkv.reshape(*qkv.shape[:-1], 3, self.num_heads, self.head_dim)
        if self.rotary_emb_dim > 0:
            qkv = self.rotary_emb(qkv, seqlen_offset=offset)
        cache = self._update_kv_cache(qkv[:, :, 1:], inference_params)
        if offset == 0:
            return self.inner_attn(qkv)
        if qkv.shape[1] != 1:
            # the reference would run these rows non-causally against the cache (mha.py:439), i.e. let each new
            # position see the ones after it; refuse instead of reproducing that
            raise RuntimeError("after the prompt pass, decoding advances one position per call")
        scale = self.inner_attn.softmax_scale or 1.0 / math.sqrt(self.head_dim)
        if self.use_flash_attn:
            return decode_attention(qkv[:, :, 0], cache, offset + 1, softmax_scale=scale)
        k, v = cache[:, :offset + 1].unbind(dim=2)
        scores = torch.einsum("bthd,bshd->bhts", qkv[:, :, 0], k * scale)
        attention = torch.softmax(scores, dim=-1, dtype=v.dtype)
        return torch.einsum("bhts,bshd->bthd", attention, v)

    def forward(self, x, x_kv=None, key_padding_mask=None, cu_seqlens=None, max_seqlen=None,
                inference_params=None, residual_out=None, **kwargs):
        """x: (batch, seqlen, hidden) or, with cu_seqlens / max_seqlen, (total, hidden)."""
        if x_kv is not None:
            raise RuntimeError("cross-attention is out of scope for this path")
        if cu_seqlens is not None:
            if max_seqlen is None or key_padding_mask is not None or not self.use_flash_attn:
                raise RuntimeError("cu_seqlens needs max_seqlen, use_flash_attn=True and no key_padding_mask")
            if self.rotary_emb_dim != 0:
                raise RuntimeError("rotary embedding is not supported with unpadded input")
        if key_padding_mask is not None and self.use_flash_attn:
            raise RuntimeError("key_padding_mask is only supported by the eager SelfAttention")
        if inference_params is not None:
            if key_padding_mask is not None or cu_seqlens is not None or max_seqlen is not None:
                raise RuntimeError("generation takes dense, unmasked batches (mha.py:409-412)")
            if inference_params.cache_position is not None:
                context = self._forward_cached_device_offsets(x, inference_params)
            else:
                context = self._forward_cached(x, inference_params)
        else:
            kw = ({"cu_seqlens": cu_seqlens, "max_seqlen": max_seqlen, **kwargs} if self.use_flash_attn
                  else {"key_padding_mask": key_padding_mask, **kwargs})
            qkv = self.Wqkv(x)
            qkv = qkv.reshape(*qkv.shape[:-1], 3, self.num_heads, self.head_dim)
            if self.rotary_emb_dim > 0:
                qkv = self.rotary_emb(qkv)
            context = self.inner_attn(qkv, **kw)
        context = context.reshape(*context.shape[:-2], self.embed_dim)
        if residual_out is not None:
            # out_proj with the residual add in its epilogue (Block's fused path); returns the residual stream
            return linear_bias_residual_(context, self.out_proj.weight, self.out_proj.bias, residual_out)
        return self.out_proj(context)
