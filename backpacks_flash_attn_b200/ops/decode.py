"""Incremental-decoding operators: one new position per sequence.

`decode_attention`   the seqlen_q = 1 form of `MHA.forward(..., inference_params)`: the new query against the reference's
                     (max_batch, max_seqlen, 2, nheads, headdim) KV cache (flash_attn/modules/mha.py:356-380, 432-440;
                     decode steps use causal=False because the new token attends to every cached key).
`sense_mix_decode`   the Backpack sense-mix of the last position against a cache of contextualisation keys and the token
                     ids of the context, sense vectors served from the (vocab, nv, d) table.  The reference regenerates
                     the whole prefix for every token instead (training/src/utils/generation.py:34-44, 62-72).

Both are HBM-bound row streams on the CUDA cores (bp_decode.cu); there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib


def _lens_arg(seqlens, batch, device):
    if seqlens is None:
        return None, 0
    if seqlens.dtype != torch.int32 or seqlens.shape != (batch,) or seqlens.device != device:
        raise RuntimeError(f"per-sequence lengths must be an int32 tensor of shape ({batch},) on {device}")
    return seqlens.contiguous(), 0


def decode_attention(q: torch.Tensor, kv_cache: torch.Tensor, seqlen_k: int, softmax_scale: float | None = None,
                     seqlens_k: torch.Tensor | None = None) -> torch.Tensor:
    """q: (batch, 1, nheads, headdim) or (batch, nheads, headdim); kv_cache: (>= batch, max_seqlen, 2, nheads, headdim)
    (a batch/sequence slice of the preallocated cache is fine); the first `seqlen_k` cached positions (or
    `seqlens_k[b]`, int32 on the device) are attended to.  Returns the context in q's shape."""
    _lib.require_cuda(q, kv_cache)
    shape = q.shape
    if q.dim() == 4:
        if q.shape[1] != 1:
            raise RuntimeError("decode_attention takes one query position per sequence")
        q = q[:, 0]
    if q.dim() != 3 or kv_cache.dim() != 5 or kv_cache.shape[2] != 2:
        raise RuntimeError("q must be (batch, [1,] nheads, headdim) and kv_cache (batch, seqlen, 2, nheads, headdim)")
    b, h, dh = q.shape
    if q.dtype not in (torch.float16, torch.bfloat16) or kv_cache.dtype != q.dtype:
        raise RuntimeError("decode_attention needs fp16/bf16 q and kv_cache of the same dtype")
    if kv_cache.shape[0] < b or kv_cache.shape[3:] != (h, dh):
        raise RuntimeError(f"kv_cache {tuple(kv_cache.shape)} does not match q {tuple(shape)}")
    if kv_cache.stride(4) != 1 or kv_cache.stride(3) != dh:
        raise RuntimeError("kv_cache must keep (nheads, headdim) contiguous")
    if seqlens_k is None and not 0 < seqlen_k <= kv_cache.shape[1]:
        raise RuntimeError(f"seqlen_k = {seqlen_k} outside the cache (max_seqlen {kv_cache.shape[1]})")
    if torch.is_grad_enabled() and q.requires_grad:
        raise RuntimeError("backward is not implemented; call under torch.no_grad()/inference_mode()")
    if softmax_scale is None:
        softmax_scale = dh ** -0.5
    q = q.contiguous()
    out = torch.empty_like(q)
    lens, _ = _lens_arg(seqlens_k, b, q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.bp_decode_attn_fwd(q.data_ptr(), kv_cache.data_ptr(), out.data_ptr(),
                                          lens.data_ptr() if lens is not None else None, b, h, dh, int(seqlen_k),
                                          kv_cache.stride(0), kv_cache.stride(1), kv_cache.stride(2),
                                          float(softmax_scale), _lib.dtype_code(q.dtype), _lib.stream_ptr(q.device)),
                   "bp_decode_attn_fwd")
    return out.reshape(shape)


def sense_mix_decode(q: torch.Tensor, k_cache: torch.Tensor, ids_cache: torch.Tensor, table: torch.Tensor,
                     seqlen: int, softmax_scale: float | None = None, seqlens: torch.Tensor | None = None) -> torch.Tensor:
    """q: (batch, nv, dk) contextualisation query of the new position; k_cache: (>= batch, max_seqlen, nv, dk) cached
    contextualisation keys (the new position's key included); ids_cache: (>= batch, max_seqlen) int64 token ids of the
    context; table: (vocab, nv, d).  Returns the Backpack output of the new position, (batch, d).

    `softmax_scale` defaults to dk ** -0.5 of the width PASSED here; callers that zero-pad dk to a multiple of 8 must
    pass the scale of the true width (as `BackpackModel` does)."""
    _lib.require_cuda(q, k_cache, ids_cache, table)
    if q.dim() != 3 or k_cache.dim() != 4 or table.dim() != 3 or ids_cache.dim() != 2:
        raise RuntimeError("expected q (batch, nv, dk), k_cache (batch, seqlen, nv, dk), ids (batch, seqlen), table (vocab, nv, d)")
    b, nv, dk = q.shape
    if q.dtype not in (torch.float16, torch.bfloat16) or k_cache.dtype != q.dtype or table.dtype != q.dtype:
        raise RuntimeError("sense_mix_decode needs fp16/bf16 q, k_cache and table of the same dtype")
    if dk % 8 != 0:
        raise RuntimeError("dk must be a multiple of 8 (zero-pad the sense key width)")
    if k_cache.shape[0] < b or k_cache.shape[2:] != (nv, dk) or k_cache.stride(3) != 1 or k_cache.stride(2) != dk \
            or k_cache.stride(1) != nv * dk:
        raise RuntimeError(f"k_cache {tuple(k_cache.shape)} must be (batch, seqlen, {nv}, {dk}) with contiguous rows")
    if ids_cache.dtype != torch.int64 or ids_cache.shape[0] < b or ids_cache.stride(1) != 1:
        raise RuntimeError("ids_cache must be int64 (batch, seqlen) with unit inner stride")
    if table.shape[1] != nv or not table.is_contiguous():
        raise RuntimeError(f"table must be a contiguous (vocab, {nv}, d) tensor, got {tuple(table.shape)}")
    if seqlens is None and not 0 < seqlen <= min(k_cache.shape[1], ids_cache.shape[1]):
        raise RuntimeError(f"seqlen = {seqlen} outside the caches")
    if softmax_scale is None:
        softmax_scale = dk ** -0.5
    vocab, _, d = table.shape
    q = q.contiguous()
    out = torch.empty((b, d), dtype=q.dtype, device=q.device)
    lens, _ = _lens_arg(seqlens, b, q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.bp_sense_mix_decode_fwd(q.data_ptr(), k_cache.data_ptr(), ids_cache.data_ptr(), table.data_ptr(),
                                               out.data_ptr(), lens.data_ptr() if lens is not None else None, b,
                                               int(seqlen), nv, dk, d, vocab, k_cache.stride(0), ids_cache.stride(0),
                                               float(softmax_scale), _lib.dtype_code(q.dtype),
                                               _lib.stream_ptr(q.device)),
                   "bp_sense_mix_decode_fwd")
    return out
