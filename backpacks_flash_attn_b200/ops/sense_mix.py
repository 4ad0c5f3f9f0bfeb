"""Fused Backpack sense-mix operator.

    out[b, i, :] = sum_l sum_{j<=i} softmax_j(scale * q_li . k_lj) * content[b, l, j, :]

replaces `torch.sum(contextualization @ content, dim=1)` together with the softmax half of
`ContextSelfAttn.forward` (training/src/models/backpack.py:116-122, 313).  alpha (b, nv, s, s) is never
materialised.  `content` may be ANY (b, nv, s, d) tensor with unit last stride -- in particular the
transposed view the reference's content model returns (backpack.py:276) and the edited sense tensors
the intervention wrappers build (training/src/models/intervened_models.py:78-101).

Training: `sense_mix` is differentiable.  The forward is the fused kernel (alpha is neither materialised nor kept for
the backward -- 2.15 GB at Backpack-Small, batch 64).  The backward recomputes the scores and is hand-derived: five
batched library GEMMs per sense with ONE own element-wise pass between them (`bp_sense_softmax_bwd`: causal softmax +
softmax backward + scale, in place); sequence lengths that pass does not take (not a multiple of 8, above 8192) fall
back to autograd through the reference's eager composition (backpack.py:116-122, 313).  A fully fused tcgen05
sense-mix backward (no (b, nv, s, s) tensor at all) is not built.

`sense_mix_table` is the inference form: the sense vectors are gathered inside the kernel from a precomputed
(vocab, nv, d) table by token id (C_l(x) is context-free, backpack.py:258), so no (b, s, nv, d) tensor exists.
"""
from __future__ import annotations

import torch

from .. import _lib


def _prepare_qk(qk: torch.Tensor, softmax_scale):
    if qk.dim() != 5 or qk.shape[2] != 2:
        raise RuntimeError("qk must be (batch, seqlen, 2, nv, dk)")
    if qk.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("sense_mix needs fp16/bf16 qk and content of the same dtype")
    dk = qk.shape[-1]
    if softmax_scale is None:
        softmax_scale = dk ** -0.5
    if dk % 8 != 0:
        # TMA needs 16-byte row strides.  Zero-padding the sense key width (e.g. d/nv = 12 at k = 64 senses of a
        # 768-wide model) leaves every q.k dot product unchanged; the softmax scale keeps using the true width.
        qk = torch.nn.functional.pad(qk, (0, (-dk) % 8))
    if not qk.is_contiguous():
        qk = qk.contiguous()
    return qk, float(softmax_scale)


def _out_buffer(b, s, d, dtype, device, out_fp32: bool):
    """`out_fp32` is the test mode of SURVEY.md §8c (T2): the accumulator is stored before the 16-bit rounding."""
    return torch.empty((b, s, d), dtype=torch.float32 if out_fp32 else dtype, device=device)


class _F32Out:
    """Context manager around the library's debug switch (not part of the C ABI)."""

    def __init__(self, hook: str, on: bool):
        self.fn = getattr(_lib.load(), hook) if on else None

    def __enter__(self):
        if self.fn is not None:
            self.fn(1)

    def __exit__(self, *exc):
        if self.fn is not None:
            self.fn(0)
        return False


def _sense_mix_eager(qk, content, scale):
    """The reference's composition (backpack.py:116-122, 313) in the tensors' own dtype."""
    s = qk.shape[1]
    q, k = qk.unbind(dim=2)
    scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
    mask = torch.triu(torch.full((s, s), -10000.0, device=scores.device), 1)
    alpha = torch.softmax(scores + mask.to(scores.dtype), dim=-1, dtype=q.dtype)
    return torch.sum(alpha @ content, dim=1)


def _empty_like_layout(t: torch.Tensor) -> torch.Tensor:
    """An uninitialised tensor with t's strides when t is a permutation of a dense tensor (the transposed view the content
    model returns), else a plain contiguous one (expanded / overlapping inputs must not alias in the gradient)."""
    order = sorted(range(t.dim()), key=lambda i: -t.stride(i))
    expect, dense = 1, True
    for i in reversed(order):
        if t.shape[i] != 1 and t.stride(i) != expect:
            dense = False
            break
        expect *= t.shape[i]
    if dense:
        return torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device)
    return torch.empty(t.shape, dtype=t.dtype, device=t.device)


def _sense_mix_backward_eager(qk, content, dout, scale, want_dqk, want_dcontent, chunk_bytes=1 << 30):
    """Gradients by autograd through the eager composition, a few batch elements at a time (any seqlen)."""
    b, s, _, nv, _ = qk.shape
    step = max(1, chunk_bytes // (4 * nv * s * s * qk.element_size()))
    dqk = torch.empty_like(qk) if want_dqk else None
    dcontent = _empty_like_layout(content) if want_dcontent else None
    for i in range(0, b, step):
        with torch.enable_grad():
            q_ = qk[i:i + step].detach().requires_grad_(want_dqk)
            c_ = content[i:i + step].detach().requires_grad_(want_dcontent)
            o = _sense_mix_eager(q_, c_, scale)
            wanted = [t for t, w in ((q_, want_dqk), (c_, want_dcontent)) if w]
            grads = list(torch.autograd.grad(o, wanted, dout[i:i + step]))
        if want_dqk:
            dqk[i:i + step] = grads.pop(0)
        if want_dcontent:
            dcontent[i:i + step] = grads.pop(0)
    return dqk, dcontent


def _sense_mix_backward(qk, content, dout, scale, want_dqk, want_dcontent, chunk_bytes=4 << 30):
    """Hand-derived backward: per sense l, with P_l = softmax(scale q_l k_l^T + causal mask),

        dA_l = dO C_l^T      dC_l = P_l^T dO      dS_l = scale P_l o (dA_l - rowsum(P_l o dA_l))
        dq_l = dS_l k_l      dk_l = dS_l^T q_l

    The five products are batched library GEMMs over the batch dimension, reading q / k / content and writing dqk /
    dcontent through their strides (no permuted copies); everything between them -- mask, softmax, softmax backward,
    scale -- is ONE pass of `bp_sense_softmax_bwd` over the two (nv, batch, s, s) score tensors, in place.  The
    autograd chain of the eager composition (what the reference trains through, backpack.py:116-122, 313) makes seven
    passes over such tensors and three dense (b, nv, s, s, d) products; this makes two and two."""
    b, s, _, nv, dk = qk.shape
    dout = dout.contiguous()
    dqk = torch.empty_like(qk) if want_dqk else None
    # same strides as `content`: the reference hands a transposed view of (b, s, nv, d), and a gradient in that layout
    # flows back through the transpose / reshape of the content model as a view instead of a 1.6 GB copy
    dcontent = _empty_like_layout(content) if want_dcontent else None
    lib = _lib.load()
    dt = _lib.dtype_code(qk.dtype)
    step = max(1, min(b, chunk_bytes // (2 * nv * s * s * qk.element_size())))
    scores = torch.empty((nv, step, s, s), dtype=qk.dtype, device=qk.device)
    dalpha = torch.empty_like(scores)
    with torch.cuda.device(qk.device):
        stream = _lib.stream_ptr(qk.device)
        for i in range(0, b, step):
            nb = min(step, b - i)
            q, k = qk[i:i + nb, :, 0], qk[i:i + nb, :, 1]            # (nb, s, nv, dk) views
            c, do = content[i:i + nb], dout[i:i + nb]
            S = scores.view(-1)[:nv * nb * s * s].view(nv, nb, s, s)      # contiguous also for a short last chunk
            dA = dalpha.view(-1)[:nv * nb * s * s].view(nv, nb, s, s)
            for l in range(nv):
                torch.bmm(q[:, :, l], k[:, :, l].transpose(1, 2), out=S[l])
                torch.bmm(do, c[:, l].transpose(1, 2), out=dA[l])
            _lib.check(lib.bp_sense_softmax_bwd(S.data_ptr(), dA.data_ptr(), nv * nb * s, s, scale, dt, stream),
                       "bp_sense_softmax_bwd")
            for l in range(nv):
                if want_dcontent:
                    torch.bmm(S[l].transpose(1, 2), do, out=dcontent[i:i + nb, l])
                if want_dqk:
                    torch.bmm(dA[l], k[:, :, l], out=dqk[i:i + nb, :, 0, l])
                    torch.bmm(dA[l].transpose(1, 2), q[:, :, l], out=dqk[i:i + nb, :, 1, l])
    return dqk, dcontent


class _SenseMixFn(torch.autograd.Function):
    """Fused forward (alpha is neither materialised nor saved); backward by recomputation from q, k and content."""

    @staticmethod
    def forward(ctx, qk, content, scale):
        ctx.save_for_backward(qk, content)
        ctx.scale = scale
        with torch.no_grad():
            return sense_mix(qk, content, softmax_scale=scale)

    @staticmethod
    def backward(ctx, dout):
        qk, content = ctx.saved_tensors
        s = qk.shape[1]
        fn = _sense_mix_backward if (s % 8 == 0 and s <= 8192 and content.stride(3) == 1) else _sense_mix_backward_eager
        dqk, dcontent = fn(qk, content, dout, ctx.scale, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dqk, dcontent, None


def sense_mix(qk: torch.Tensor, content: torch.Tensor, softmax_scale: float | None = None,
              return_lse: bool = False, out_fp32: bool = False):
    """qk: (batch, seqlen, 2, nv, dk) -- ContextSelfAttn.Wqkv output viewed as at backpack.py:112-116;
    content: (batch, nv, seqlen, d).  Returns (batch, seqlen, d) [and lse (batch, nv, seqlen) fp32]."""
    _lib.require_cuda(qk, content)
    if torch.is_grad_enabled() and (qk.requires_grad or content.requires_grad):
        if return_lse or out_fp32:
            raise RuntimeError("return_lse / out_fp32 are inference-only options")
        scale = float(softmax_scale) if softmax_scale is not None else qk.shape[-1] ** -0.5
        return _SenseMixFn.apply(qk, content, scale)
    qk, scale = _prepare_qk(qk, softmax_scale)
    if content.dtype != qk.dtype:
        raise RuntimeError("sense_mix needs fp16/bf16 qk and content of the same dtype")
    b, s, _, nv, dk = qk.shape
    if content.dim() != 4 or content.shape[:3] != (b, nv, s):
        raise RuntimeError(f"content must be (batch, nv, seqlen, d) = ({b}, {nv}, {s}, d), got {tuple(content.shape)}")
    d = content.shape[3]
    if content.stride(3) != 1 or any(st % 8 for st in content.stride()[:3]) or content.data_ptr() % 16:
        content = content.contiguous()
    lse = torch.empty((b, nv, s), dtype=torch.float32, device=qk.device)
    out = _out_buffer(b, s, d, qk.dtype, qk.device, out_fp32)
    lib = _lib.load()
    dt = _lib.dtype_code(qk.dtype)
    with torch.cuda.device(qk.device), _F32Out("bp_debug_set_sense_out_f32", out_fp32):
        stream = _lib.stream_ptr(qk.device)
        _lib.check(lib.bp_sense_lse_fwd(qk.data_ptr(), lse.data_ptr(), b, s, nv, dk, scale, dt, stream),
                   "bp_sense_lse_fwd")
        _lib.check(lib.bp_sense_mix_fwd(qk.data_ptr(), content.data_ptr(), lse.data_ptr(), out.data_ptr(),
                                        b, s, nv, dk, d, content.stride(0), content.stride(1), content.stride(2),
                                        scale, dt, stream), "bp_sense_mix_fwd")
    return (out, lse) if return_lse else out


def sense_mix_table(qk: torch.Tensor, table: torch.Tensor, input_ids: torch.Tensor,
                    softmax_scale: float | None = None, return_lse: bool = False, out_fp32: bool = False):
    """qk: (batch, seqlen, 2, nv, dk); table: (vocab, nv, d) sense vectors of every vocabulary item
    (`BackpackModel.build_sense_table()`); input_ids: (batch, seqlen) int64.  Equivalent to
    `sense_mix(qk, table[input_ids].transpose(1, 2))` without ever forming that tensor."""
    _lib.require_cuda(qk, table, input_ids)
    qk, scale = _prepare_qk(qk, softmax_scale)
    b, s, _, nv, dk = qk.shape
    if table.dtype != qk.dtype:
        raise RuntimeError("sense_mix_table needs fp16/bf16 qk and table of the same dtype")
    if table.dim() != 3 or table.shape[1] != nv or not table.is_contiguous():
        raise RuntimeError(f"table must be a contiguous (vocab, nv, d) = (vocab, {nv}, d) tensor, got {tuple(table.shape)}")
    if input_ids.shape != (b, s) or input_ids.dtype != torch.int64:
        raise RuntimeError(f"input_ids must be int64 of shape ({b}, {s})")
    if not input_ids.is_contiguous():
        input_ids = input_ids.contiguous()
    vocab, _, d = table.shape
    lse = torch.empty((b, nv, s), dtype=torch.float32, device=qk.device)
    out = _out_buffer(b, s, d, qk.dtype, qk.device, out_fp32)
    lib = _lib.load()
    dt = _lib.dtype_code(qk.dtype)
    with torch.cuda.device(qk.device), _F32Out("bp_debug_set_sense_out_f32", out_fp32):
        stream = _lib.stream_ptr(qk.device)
        _lib.check(lib.bp_sense_lse_fwd(qk.data_ptr(), lse.data_ptr(), b, s, nv, dk, scale, dt, stream),
                   "bp_sense_lse_fwd")
        _lib.check(lib.bp_sense_mix_table_fwd(qk.data_ptr(), table.data_ptr(), input_ids.data_ptr(), lse.data_ptr(),
                                              out.data_ptr(), b, s, nv, dk, d, vocab, scale, dt, stream),
                   "bp_sense_mix_table_fwd")
    return (out, lse) if return_lse else out
