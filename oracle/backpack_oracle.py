"""CPU oracle for the Backpack forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, functional (no nn.Module) PyTorch restatement of the
reference's pure-PyTorch path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``backpacks_flash_attn_b200`` never does (it fails loudly without its CUDA
library instead of falling back to this code).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imported the real reference
from ``/root/reference`` (with the three import shims of SURVEY.md §8c) in the build
container and wrote the fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function below against them (fp32, atol 2e-6) and against the literal
values quoted in SURVEY.md §8c / Appendix B.

All citations are ``path:line`` relative to the reference repository root.
Weights are passed as a flat ``dict`` using the reference's own state-dict key names
(SURVEY.md §8c), so a reference checkpoint can be fed to the oracle unchanged.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
MASK_VALUE = -10000.0  # additive causal mask used by the eager path (mha.py:218, backpack.py:119)


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """The subset of GPT2Config / BackpackConfig (backpack.py:146-154) the forward reads."""
    n_embd: int = 768
    n_head: int = 12
    n_layer: int = 12
    n_positions: int = 1024
    vocab_size: int = 50257
    num_content_vectors: int = 16
    pad_vocab_size_multiple: int = 8
    layer_norm_epsilon: float = 1e-5
    scale_attn_weights: bool = True
    scale_attn_by_inverse_layer_idx: bool = True
    n_inner: Optional[int] = None
    shrink_final_inner: bool = False

    @property
    def padded_vocab(self) -> int:
        # backpack.py:285-288 and gpt.py:182-185 round the vocabulary up in place.
        m = self.pad_vocab_size_multiple
        return self.vocab_size + (-self.vocab_size) % m

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head

    def mha_softmax_scale(self, layer_idx: int) -> float:
        # gpt.py:46-50 -- 1/sqrt(dh), further divided by (layer_idx+1).
        s = self.head_dim ** -0.5 if self.scale_attn_weights else 1.0
        if self.scale_attn_by_inverse_layer_idx:
            s /= float(layer_idx + 1)
        return s


MICRO = dict(n_embd=384, n_head=6, n_layer=6, n_positions=512)      # gpt2-micro.yaml:4-6
SMALL = dict(n_embd=768, n_head=12, n_layer=12, n_positions=1024)   # gpt2-small.yaml:4-6


# --------------------------------------------------------------------------------------
# deterministic, RNG-order independent weights (SURVEY.md §8c recipe)
# --------------------------------------------------------------------------------------
def canonical_param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """Unique parameter names (tied embeddings listed once) and shapes of the reference model."""
    d, nv, V = cfg.n_embd, cfg.num_content_vectors, cfg.padded_vocab
    inner = cfg.n_inner or 4 * d
    final_inner = d if cfg.shrink_final_inner else inner
    g = "transformer.gpt2_model."
    c = "transformer.content_model."
    shapes = {
        g + "embeddings.word_embeddings.weight": (V, d),
        g + "embeddings.position_embeddings.weight": (cfg.n_positions, d),
        g + "ln_0.weight": (d,), g + "ln_0.bias": (d,),
    }
    for i in range(cfg.n_layer):
        p = f"{g}layers.{i}."
        shapes.update({
            p + "mixer.Wqkv.weight": (3 * d, d), p + "mixer.Wqkv.bias": (3 * d,),
            p + "mixer.out_proj.weight": (d, d), p + "mixer.out_proj.bias": (d,),
            p + "norm1.weight": (d,), p + "norm1.bias": (d,),
            p + "mlp.fc1.weight": (inner, d), p + "mlp.fc1.bias": (inner,),
            p + "mlp.fc2.weight": (d, inner), p + "mlp.fc2.bias": (d,),
            p + "norm2.weight": (d,), p + "norm2.bias": (d,),
        })
    shapes.update({
        c + "ln_0.weight": (d,), c + "ln_0.bias": (d,),
        c + "layers.0.norm1.weight": (d,), c + "layers.0.norm1.bias": (d,),
        c + "layers.0.mlp.fc1.weight": (final_inner, d), c + "layers.0.mlp.fc1.bias": (final_inner,),
        c + "layers.0.mlp.fc2.weight": (d, final_inner), c + "layers.0.mlp.fc2.bias": (d,),
        c + "layers.0.norm2.weight": (d,), c + "layers.0.norm2.bias": (d,),
        c + "final_mlp.fc1.weight": (final_inner, d), c + "final_mlp.fc1.bias": (final_inner,),
        c + "final_mlp.fc2.weight": (nv * d, final_inner), c + "final_mlp.fc2.bias": (nv * d,),
        "transformer.contextualization_attn.Wqkv.weight": (2 * d, d),
        "transformer.contextualization_attn.Wqkv.bias": (2 * d,),
    })
    return shapes


def name_seeded_tensor(name: str, shape: tuple) -> Tensor:
    """One parameter of the golden-vector recipe: seed = crc32(name), fan-in scaled matrices."""
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    w = torch.randn(shape, generator=g)
    if len(shape) == 2:
        return w * shape[1] ** -0.5
    if name.endswith(("ln_0.weight", "norm1.weight", "norm2.weight")):
        return 1.0 + 0.1 * w
    return 0.02 * w


def name_seeded_weights(cfg: OracleConfig) -> Dict[str, Tensor]:
    return {k: name_seeded_tensor(k, s) for k, s in canonical_param_shapes(cfg).items()}


# --------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------
def attention_fp32_ref(q: Tensor, k: Tensor, v: Tensor, softmax_scale: Optional[float] = None,
                       causal: bool = False) -> tuple[Tensor, Tensor]:
    """Exact softmax attention in fp32 with -inf masking: the judge the reference's own
    fmha tests use (tests/test_flash_attn.py:129-178, upcast=True, no dropout/padding).

    q, k, v: (b, s, h, d).  Returns (out (b, s, h, d) fp32, lse (b, h, s) fp32) where
    lse = log sum_j exp(scale * q.k_j) as written by mha_fwd (fmha_api.cpp:276).
    """
    q, k, v = q.float(), k.float(), v.float()
    d = q.shape[-1]
    scale = softmax_scale if softmax_scale is not None else d ** -0.5
    scores = torch.einsum("bthd,bshd->bhts", q, k) * scale
    if causal:
        sq, sk = scores.shape[-2:]
        keep = torch.ones(sq, sk, dtype=torch.bool, device=scores.device).tril()  # top-left aligned, col <= row (mask.h:70)
        scores = scores.masked_fill(~keep, float("-inf"))
    lse = torch.logsumexp(scores, dim=-1)
    p = torch.softmax(scores, dim=-1)
    return torch.einsum("bhts,bshd->bthd", p, v), lse


def self_attention_eager(qkv: Tensor, softmax_scale: Optional[float] = None,
                         causal: bool = True) -> Tensor:
    """The reference's non-flash attention (flash_attn/modules/mha.py:195-224), in the
    tensor's own dtype: K pre-multiplied by the scale, additive -10000 causal mask,
    softmax in the activation dtype.  qkv: (b, s, 3, h, d) -> (b, s, h, d).
    """
    q, k, v = qkv.unbind(dim=2)
    scale = softmax_scale or 1.0 / math.sqrt(q.shape[-1])
    scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
    if causal:
        s = qkv.shape[1]
        mask = torch.full((s, s), MASK_VALUE, device=scores.device).triu(1)
        scores = scores + mask.to(scores.dtype)
    probs = torch.softmax(scores, dim=-1, dtype=v.dtype)
    return torch.einsum("bhts,bshd->bthd", probs, v)


def context_weights_eager(hidden: Tensor, wqk: Tensor, bqk: Tensor, nv: int) -> Tensor:
    """ContextSelfAttn.forward (training/src/models/backpack.py:107-122): project to 2d,
    split into (q, k) x nv senses of width d/nv, causal softmax -> alpha (b, nv, s, s)."""
    b, s, d = hidden.shape
    qk = F.linear(hidden, wqk, bqk).reshape(b, s, 2, nv, d // nv)
    q, k = qk.unbind(dim=2)
    scale = 1.0 / math.sqrt(d // nv)
    scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
    scores = scores + torch.full((s, s), MASK_VALUE, device=scores.device).triu(1).to(scores.dtype)
    return torch.softmax(scores, dim=-1, dtype=q.dtype)


def sense_sum(alpha: Tensor, content: Tensor) -> Tensor:
    """o_i = sum_l sum_j alpha[l,i,j] C_l(x_j)  (backpack.py:313).
    alpha (b, nv, s, s), content (b, nv, s, d) -> (b, s, d)."""
    return torch.sum(alpha @ content, dim=1)


def sense_mix_eager(qk: Tensor, content: Tensor) -> Tensor:
    """The reference's own composition from the projected (q, k) on (backpack.py:116-122 + :313), in the
    tensors' dtype: k pre-scaled, additive -10000 mask, softmax in that dtype, alpha @ content, sum over
    senses.  qk (b, s, 2, nv, dk); content (b, nv, s, d)."""
    q, k = qk.unbind(dim=2)
    s = qk.shape[1]
    scale = 1.0 / math.sqrt(q.shape[-1])
    scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
    scores = scores + torch.full((s, s), MASK_VALUE, device=scores.device).triu(1).to(scores.dtype)
    alpha = torch.softmax(scores, dim=-1, dtype=q.dtype)
    return torch.sum(alpha @ content, dim=1)


def sense_mix_fp32_ref(qk: Tensor, content: Tensor, softmax_scale: Optional[float] = None
                       ) -> tuple[Tensor, Tensor]:
    """fp32 judge for the fused sense-mix operator: exact causal softmax (-inf mask, scale
    applied to the fp32 scores) followed by the sense sum.

    qk: (b, s, 2, nv, dk); content: (b, nv, s, d) (any strides).
    Returns (out (b, s, d) fp32, lse (b, nv, s) fp32)."""
    q, k = qk.float().unbind(dim=2)
    scale = softmax_scale if softmax_scale is not None else q.shape[-1] ** -0.5
    s = q.shape[1]
    scores = torch.einsum("bthd,bshd->bhts", q, k) * scale
    scores = scores.masked_fill(~torch.ones(s, s, dtype=torch.bool, device=scores.device).tril(), float("-inf"))
    lse = torch.logsumexp(scores, dim=-1)
    alpha = torch.softmax(scores, dim=-1)
    return torch.sum(alpha @ content.float(), dim=1), lse


def gelu_tanh(x: Tensor) -> Tensor:
    """gelu_new / approximate='tanh' (gpt.py:87-89; fused epilogue fused_dense_cuda.cu:128)."""
    return F.gelu(x, approximate="tanh")


def mlp(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    """Mlp.forward (flash_attn/modules/mlp.py:26-30) == FusedDenseGeluDense forward
    (flash_attn/ops/fused_dense.py:349-354)."""
    return F.linear(gelu_tanh(F.linear(x, w1, b1)), w2, b2)


def add_layer_norm(x0: Tensor, x1: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float,
                   residual_in_fp32: bool = True, fused: bool = False) -> tuple[Tensor, Tensor]:
    """residual add + LayerNorm, pre-norm bookkeeping.

    fused=False: the un-fused Block path (block.py:75-76, gpt.py:232-234): the fp32 residual
    is rounded to the weight dtype *before* LayerNorm.
    fused=True : dropout_add_ln_fwd semantics (csrc/layer_norm/ln_fwd_kernels.cuh:98-188,
    flash_attn/ops/layer_norm.py:10-24) with p=0: statistics and normalisation from the
    fp32 sum, output rounded once.
    Returns (z in gamma.dtype, residual)."""
    res = x0.float() if x1 is None else x0.float() + x1.float()
    if not residual_in_fp32 and x1 is None:
        res = res.to(x0.dtype)
    d = res.shape[-1]
    if fused:
        z = F.layer_norm(res.float(), (d,), gamma.float(), beta.float(), eps).to(gamma.dtype)
    else:
        z = F.layer_norm(res.to(gamma.dtype), (d,), gamma, beta, eps)
    return z, res


def apply_rotary_ref(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """Non-interleaved (GPT-NeoX) rotary on the first 2*cos.shape[-1] features
    (flash_attn/layers/rotary.py:13-28).  x: (b, s, h, d); cos, sin: (s, rot/2)."""
    ro = cos.shape[-1] * 2
    x1, x2 = x[..., : ro // 2], x[..., ro // 2: ro]
    c, s_ = cos[: x.shape[1], None, :], sin[: x.shape[1], None, :]
    rot = torch.cat([x1 * c - x2 * s_, x1 * s_ + x2 * c], dim=-1)
    return torch.cat([rot, x[..., ro:]], dim=-1)


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------
def _ln(w: Dict[str, Tensor], prefix: str):
    return w[prefix + ".weight"], w[prefix + ".bias"]


def _lin(w: Dict[str, Tensor], prefix: str):
    return w[prefix + ".weight"], w[prefix + ".bias"]


def gpt_trunk(ids: Tensor, w: Dict[str, Tensor], cfg: OracleConfig, fused_ln: bool = False) -> Tensor:
    """GPTModel.forward (flash_attn/models/gpt.py:224-246) with prenorm Blocks
    (flash_attn/modules/block.py:70-106) and eager MHA (mha.py:416-431, 466).
    Returns the last block's norm2 output (b, s, d)."""
    g = "transformer.gpt2_model."
    b, s = ids.shape
    eps = cfg.layer_norm_epsilon
    emb = w[g + "embeddings.word_embeddings.weight"][ids] \
        + w[g + "embeddings.position_embeddings.weight"][torch.arange(s, device=ids.device)]  # embedding.py:27-39
    h, res = add_layer_norm(emb, None, *_ln(w, g + "ln_0"), eps, fused=fused_ln)
    for i in range(cfg.n_layer):
        p = f"{g}layers.{i}."
        qkv = F.linear(h, *_lin(w, p + "mixer.Wqkv")).reshape(b, s, 3, cfg.n_head, cfg.head_dim)
        ctx = self_attention_eager(qkv, cfg.mha_softmax_scale(i), causal=True)
        mix = F.linear(ctx.reshape(b, s, cfg.n_embd), *_lin(w, p + "mixer.out_proj"))
        h, res = add_layer_norm(mix, res, *_ln(w, p + "norm1"), eps, fused=fused_ln)
        m = mlp(h, *_lin(w, p + "mlp.fc1"), *_lin(w, p + "mlp.fc2"))
        h, res = add_layer_norm(m, res, *_ln(w, p + "norm2"), eps, fused=fused_ln)
    return h


def content_vectors(ids: Tensor, w: Dict[str, Tensor], cfg: OracleConfig, fused_ln: bool = False) -> Tensor:
    """BackpackContentModule.forward (backpack.py:251-276): word embedding only (no positions),
    ln_0, one Identity-mixer block, final MLP to nv*d, viewed (b, nv, s, d) (transposed view)."""
    g = "transformer.gpt2_model."
    c = "transformer.content_model."
    b, s = ids.shape
    eps = cfg.layer_norm_epsilon
    emb = w[g + "embeddings.word_embeddings.weight"][ids]
    h, res = add_layer_norm(emb, None, *_ln(w, c + "ln_0"), eps, fused=fused_ln)
    # Identity mixer still passes through the block's residual add (block.py:72-76).
    h, res = add_layer_norm(h, res, *_ln(w, c + "layers.0.norm1"), eps, fused=fused_ln)
    m = mlp(h, *_lin(w, c + "layers.0.mlp.fc1"), *_lin(w, c + "layers.0.mlp.fc2"))
    h, res = add_layer_norm(m, res, *_ln(w, c + "layers.0.norm2"), eps, fused=fused_ln)
    out = mlp(h, *_lin(w, c + "final_mlp.fc1"), *_lin(w, c + "final_mlp.fc2"))
    return out.reshape(b, s, cfg.num_content_vectors, cfg.n_embd).transpose(1, 2)


def backpack_hidden(ids: Tensor, w: Dict[str, Tensor], cfg: OracleConfig, fused_ln: bool = False,
                    return_parts: bool = False):
    """BackpackModel.forward (backpack.py:297-314)."""
    ctx_h = gpt_trunk(ids, w, cfg, fused_ln)
    alpha = context_weights_eager(ctx_h, *_lin(w, "transformer.contextualization_attn.Wqkv"),
                                  cfg.num_content_vectors)
    content = content_vectors(ids, w, cfg, fused_ln)
    hid = sense_sum(alpha, content)
    if return_parts:
        return hid, dict(ctx_h=ctx_h, alpha=alpha, content=content)
    return hid


def backpack_logits(ids: Tensor, w: Dict[str, Tensor], cfg: OracleConfig, fused_ln: bool = False) -> Tensor:
    """BackpackLMHeadModel.forward (backpack.py:342-351); lm_head is tied to wte (:339-340)."""
    hid = backpack_hidden(ids, w, cfg, fused_ln)
    return F.linear(hid, w["transformer.gpt2_model.embeddings.word_embeddings.weight"])


def cast_weights(w: Dict[str, Tensor], dtype: torch.dtype) -> Dict[str, Tensor]:
    return {k: v.to(dtype) for k, v in w.items()}


# --------------------------------------------------------------------------------------
# parity metrics (reference test rule, tests/test_flash_attn.py:426-428)
# --------------------------------------------------------------------------------------
def max_abs(a: Tensor, b: Tensor) -> float:
    return (a.float() - b.float()).abs().max().item()


def mean_abs(a: Tensor, b: Tensor) -> float:
    return (a.float() - b.float()).abs().mean().item()


def bf16_ulp_histogram(ours: Tensor, exact_fp32: Tensor, max_ulp: int = 4) -> Dict[int, int]:
    """Distance, in bf16 ulps of the correctly rounded value, between a bf16 result and the
    rounded fp32 oracle."""
    a = ours.to(torch.bfloat16).view(torch.int16).to(torch.int32)
    b = exact_fp32.to(torch.bfloat16).view(torch.int16).to(torch.int32)
    # map sign-magnitude to a monotone integer line
    a = torch.where(a < 0, -(a & 0x7FFF), a)
    b = torch.where(b < 0, -(b & 0x7FFF), b)
    dist = (a - b).abs().clamp(max=max_ulp)
    return {i: int((dist == i).sum()) for i in range(max_ulp + 1)}
