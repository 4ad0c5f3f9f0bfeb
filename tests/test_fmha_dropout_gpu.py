"""GPU parity of the attention-dropout variants (bp_fmha_fwd_dropout / bp_fmha_bwd_dropout).

The reference's test recovers the mask its kernel drew from the returned S_dmask and feeds it to the fp32 reference
(tests/test_flash_attn.py:383-400, 129-178 `dropout_mask`).  Here the mask is a documented counter-based function of
the seed (csrc/bp_common.cuh), restated in Python by `attention_dropout_mask`; the oracle applies that mask:
    out = ((D o softmax(S)) / (1 - p)) V
and the same criteria as without dropout apply: forward <= 2 x, backward <= 4 x the same-precision PyTorch error
(tests/test_flash_attn.py:404-420), plus the dropout-fraction check of :411-414."""
import pytest
import torch

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from backpacks_flash_attn_b200 import flash_attn_interface as F
    return F


def _ref(qkv, mask, p_eff, causal, dtype=None):
    """softmax attention with a given keep mask; fp32 (-inf masking) when dtype is None, else the reference's eager
    composition in `dtype` (additive -10000 mask, softmax in the activation dtype, mha.py:195-224)."""
    q, k, v = (qkv if dtype is None else qkv).unbind(2)
    if dtype is None:
        q, k, v = q.float(), k.float(), v.float()
    d = q.shape[-1]
    s = q.shape[1]
    if dtype is not None:
        scores = torch.einsum("bthd,bshd->bhts", q, k * d ** -0.5)
    else:
        scores = torch.einsum("bthd,bshd->bhts", q, k) * d ** -0.5
    if causal:
        if dtype is None:
            scores = scores.masked_fill(~torch.ones(s, s, dtype=torch.bool, device=q.device).tril(), float("-inf"))
        else:
            scores = scores + torch.full((s, s), -10000.0, device=q.device).triu(1).to(scores.dtype)
    probs = torch.softmax(scores, dim=-1, dtype=v.dtype)
    probs = probs * mask.to(probs.dtype) / (1.0 - p_eff)
    return torch.einsum("bhts,bshd->bthd", probs, v)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("d", [64, 128, 40])
@pytest.mark.parametrize("s", [97, 128, 257, 512, 1024])
@pytest.mark.parametrize("p", [0.17, 0.1])
def test_fmha_dropout_forward_and_backward_match_the_oracle(p, s, d, causal, dtype):
    F = _ops()
    b, h = 2, (3 if d <= 64 else 2)
    g0 = torch.Generator(device="cuda").manual_seed(s * 7 + d)
    qkv = torch.randn(b, s, 3, h, d, device="cuda", generator=g0).to(dtype)
    g = torch.randn(b, s, h, d, device="cuda", generator=g0).to(dtype)
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    torch.manual_seed(1000 + s + d)
    x = qkv.reshape(b * s, 3, h, d).clone().requires_grad_(True)
    out = F.flash_attn_unpadded_qkvpacked_func(x, cu, s, p, causal=causal)
    dqkv, = torch.autograd.grad(out, x, g.reshape(b * s, h, d))
    torch.manual_seed(1000 + s + d)
    seed = F._new_seed()                                   # the seed the call above drew
    p_eff = F.effective_dropout_p(p)
    mask = F.attention_dropout_mask(seed, b, h, s, s, p, device="cuda")
    vis = torch.ones(s, s, dtype=torch.bool, device="cuda").tril() if causal else torch.ones(s, s, dtype=torch.bool, device="cuda")
    frac = 1.0 - mask[:, :, vis].float().mean().item()
    assert abs(frac - p_eff) < 0.01 + 2.0 / s, (frac, p_eff)                        # tests/test_flash_attn.py:411-414
    xr = qkv.float().requires_grad_(True)
    ref = _ref(xr, mask, p_eff, causal)
    dref, = torch.autograd.grad(ref, xr, g.float())
    xp = qkv.clone().requires_grad_(True)
    pt = _ref(xp, mask, p_eff, causal, dtype=dtype)
    dpt, = torch.autograd.grad(pt, xp, g)
    out = out.reshape(b, s, h, d)
    assert O.max_abs(out, ref) <= 2 * O.max_abs(pt, ref) + 1e-5
    assert O.mean_abs(out, ref) <= 2 * O.mean_abs(pt, ref) + 1e-6
    dqkv = dqkv.reshape(b, s, 3, h, d).float()
    for i, name in enumerate(("dQ", "dK", "dV")):
        err, err_pt = O.max_abs(dqkv[:, :, i], dref[:, :, i]), O.max_abs(dpt[:, :, i].float(), dref[:, :, i])
        assert err <= 4 * err_pt + 1e-5, f"{name}: {err:.3e} vs eager {err_pt:.3e}"
        assert O.mean_abs(dqkv[:, :, i], dref[:, :, i]) <= 2 * O.mean_abs(dpt[:, :, i].float(), dref[:, :, i]) + 1e-6


def test_fmha_dropout_seed_semantics():
    """Same seed -> same result (forward and backward, bit-wise); another seed -> another mask; p = 0 -> the plain
    kernels; the mask does not depend on how the batch is packed."""
    F = _ops()
    b, s, h, d = 3, 300, 4, 64
    qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
    g = torch.randn(b * s, h, d, device="cuda").bfloat16()
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")

    def run(seed_value, p):
        torch.manual_seed(seed_value)
        x = qkv.clone().requires_grad_(True)
        o = F.flash_attn_unpadded_qkvpacked_func(x, cu, s, p, causal=True)
        dx, = torch.autograd.grad(o, x, g)
        return o.detach(), dx

    o1, d1 = run(3, 0.1)
    o2, d2 = run(3, 0.1)
    o3, d3 = run(4, 0.1)
    assert torch.equal(o1, o2) and torch.equal(d1, d2)
    assert not torch.equal(o1, o3)
    o0, d0 = run(3, 0.0)
    with torch.no_grad():
        plain = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
    assert torch.equal(o0, plain)
    # inference-mode call with dropout draws a mask too (module in train mode under no_grad)
    torch.manual_seed(3)
    with torch.no_grad():
        o_ng = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.1, causal=True)
    assert torch.equal(o_ng, o1)


def test_fmha_dropout_varlen_and_cross_lengths():
    F = _ops()
    torch.manual_seed(0)
    lens_q, lens_k = [5, 130, 64, 257], [70, 130, 300, 31]
    h, d, p = 2, 64, 0.25
    q = torch.randn(sum(lens_q), h, d, device="cuda").bfloat16().requires_grad_(True)
    kv = torch.randn(sum(lens_k), 2, h, d, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(sum(lens_q), h, d, device="cuda").bfloat16()
    cq = torch.tensor([0] + list(torch.tensor(lens_q).cumsum(0)), dtype=torch.int32, device="cuda")
    ck = torch.tensor([0] + list(torch.tensor(lens_k).cumsum(0)), dtype=torch.int32, device="cuda")
    torch.manual_seed(77)
    out = F.flash_attn_unpadded_kvpacked_func(q, kv, cq, ck, max(lens_q), max(lens_k), p, causal=False)
    dq, dkv = torch.autograd.grad(out, (q, kv), g)
    torch.manual_seed(77)
    seed = F._new_seed()
    p_eff = F.effective_dropout_p(p)
    mask = F.attention_dropout_mask(seed, len(lens_q), h, max(lens_q), max(lens_k), p, device="cuda")
    sq = sk = 0
    for i, (nq, nk) in enumerate(zip(lens_q, lens_k)):
        qf = q.detach()[sq:sq + nq].float().requires_grad_(True)
        kf = kv.detach()[sk:sk + nk, 0].float().requires_grad_(True)
        vf = kv.detach()[sk:sk + nk, 1].float().requires_grad_(True)
        probs = torch.softmax(torch.einsum("thd,shd->hts", qf, kf) * d ** -0.5, -1) * mask[i, :, :nq, :nk] / (1 - p_eff)
        ref = torch.einsum("hts,shd->thd", probs, vf)
        rq, rk, rv = torch.autograd.grad(ref, (qf, kf, vf), g[sq:sq + nq].float())
        assert O.max_abs(out[sq:sq + nq], ref) < 3e-2
        assert O.max_abs(dq[sq:sq + nq], rq) < 5e-2
        assert O.max_abs(dkv[sk:sk + nk, 0], rk) < 5e-2 and O.max_abs(dkv[sk:sk + nk, 1], rv) < 5e-2
        sq, sk = sq + nq, sk + nk


def test_mha_trains_with_attention_and_residual_dropout():
    """The reference's training configuration (attn_pdrop = resid_pdrop = embd_pdrop = 0.1) runs on the fused path."""
    import torch.nn.functional as Fn
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
    from backpacks_flash_attn_b200.utils.weights import name_seeded_
    cfg = flash_config(n_embd=128, n_head=2, n_layer=2, n_positions=256, num_content_vectors=4)
    assert cfg.attn_pdrop == 0.1 and cfg.resid_pdrop == 0.1
    model = name_seeded_(BackpackLMHeadModel(cfg)).to("cuda", torch.bfloat16).train()
    ids = torch.randint(0, 50257, (2, 200), device="cuda")
    torch.manual_seed(1)
    logits = model(ids).logits
    loss = Fn.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), ids[:, 1:].reshape(-1))
    loss.backward()
    assert torch.isfinite(loss)
    for n, prm in model.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad.float()).all(), n
    torch.manual_seed(1)
    again = model(ids).logits
    assert torch.equal(logits, again)                     # the dropout masks follow torch.manual_seed
    model.eval()
    with torch.no_grad():
        e1, e2 = model(ids).logits, model(ids).logits
    assert torch.equal(e1, e2)
