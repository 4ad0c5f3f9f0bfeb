"""GPU parity for the HBM-bound satellites: residual-add LayerNorm and rotary."""
import pytest
import torch

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cols", [768, 384, 256, 1024, 1536, 2048, 3072, 4096, 8192, 40])
@pytest.mark.parametrize("has_residual", [True, False])
@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32", "bf16w32"])
def test_dropout_add_layer_norm_prenorm(cols, has_residual, dt):
    """Mirrors tests/ops/test_dropout_layer_norm.py:109-165 (eval mode, prenorm): |ours-ref| <= 4|pt-ref|+1e-4."""
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    xdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32, "bf16w32": torch.bfloat16}[dt]
    wdt = torch.float32 if dt in ("fp32", "bf16w32") else xdt
    torch.manual_seed(cols)
    rows = 517
    x0 = torch.randn(rows, cols, device="cuda").to(xdt)
    x1 = torch.randn(rows, cols, device="cuda") if has_residual else None
    gamma = (1 + 0.1 * torch.randn(cols, device="cuda")).to(wdt)
    beta = (0.1 * torch.randn(cols, device="cuda")).to(wdt)
    z, res = dropout_add_layer_norm(x0, x1, gamma, beta, 0.0, 1e-5, prenorm=True, residual_in_fp32=True)
    assert z.dtype == xdt and res.dtype == torch.float32
    res_ref = x0.float() + (x1 if has_residual else 0)
    assert torch.equal(res, res_ref)
    ref = torch.nn.functional.layer_norm(res_ref, (cols,), gamma.float(), beta.float(), 1e-5)
    pt = torch.nn.functional.layer_norm(res_ref.to(xdt), (cols,), gamma.to(xdt), beta.to(xdt), 1e-5)
    assert (z.float() - ref).abs().max() <= 4 * (pt.float() - ref).abs().max() + 1e-4
    z_or, _ = O.add_layer_norm(x0, x1, gamma, beta, 1e-5, fused=True)
    assert (z.float() - z_or.float()).abs().max() <= (2e-2 if xdt != torch.float32 else 1e-5)


def test_layer_norm_config3_shape_and_postnorm():
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    x0 = torch.randn(65536, 768, device="cuda").bfloat16()
    x1 = torch.randn(65536, 768, device="cuda")
    g = torch.ones(768, device="cuda").bfloat16()
    b = torch.zeros(768, device="cuda").bfloat16()
    z, res = dropout_add_layer_norm(x0, x1, g, b, 0.0, 1e-5, prenorm=True)
    ref = torch.nn.functional.layer_norm(x0.float() + x1, (768,))
    assert (z.float() - ref).abs().max() < 4e-2
    z2 = dropout_add_layer_norm(x0, x1, g, b, 0.0, 1e-5, prenorm=False)
    assert torch.equal(z, z2)
    # dropout_p > 0: z = LN(dropout(x0) + x1) -- the residual stream tells which elements of x0 were kept
    zd, rd = dropout_add_layer_norm(x0, x1, g, b, 0.25, 1e-5, prenorm=True)
    kept = ((rd - x1).abs() > 0) | (x0 == 0)
    frac = 1.0 - kept.float().mean().item()
    assert 0.24 < frac < 0.26, frac
    want = torch.where(kept, x0.float() / 0.75, torch.zeros((), device="cuda")) + x1
    assert (rd - want).abs().max() < 2e-2                        # x0 / (1 - p) is rounded to bf16 before the add
    assert (zd.float() - torch.nn.functional.layer_norm(rd, (768,))).abs().max() < 4e-2
    with pytest.raises(RuntimeError, match="rowscale"):
        dropout_add_layer_norm(x0, x1, g, b, 0.0, 1e-5, rowscale=torch.ones(65536, device="cuda"))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rotary_dim", [64, 32, 16, 8, 2])
def test_rotary_qkv_inplace(dtype, rotary_dim):
    """Mirrors tests/test_rotary.py:14-38: CUDA vs apply_rotary_emb_torch on q and k of a packed qkv."""
    from backpacks_flash_attn_b200.layers.rotary import RotaryEmbedding
    torch.manual_seed(0)
    b, s, h, d = 3, 217, 4, 64
    qkv = torch.randn(b, s, 3, h, d, device="cuda").to(dtype)
    rot = RotaryEmbedding(rotary_dim, device="cuda")
    orig = qkv.clone()
    out = rot(qkv)
    assert out.data_ptr() == qkv.data_ptr()                     # in place (rotary.py:98-104)
    cos, sin = rot._cos_cached, rot._sin_cached
    q_ref = O.apply_rotary_ref(orig[:, :, 0].float(), cos.float(), sin.float())
    k_ref = O.apply_rotary_ref(orig[:, :, 1].float(), cos.float(), sin.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert (out[:, :, 0].float() - q_ref).abs().max() < tol
    assert (out[:, :, 1].float() - k_ref).abs().max() < tol
    assert torch.equal(out[:, :, 2], orig[:, :, 2])             # v untouched
    assert torch.equal(out[..., rotary_dim:][:, :, :2], orig[..., rotary_dim:][:, :, :2])


def test_rotary_xpos_scale():
    from backpacks_flash_attn_b200.layers.rotary import RotaryEmbedding
    torch.manual_seed(1)
    qkv = torch.randn(2, 64, 3, 2, 32, device="cuda").bfloat16()
    orig = qkv.clone()
    rot = RotaryEmbedding(32, scale_base=512, device="cuda")
    out = rot(qkv)
    q_ref = O.apply_rotary_ref(orig[:, :, 0].float(), rot._cos_cached.float(), rot._sin_cached.float())
    k_ref = O.apply_rotary_ref(orig[:, :, 1].float(), rot._cos_k_cached.float(), rot._sin_k_cached.float())
    assert (out[:, :, 0].float() - q_ref).abs().max() < 2e-2
    assert (out[:, :, 1].float() - k_ref).abs().max() < 2e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cols", [768, 384, 1024, 8192, 40])
def test_layer_norm_from_fp32_residual(cols, dtype):
    """bp_ln_fwd: LayerNorm of the already-summed fp32 residual stream, 16-bit output (Block's fused path)."""
    from backpacks_flash_attn_b200.ops.layer_norm import layer_norm_from_residual
    torch.manual_seed(cols)
    x = torch.randn(777, cols, device="cuda") * 2 + 0.3
    g = (1 + 0.1 * torch.randn(cols, device="cuda")).to(dtype)
    b = (0.1 * torch.randn(cols, device="cuda")).to(dtype)
    with torch.no_grad():
        z = layer_norm_from_residual(x, g, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x, (cols,), g.float(), b.float(), 1e-5)
    pt = torch.nn.functional.layer_norm(x.to(dtype), (cols,), g, b, 1e-5)
    assert z.dtype == dtype and z.shape == x.shape
    assert (z.float() - ref).abs().max() <= 4 * (pt.float() - ref).abs().max() + 1e-4   # the reference's LN rule
