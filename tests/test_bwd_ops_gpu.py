"""GPU parity of the backward operators around the attention kernel: bp_ln_residual_bwd and
bp_bias_act_bwd behind the reference-shaped autograd operators, against fp32 autograd of the oracle's definitions.

Criteria are the reference's own:
  * LayerNorm (tests/ops/test_dropout_layer_norm.py:101-106): input gradients <= 4 x the same-precision PyTorch error
    + 1e-4, weight / bias gradients <= 2 x + 3e-5;
  * fused dense (tests/ops/test_fused_dense.py:20, 55-59, 110-117): allclose to the same-precision PyTorch module,
    rtol 3e-3, atol 1e-2 (bf16) / 1e-3 (fp16), x10 for weight and x5 for bias gradients.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hidden", [768, 256, 384, 1024, 1536, 2048, 200])
@pytest.mark.parametrize("has_residual", [True, False])
@pytest.mark.parametrize("prenorm", [True, False])
@pytest.mark.parametrize("in_dtype,res_dtype,w_dtype", [
    (torch.bfloat16, torch.float32, torch.bfloat16), (torch.float16, torch.float32, torch.float16),
    (torch.bfloat16, torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16, torch.float16),
    (torch.float32, torch.float32, torch.float32), (torch.bfloat16, torch.float32, torch.float32)])
def test_dropout_add_layer_norm_backward(in_dtype, res_dtype, w_dtype, prenorm, has_residual, hidden):
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    torch.manual_seed(hidden + prenorm)
    rows = (8, 131)
    x0_pt = torch.randn(*rows, hidden, device="cuda", dtype=in_dtype, requires_grad=True)
    x0 = x0_pt.detach().clone().requires_grad_()
    x0_ref = x0_pt.detach().clone().float().requires_grad_()
    if has_residual:
        x1_pt = torch.randn(*rows, hidden, device="cuda", dtype=res_dtype, requires_grad=True)
        x1 = x1_pt.detach().clone().requires_grad_()
        x1_ref = x1_pt.detach().clone().float().requires_grad_()
    else:
        x1 = x1_pt = x1_ref = None
    w_pt = torch.randn(hidden, device="cuda", dtype=w_dtype, requires_grad=True)
    b_pt = torch.randn(hidden, device="cuda", dtype=w_dtype, requires_grad=True)
    w, b = w_pt.detach().clone().requires_grad_(), b_pt.detach().clone().requires_grad_()
    w_ref, b_ref = w_pt.detach().float().requires_grad_(), b_pt.detach().float().requires_grad_()
    residual_in_fp32 = (not has_residual) and res_dtype == torch.float32
    res = dropout_add_layer_norm(x0, x1, w, b, 0.0, 1e-5, prenorm=prenorm, residual_in_fp32=residual_in_fp32)
    out, resid = res if prenorm else (res, None)
    assert out.dtype == in_dtype
    # same-precision PyTorch (as the reference's test builds it) and fp32 reference
    r_pt = (x0_pt.float() + x1_pt.float()).to(res_dtype) if has_residual else x0_pt.float().to(res_dtype)
    r_ref = x0_ref + x1_ref if has_residual else x0_ref
    out_pt = F.layer_norm(r_pt.to(w_dtype), (hidden,), w_pt, b_pt, 1e-5).to(in_dtype)
    out_ref = F.layer_norm(r_ref, (hidden,), w_ref, b_ref, 1e-5)
    assert O.max_abs(out, out_ref) <= 4 * O.max_abs(out_pt, out_ref) + 1e-4
    g = torch.randn_like(out) / rows[0]
    if prenorm:
        assert resid.dtype == (res_dtype if (has_residual or residual_in_fp32) else in_dtype)
        g2 = torch.randn(*rows, hidden, device="cuda") / rows[0]
        (out.float() * g.float()).sum().add((resid.float() * g2).sum()).backward()
        (out_pt.float() * g.float()).sum().add((r_pt.float() * g2).sum()).backward()
        (out_ref * g.float()).sum().add((r_ref * g2).sum()).backward()
    else:
        out.backward(g)
        out_pt.backward(g)
        out_ref.backward(g.float())
    assert x0.grad.dtype == in_dtype
    assert O.max_abs(x0.grad, x0_ref.grad) <= 4 * O.max_abs(x0_pt.grad, x0_ref.grad) + 1e-4
    if has_residual:
        assert x1.grad.dtype == res_dtype
        assert O.max_abs(x1.grad, x1_ref.grad) <= 4 * O.max_abs(x1_pt.grad, x1_ref.grad) + 1e-4
    assert O.max_abs(w.grad, w_ref.grad) <= 2 * O.max_abs(w_pt.grad, w_ref.grad) + 3e-5
    assert O.max_abs(b.grad, b_ref.grad) <= 2 * O.max_abs(b_pt.grad, b_ref.grad) + 3e-5


def test_layer_norm_backward_full_size_and_determinism():
    """Config-3 residual stream (65536 x 768): more rows than resident warps, deterministic column sums."""
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    torch.manual_seed(0)
    x0 = torch.randn(65536, 768, device="cuda").bfloat16().requires_grad_()
    x1 = torch.randn(65536, 768, device="cuda").requires_grad_()
    w = (1 + 0.1 * torch.randn(768, device="cuda")).bfloat16().requires_grad_()
    b = (0.1 * torch.randn(768, device="cuda")).bfloat16().requires_grad_()
    z, r = dropout_add_layer_norm(x0, x1, w, b, 0.0, 1e-5, prenorm=True)
    g, g2 = torch.randn_like(z) / 256, torch.randn_like(r) / 256
    first = torch.autograd.grad((z, r), (x0, x1, w, b), (g, g2), retain_graph=True)
    again = torch.autograd.grad((z, r), (x0, x1, w, b), (g, g2), retain_graph=True)
    for a, c in zip(first, again):
        assert torch.equal(a, c)
    xr0, xr1 = x0.detach().float().requires_grad_(), x1.detach().clone().requires_grad_()
    wr, br = w.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    rr = xr0 + xr1
    zr = F.layer_norm(rr, (768,), wr, br, 1e-5)
    ref = torch.autograd.grad((zr, rr), (xr0, xr1, wr, br), (g.float(), g2))
    assert O.max_abs(first[0], ref[0]) < 2e-3 and O.max_abs(first[1], ref[1]) < 1e-4
    assert O.max_abs(first[2], ref[2]) / ref[2].abs().max().item() < 1e-2
    assert O.max_abs(first[3], ref[3]) / ref[3].abs().max().item() < 1e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("has_bias", [True, False])
@pytest.mark.parametrize("out_features", [1024, 4096, 776])
@pytest.mark.parametrize("in_features", [1024, 4096, 200])
def test_fused_dense_backward(in_features, out_features, has_bias, dtype):
    """tests/ops/test_fused_dense.py:12-59 with this library's GEMM forced on ("own" backend)."""
    from backpacks_flash_attn_b200.ops import fused_dense as FD
    rtol, atol = (3e-3, 1e-2) if dtype == torch.bfloat16 else (3e-3, 1e-3)
    torch.manual_seed(0)
    x_pt = torch.randn(8, 512, in_features, device="cuda", dtype=dtype, requires_grad=True)
    x = x_pt.detach().clone().requires_grad_()
    model_pt = torch.nn.Linear(in_features, out_features, bias=has_bias, device="cuda", dtype=dtype)
    model = FD.FusedDense(in_features, out_features, bias=has_bias, device="cuda", dtype=dtype)
    model.load_state_dict(model_pt.state_dict())
    FD.set_linear_backend("own")
    try:
        out = model(x)
        out_pt = model_pt(x_pt)
        assert torch.allclose(out, out_pt, rtol=rtol, atol=atol)
        g = torch.randn_like(out) / 32
        out.backward(g)
        out_pt.backward(g)
    finally:
        FD.set_linear_backend("auto")
    assert torch.allclose(x.grad, x_pt.grad, rtol=rtol, atol=atol)
    assert torch.allclose(model.weight.grad, model_pt.weight.grad, rtol=rtol, atol=atol * 10)
    if has_bias:
        assert torch.allclose(model.bias.grad, model_pt.bias.grad, rtol=rtol, atol=atol * 5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("out_features", [1024, 776])
@pytest.mark.parametrize("in_features,hidden", [(1024, 4096), (768, 3072), (200, 512)])
def test_fused_dense_gelu_dense_backward(in_features, hidden, out_features, dtype):
    """tests/ops/test_fused_dense.py:62-117: FusedDenseGeluDense against Linear -> gelu(tanh) -> Linear."""
    from backpacks_flash_attn_b200.ops.fused_dense import FusedDenseGeluDense
    rtol, atol = (3e-3, 3e-2) if dtype == torch.bfloat16 else (3e-3, 1e-3)
    torch.manual_seed(0)
    x_pt = torch.randn(8, 512, in_features, device="cuda", dtype=dtype, requires_grad=True)
    x = x_pt.detach().clone().requires_grad_()
    fc1 = torch.nn.Linear(in_features, hidden, device="cuda", dtype=dtype)
    fc2 = torch.nn.Linear(hidden, out_features, device="cuda", dtype=dtype)
    model = FusedDenseGeluDense(in_features, hidden, out_features, device="cuda", dtype=dtype)
    model.fc1.load_state_dict(fc1.state_dict())
    model.fc2.load_state_dict(fc2.state_dict())
    out_pt = fc2(F.gelu(fc1(x_pt), approximate="tanh"))
    out = model(x)
    assert torch.allclose(out, out_pt, rtol=rtol, atol=atol)
    g = torch.randn_like(out) / 32
    out.backward(g)
    out_pt.backward(g)
    assert torch.allclose(x.grad, x_pt.grad, rtol=rtol, atol=atol)
    assert torch.allclose(model.fc1.weight.grad, fc1.weight.grad, rtol=rtol, atol=atol * 10)
    assert torch.allclose(model.fc1.bias.grad, fc1.bias.grad, rtol=rtol, atol=atol * 5)
    assert torch.allclose(model.fc2.weight.grad, fc2.weight.grad, rtol=rtol, atol=atol * 10)
    assert torch.allclose(model.fc2.bias.grad, fc2.bias.grad, rtol=rtol, atol=atol * 5)
    # and against fp32 autograd of the oracle's MLP with the 2x rule on the input gradient
    xr = x_pt.detach().float().requires_grad_()
    params = [p.detach().float().requires_grad_() for p in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    O.mlp(xr, *params).backward(g.float())
    assert O.max_abs(x.grad, xr.grad) <= 2 * O.max_abs(x_pt.grad, xr.grad) + 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k", [(4096, 3072, 768), (1000, 776, 200), (256, 256, 64), (65536, 3072, 768)])
def test_pre_activation_output_of_the_forward_gemm(dtype, m, n, k):
    """bp_linear_bias_act_aux_fwd (the reference's linear_gelu_forward with save_gelu_in, fused_dense.py:220-222): the
    activated output is bit-identical to the plain kernel's, the pre-activation to the activation-free kernel's."""
    from backpacks_flash_attn_b200.ops.fused_dense import _linear_bias_act_aux, linear_bias_act
    torch.manual_seed(m + n)
    x = torch.randn(m, k, device="cuda", dtype=dtype)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(dtype)
    b = torch.randn(n, device="cuda", dtype=dtype)
    out, pre = _linear_bias_act_aux(x, w, b, "gelu_tanh")
    assert torch.equal(out, linear_bias_act(x, w, b, "gelu_tanh"))
    assert torch.equal(pre, linear_bias_act(x, w, b, "none"))
    out2, pre2 = _linear_bias_act_aux(x, w, None, "gelu_tanh")
    assert torch.equal(out2, linear_bias_act(x, w, None, "gelu_tanh")) and torch.equal(pre2, linear_bias_act(x, w, None, "none"))
    if m <= 4096:
        ref = x.float() @ w.float().t() + b.float()
        assert (pre.float() - ref).abs().max() <= 2 * (F.linear(x, w, b).float() - ref).abs().max() + 1e-3


@pytest.mark.parametrize("checkpoint_lvl", [0, 1, 2])
def test_checkpoint_levels_give_the_same_gradients(checkpoint_lvl):
    """checkpoint_lvl 0 / 1 keep the pre-activation the forward kernel wrote, 2 recomputes it in the backward
    (fused_dense.py:262-266): same bits either way, and the launch counters show which path ran."""
    from backpacks_flash_attn_b200 import _lib
    from backpacks_flash_attn_b200.ops.fused_dense import FusedDenseGeluDense
    torch.manual_seed(1)
    ref = FusedDenseGeluDense(256, 1024, 256, checkpoint_lvl=2, device="cuda", dtype=torch.bfloat16)
    mod = FusedDenseGeluDense(256, 1024, 256, checkpoint_lvl=checkpoint_lvl, device="cuda", dtype=torch.bfloat16)
    mod.load_state_dict(ref.state_dict())
    x = torch.randn(4, 300, 256, device="cuda", dtype=torch.bfloat16)
    g = torch.randn(4, 300, 256, device="cuda", dtype=torch.bfloat16)
    grads = []
    for m_ in (ref, mod):
        xi = x.clone().requires_grad_()
        before = dict(_lib.launch_counts)
        m_(xi).backward(g)
        used_aux = _lib.launch_counts.get("bp_linear_bias_act_aux_fwd", 0) - before.get("bp_linear_bias_act_aux_fwd", 0)
        grads.append((xi.grad, m_.fc1.weight.grad, m_.fc1.bias.grad, m_.fc2.weight.grad, used_aux))
    assert grads[0][4] == 0 and grads[1][4] == (0 if checkpoint_lvl == 2 else 1)
    for a, c in zip(grads[0][:4], grads[1][:4]):
        assert torch.equal(a, c)


def test_bias_act_backward_kernel_directly():
    from backpacks_flash_attn_b200.ops.fused_dense import bias_act_backward
    torch.manual_seed(1)
    for m, n in [(65536, 3072), (1000, 776), (7, 8)]:
        d = torch.randn(m, n, device="cuda").bfloat16()
        pre = (2 * torch.randn(m, n, device="cuda")).bfloat16()
        dpre, db = bias_act_backward(d, pre, "gelu_tanh", True)
        pr = pre.float().requires_grad_()
        F.gelu(pr, approximate="tanh").backward(d.float())
        assert O.max_abs(dpre, pr.grad) < 2e-2
        want = dpre.float().sum(0)                                        # the output is rounded to bf16 once
        assert O.max_abs(db, want) <= 2 ** -8 * want.abs().max().item() + 1e-3 * m ** 0.5
        same, db2 = bias_act_backward(d, None, "none", True)
        want2 = d.float().sum(0)
        assert same is d and O.max_abs(db2, want2) <= 2 ** -8 * want2.abs().max().item() + 1e-3 * m ** 0.5
        dpre_b, db_b = bias_act_backward(d, pre, "gelu_tanh", True)
        assert torch.equal(dpre, dpre_b) and torch.equal(db, db_b)        # deterministic
