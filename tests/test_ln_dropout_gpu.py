"""GPU parity of the residual dropout inside the LayerNorm kernels (bp_ln_residual_fwd_dropout / _bwd_dropout behind
`dropout_add_layer_norm(..., dropout_p > 0)`), mirroring the reference's training test
(tests/ops/test_dropout_layer_norm.py:54-106): the keep mask the operator returns is applied in a same-precision PyTorch
composition and in an fp32 one, and the rules are the reference's -- output and input gradients <= 4 x the PyTorch error
+ 1e-4, weight / bias gradients <= 2 x + 3e-5.  The mask is the Python restatement of the kernels' hash
(ops/layer_norm.layer_norm_dropout_mask); the residual output proves the kernel applied exactly that mask."""
import pytest
import torch
import torch.nn.functional as F

from backpacks_flash_attn_b200 import _lib
from backpacks_flash_attn_b200.flash_attn_interface import effective_dropout_p
from backpacks_flash_attn_b200.ops.layer_norm import DropoutAddLayerNorm, dropout_add_layer_norm, layer_norm_dropout_mask
from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hidden", [768, 256, 1024, 200, 2048])
@pytest.mark.parametrize("has_residual", [True, False])
@pytest.mark.parametrize("prenorm", [True, False])
@pytest.mark.parametrize("in_dtype,res_dtype,w_dtype", [
    (torch.bfloat16, torch.float32, torch.bfloat16), (torch.float16, torch.float16, torch.float16),
    (torch.float32, torch.float32, torch.float32)])
def test_dropout_add_layer_norm_training(in_dtype, res_dtype, w_dtype, prenorm, has_residual, hidden):
    torch.manual_seed(hidden + 2 * prenorm + has_residual)
    p = 0.37
    pe = effective_dropout_p(p)
    rows = (6, 133)
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
    before = dict(_lib.launch_counts)
    res = dropout_add_layer_norm(x0, x1, w, b, p, 1e-5, prenorm=prenorm, residual_in_fp32=residual_in_fp32,
                                 return_dropout_mask=True)
    out, resid, mask = res if prenorm else (res[0], None, res[1])
    assert _lib.launch_counts.get("bp_ln_residual_fwd_dropout", 0) > before.get("bp_ln_residual_fwd_dropout", 0)
    assert mask.shape == x0.shape and mask.dtype == torch.bool
    assert abs(mask.float().mean().item() - (1 - pe)) < 0.01
    keep = mask.float() / (1 - pe)
    # same-precision PyTorch composition (as the reference's test builds it) and the fp32 one
    if has_residual:
        r_pt = ((x0_pt.float() * keep) + x1_pt.float()).to(res_dtype)
        r_ref = x0_ref * keep + x1_ref
    else:
        r_pt = (x0_pt.float() * keep).to(res_dtype if residual_in_fp32 else in_dtype)
        r_ref = x0_ref * keep
    out_pt = F.layer_norm(r_pt.to(w_dtype), (hidden,), w_pt, b_pt, 1e-5).to(in_dtype)
    out_ref = F.layer_norm(r_ref, (hidden,), w_ref, b_ref, 1e-5)
    assert O.max_abs(out, out_ref) <= 4 * O.max_abs(out_pt, out_ref) + 1e-4
    g = torch.randn_like(out) / rows[0]
    if prenorm:
        assert resid.dtype == r_pt.dtype
        # the residual stream shows the mask the kernel applied: dropped elements carry x1 alone
        shown = (resid.float() - (x1.detach().float() if has_residual else 0)) != 0
        assert not (shown & ~mask).any()                      # a dropped element never contributes
        if resid.dtype == torch.float32 or not has_residual:  # (a 16-bit sum can absorb a small kept x0)
            assert torch.equal(shown | (x0.detach() == 0), mask | (x0.detach() == 0))
        assert O.max_abs(resid, r_ref.detach()) <= 4 * O.max_abs(r_pt.detach(), r_ref.detach()) + 1e-4
        g2 = torch.randn(*rows, hidden, device="cuda") / rows[0]
        (out.float() * g.float()).sum().add((resid.float() * g2).sum()).backward()
        (out_pt.float() * g.float()).sum().add((r_pt.float() * g2).sum()).backward()
        (out_ref * g.float()).sum().add((r_ref * g2).sum()).backward()
    else:
        out.backward(g)
        out_pt.backward(g)
        out_ref.backward(g.float())
    assert _lib.launch_counts.get("bp_ln_residual_bwd_dropout", 0) > before.get("bp_ln_residual_bwd_dropout", 0)
    assert x0.grad.dtype == in_dtype
    assert (x0.grad[~mask] == 0).all()
    assert O.max_abs(x0.grad, x0_ref.grad) <= 4 * O.max_abs(x0_pt.grad, x0_ref.grad) + 1e-4
    if has_residual:
        assert x1.grad.dtype == res_dtype
        assert O.max_abs(x1.grad, x1_ref.grad) <= 4 * O.max_abs(x1_pt.grad, x1_ref.grad) + 1e-4
    assert O.max_abs(w.grad, w_ref.grad) <= 2 * O.max_abs(w_pt.grad, w_ref.grad) + 3e-5
    assert O.max_abs(b.grad, b_ref.grad) <= 2 * O.max_abs(b_pt.grad, b_ref.grad) + 3e-5


def test_seed_semantics_and_zero_probability():
    x0 = torch.randn(300, 768, device="cuda", dtype=torch.bfloat16)
    x1 = torch.randn(300, 768, device="cuda")
    w = torch.randn(768, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(768, device="cuda", dtype=torch.bfloat16)
    z1, r1, m1 = dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5, prenorm=True, return_dropout_mask=True, seed=77)
    z2, r2, m2 = dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5, prenorm=True, return_dropout_mask=True, seed=77)
    z3, r3, m3 = dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5, prenorm=True, return_dropout_mask=True, seed=78)
    assert torch.equal(z1, z2) and torch.equal(r1, r2) and torch.equal(m1, m2)
    assert not torch.equal(m1, m3) and not torch.equal(r1, r3)
    assert torch.equal(m1, layer_norm_dropout_mask(77, 300, 768, 0.1, device="cuda"))
    assert torch.equal(m1.cpu(), layer_norm_dropout_mask(77, 300, 768, 0.1))
    # rows and columns are both decorrelated
    assert not torch.equal(m1[0], m1[1]) and not torch.equal(m1[:, 0], m1[:, 1])
    # the default seed follows torch.manual_seed
    torch.manual_seed(5)
    a = dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5)
    torch.manual_seed(5)
    assert torch.equal(a, dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5))
    assert not torch.equal(a, dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5))
    # dropout_p == 0 through the dropout entry point is the plain kernel
    lib = _lib.load()
    za, zb = torch.empty_like(x0), torch.empty_like(x0)
    ra, rb = torch.empty_like(x1), torch.empty_like(x1)
    st = _lib.stream_ptr(x0.device)
    common = (x0.data_ptr(), x1.data_ptr(), w.data_ptr(), b.data_ptr())
    _lib.check(lib.bp_ln_residual_fwd(*common, za.data_ptr(), ra.data_ptr(), None, None, 300, 768, 1e-5, 1, 2, 1, st), "a")
    _lib.check(lib.bp_ln_residual_fwd_dropout(*common, zb.data_ptr(), rb.data_ptr(), None, None, 300, 768, 1e-5, 1, 2, 1,
                                              0.0, 123, st), "b")
    assert torch.equal(za, zb) and torch.equal(ra, rb)
    assert lib.bp_ln_residual_fwd_dropout(*common, zb.data_ptr(), rb.data_ptr(), None, None, 300, 768, 1e-5, 1, 2, 1,
                                          1.0, 123, st) == -1
    assert "dropout_p" in _lib.last_error()


def test_module_applies_dropout_only_in_training_mode():
    m = DropoutAddLayerNorm(256, prenorm=True, p=0.5, residual_in_fp32=True, device="cuda", dtype=torch.bfloat16)
    x0 = torch.randn(64, 256, device="cuda", dtype=torch.bfloat16)
    x1 = torch.randn(64, 256, device="cuda")
    m.eval()
    z, r = m(x0, x1)
    assert torch.equal(r, x0.float() + x1)
    m.train()
    z, r = m(x0, x1)
    dropped = (r == x1).float().mean().item()
    assert 0.45 < dropped < 0.55, dropped


def test_full_size_rows_and_determinism():
    x0 = torch.randn(65536, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    x1 = torch.randn(65536, 768, device="cuda", requires_grad=True)
    w = torch.ones(768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    b = torch.zeros(768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    grads = []
    for _ in range(2):
        z, r, mask = dropout_add_layer_norm(x0, x1, w, b, 0.1, 1e-5, prenorm=True, return_dropout_mask=True, seed=9)
        g = torch.autograd.grad((z.float().square().sum() + r.sum()), (x0, x1, w, b))
        grads.append(g)
    for a, c in zip(*grads):
        assert torch.equal(a, c)
    pe = effective_dropout_p(0.1)
    assert abs(mask.float().mean().item() - (1 - pe)) < 1e-3
    want = x0.detach().float() * mask / (1 - pe) + x1.detach()
    assert (r.detach() - want).abs().max() < 1e-5
    assert (grads[0][0][~mask] == 0).all()
