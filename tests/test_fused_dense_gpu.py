"""GPU parity of bp_linear_bias_act_fwd (FusedDenseGeluDense.fc1 + GELU, FusedDense) -- mirrors
tests/ops/test_fused_dense.py:12-117 of the reference: compare with nn.Linear / F.gelu(approximate='tanh')
computed in fp32 on the same 16-bit inputs; rtol 3e-3, atol 3e-2 (bf16) / 1e-2... as the reference states,
plus the stricter "<= 2x the same-precision PyTorch error" rule used for the attention kernels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("act", ["gelu_tanh", "none"])
@pytest.mark.parametrize("has_bias", [True, False])
@pytest.mark.parametrize("m,n,k", [(1024, 3072, 768), (517, 1024, 1024), (128, 256, 64), (1, 8, 8),
                                    (300, 776, 200), (4096, 768, 3072), (2048, 12288, 768), (777, 50264, 768)])
def test_linear_bias_act(m, n, k, has_bias, act, dtype):
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k, device="cuda").to(dtype)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(dtype)
    b = (torch.randn(n, device="cuda") * 0.5).to(dtype) if has_bias else None
    out = linear_bias_act(x, w, b, act)
    pre = F.linear(x.float(), w.float(), b.float() if has_bias else None)
    ref = F.gelu(pre, approximate="tanh") if act == "gelu_tanh" else pre
    pt_pre = F.linear(x, w, b)
    pt = F.gelu(pt_pre, approximate="tanh") if act == "gelu_tanh" else pt_pre
    assert out.shape == (m, n) and out.dtype == dtype
    torch.testing.assert_close(out.float(), ref, rtol=3e-3, atol=3e-2 if dtype == torch.bfloat16 else 1e-2)
    assert O.max_abs(out, ref) <= 2 * O.max_abs(pt, ref) + 1e-3


def test_fused_dense_gelu_dense_module_matches_reference_fixture(golden_dir):
    """Mlp fixture generated from the real reference (tests/golden/ops.npz: mlp_*)."""
    from backpacks_flash_attn_b200.ops.fused_dense import FusedDenseGeluDense
    g = np.load(f"{golden_dir}/ops.npz")
    mod = FusedDenseGeluDense(64, 256, 64, device="cuda", dtype=torch.bfloat16).eval()
    with torch.no_grad():
        mod.fc1.weight.copy_(torch.from_numpy(g["mlp_w1"]))
        mod.fc1.bias.copy_(torch.from_numpy(g["mlp_b1"]))
        mod.fc2.weight.copy_(torch.from_numpy(g["mlp_w2"]))
        mod.fc2.bias.copy_(torch.from_numpy(g["mlp_b2"]))
        y = mod(torch.from_numpy(g["mlp_x"]).cuda().bfloat16())
    assert (y.float().cpu() - torch.from_numpy(g["mlp_y"])).abs().max() < 3e-2


def test_linear_bias_act_3d_input_and_errors():
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act, FusedDense
    x = torch.randn(4, 33, 64, device="cuda").bfloat16()
    w = torch.randn(128, 64, device="cuda").bfloat16() * 0.1
    out = linear_bias_act(x, w, None, "none")
    assert out.shape == (4, 33, 128)
    assert O.max_abs(out, F.linear(x.float(), w.float())) < 3e-2
    with pytest.raises(RuntimeError, match="same dtype"):
        linear_bias_act(x, w.half(), None)
    lin = FusedDense(64, 128, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        assert lin(x).shape == (4, 33, 128)
