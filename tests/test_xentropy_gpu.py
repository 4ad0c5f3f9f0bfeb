"""GPU parity of bp_xentropy_fwd / bp_xentropy_bwd behind the reference-shaped CrossEntropyLoss -- mirrors
tests/losses/test_cross_entropy.py:14-40 of the reference: against torch.nn.CrossEntropyLoss on the fp32-upcast
logits, rtol 1e-5 / atol 1e-6 (fp32) or 1e-3 / 1e-4 (16-bit), with label smoothing, ignored targets and the in-place
backward; plus the padded Backpack vocabulary (50264, the vectorised path) and strided rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("inplace_backward", [False, True])
@pytest.mark.parametrize("smoothing", [0.0, 0.9])
@pytest.mark.parametrize("vocab_size", [50257, 50264, 1000])
def test_cross_entropy_loss(vocab_size, smoothing, inplace_backward, dtype):
    from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLossApex
    rtol, atol = (1e-5, 1e-6) if dtype == torch.float32 else (1e-3, 1e-4)
    if dtype == torch.bfloat16:
        rtol, atol = 1e-2, 1e-4          # the gradient is ROUNDED to bf16 (8 bits); the reference's test stops at fp16
    torch.random.manual_seed(0)
    n = 8 * 128
    x_pt = torch.randn(n, vocab_size, device="cuda", dtype=dtype, requires_grad=True)
    x = x_pt.detach().clone().requires_grad_()
    y = torch.randint(0, vocab_size, (n,), dtype=torch.long, device="cuda")
    y[torch.randperm(n)[:10]] = -100
    model_pt = torch.nn.CrossEntropyLoss(label_smoothing=smoothing)
    model = CrossEntropyLossApex(label_smoothing=smoothing, inplace_backward=inplace_backward)
    out = model(x, y)
    out_pt = model_pt(x_pt.float(), y)
    assert out.dtype == torch.float32
    assert torch.allclose(out, out_pt, rtol=1e-5 if dtype == torch.float32 else 1e-3, atol=1e-6 if dtype == torch.float32 else 1e-4)
    g = torch.randn_like(out)
    out_pt.backward(g)
    out.backward(g)
    assert torch.allclose(x.grad.float(), x_pt.grad.float(), rtol=rtol, atol=atol)


def test_cross_entropy_none_reduction_strided_rows_and_errors():
    from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss, SoftmaxCrossEntropyLossFn
    torch.manual_seed(1)
    big = torch.randn(300, 50264 + 8, device="cuda").bfloat16()
    x = big[:, :50264]                                   # row stride 50272: still 16-byte aligned rows
    y = torch.randint(0, 50264, (300,), device="cuda")
    y[5] = -100
    loss = CrossEntropyLoss(reduction="none")(x, y)
    ref = torch.nn.functional.cross_entropy(x.float(), y, reduction="none")
    assert loss[5] == 0 and torch.allclose(loss, ref, rtol=1e-3, atol=1e-4)
    xr = x.clone().requires_grad_()
    out = SoftmaxCrossEntropyLossFn.apply(xr, y, 0.0, -100, True)
    out.sum().backward()
    xf = x.float().requires_grad_()
    torch.nn.functional.cross_entropy(xf, y, reduction="sum").backward()
    assert torch.allclose(xr.grad.float(), xf.grad, rtol=1e-2, atol=1e-4)
    assert xr.grad[5].abs().max() == 0                    # ignored target: no gradient
    with pytest.raises(RuntimeError, match="CUDA"):
        CrossEntropyLoss()(x.cpu(), y.cpu())
    with pytest.raises(RuntimeError, match="int64"):
        SoftmaxCrossEntropyLossFn.apply(x, y.int())
    with pytest.raises(NotImplementedError):
        CrossEntropyLoss(reduction="sum")


def test_cross_entropy_full_size_matches_the_lm_head_statistics():
    """Config-3 sized logits (16384 x 50264 here): the loss kernel against the fused LM-head statistics path."""
    from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss
    torch.manual_seed(2)
    x = (2 * torch.randn(16384, 50264, device="cuda")).bfloat16()
    y = torch.randint(0, 50257, (16384,), device="cuda")
    loss = CrossEntropyLoss(reduction="none")(x, y)
    lse = torch.logsumexp(x.float(), -1)
    ref = lse - x.float().gather(1, y[:, None])[:, 0]
    assert (loss - ref).abs().max() < 2e-3
