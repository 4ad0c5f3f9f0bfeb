"""GPU parity of bp_lm_head_stats_fwd (LM head with the softmax statistics fused into the GEMM epilogue) against plain
fp32 PyTorch on the same 16-bit inputs, and of the model-level entry points built on it.

Reference semantics: logits = lm_head(hidden) (training/src/models/backpack.py:349), cross-entropy over them
(flash_attn/losses/cross_entropy.py:19-129 -> csrc/xentropy/xentropy_kernel.cu:430-760), greedy next token =
logits[:, -1].argmax (training/src/utils/generation.py:34-44).  The reference's own xentropy test compares with
torch.nn.CrossEntropyLoss at rtol 1e-5 / atol 1e-6 in fp32 (tests/losses/test_cross_entropy.py); here the logits come
out of a 16-bit GEMM, so the bar is: |lse - lse_fp32| < 2e-3 and the same for the target logit (fp32 accumulators
against an fp32 matmul of the same inputs), arg-max exact wherever the fp32 top-2 margin exceeds that noise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, n_valid):
    logits = F.linear(x.float(), w.float())[..., :n_valid]
    return logits, torch.logsumexp(logits, -1), logits.argmax(-1), logits.max(-1).values


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k,n_valid", [(1024, 50264, 768, 50264), (1024, 50264, 768, 50257), (300, 1000, 64, 1000),
                                            (257, 520, 128, 513), (1, 256, 64, 256), (70000, 512, 64, 512), (5000, 4096, 256, 4096)])
def test_lm_head_stats_matches_fp32(m, n, k, n_valid, dtype):
    from backpacks_flash_attn_b200.ops.lm_head import lm_head_stats
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = torch.randn(m, k, device="cuda", generator=g).to(dtype)
    w = (torch.randn(n, k, device="cuda", generator=g) * k ** -0.5 * 2).to(dtype)
    t = torch.randint(0, n_valid, (m,), device="cuda", generator=g)
    t[0], t[-1] = 0, n_valid - 1                      # first / last vocabulary column
    out = lm_head_stats(x, w, t, n_valid=n_valid)
    logits, lse, amax, mx = _ref(x, w, n_valid)
    assert (out["lse"] - lse).abs().max() < 2e-3
    assert (out["max_logit"] - mx).abs().max() < 2e-3
    assert (out["target_logit"] - logits.gather(-1, t[:, None])[:, 0]).abs().max() < 2e-3
    top2 = logits.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4e-3         # rows whose arg-max is not a numerical coin toss
    assert clear.float().mean() > 0.9
    assert torch.equal(out["argmax"][clear].long(), amax[clear])
    # the arg-max always points at a logit within the noise of the true maximum
    assert (logits.gather(-1, out["argmax"].long()[:, None])[:, 0] - mx).abs().max() < 4e-3
    again = lm_head_stats(x, w, t, n_valid=n_valid)
    assert all(torch.equal(out[k_], again[k_]) for k_ in out)          # deterministic


def test_lm_head_stats_ties_and_shapes():
    """Exact ties resolve to the first column (torch.argmax order); leading dims are kept; targets optional."""
    from backpacks_flash_attn_b200.ops.lm_head import lm_head_stats
    x = torch.ones(2, 5, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(600, 64, device="cuda", dtype=torch.bfloat16)
    w[[7, 300, 599]] = 1.0                            # three identical maxima in different half-tiles
    out = lm_head_stats(x, w)
    assert out["argmax"].shape == (2, 5) and (out["argmax"] == 7).all() and "target_logit" not in out
    assert (out["max_logit"] - 64.0).abs().max() == 0
    with pytest.raises(RuntimeError, match="same dtype"):
        lm_head_stats(x, w.half())
    with pytest.raises(RuntimeError, match="targets must be"):
        lm_head_stats(x, w, torch.zeros(2, 4, dtype=torch.int64, device="cuda"))


def test_lm_head_cross_entropy_matches_torch():
    from backpacks_flash_attn_b200.ops.lm_head import lm_head_cross_entropy
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(4, 333, 256, device="cuda", generator=g).bfloat16()
    w = (torch.randn(5000, 256, device="cuda", generator=g) * 0.1).bfloat16()
    t = torch.randint(0, 5000, (4, 333), device="cuda", generator=g)
    t[0, :50] = -100                                   # ignored positions
    logits = F.linear(x.float(), w.float())
    for red in ("mean", "sum", "none"):
        want = F.cross_entropy(logits.view(-1, 5000), t.view(-1), ignore_index=-100, reduction=red)
        got = lm_head_cross_entropy(x, w, t, reduction=red)
        torch.testing.assert_close(got.reshape(want.shape), want, rtol=2e-4, atol=2e-3)
    # against the reference's own pipeline (16-bit logits, then fp32 cross-entropy): ours is at least as close to fp32
    ref16 = F.cross_entropy(F.linear(x, w).float().view(-1, 5000), t.view(-1), ignore_index=-100)
    exact = F.cross_entropy(logits.view(-1, 5000), t.view(-1), ignore_index=-100)
    assert (lm_head_cross_entropy(x, w, t) - exact).abs() <= (ref16 - exact).abs() + 1e-4


def test_model_last_tokens_loss_and_greedy_token():
    """BackpackLMHeadModel: forward(num_last_tokens=1) == forward()[:, -1:]; token_stats / loss agree with the logits
    the full forward writes (which the fused path never forms)."""
    from backpacks_flash_attn_b200 import _lib
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
    from backpacks_flash_attn_b200.utils.weights import name_seeded_
    model = name_seeded_(BackpackLMHeadModel(flash_config(n_embd=128, n_head=2, n_layer=2, n_positions=256)).eval())
    model = model.to("cuda", torch.bfloat16)
    ids = torch.randint(0, 50257, (3, 200), device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    with torch.inference_mode():
        full = model(ids).logits
        last = model(ids, num_last_tokens=1).logits
        assert last.shape == (3, 1, 50264)
        assert (last.float() - full[:, -1:].float()).abs().max() < 3e-2     # another GEMM shape, same values up to rounding
        before = _lib.launch_counts.get("bp_lm_head_stats_fwd", 0)
        st = model.token_stats(ids, targets=ids)
        assert _lib.launch_counts.get("bp_lm_head_stats_fwd", 0) == before + 1
        lf = full.float()
        assert (st["lse"] - torch.logsumexp(lf, -1)).abs().max() < 3e-2      # full holds bf16-rounded logits
        agree = (st["argmax"].long() == lf.argmax(-1)).float().mean()
        assert agree > 0.97, agree
        labels = torch.roll(ids, -1, 1)
        labels[:, -1] = -100
        want = F.cross_entropy(lf.view(-1, lf.shape[-1]), labels.view(-1), ignore_index=-100)
        assert (model.loss(ids, labels) - want).abs() < 2e-2
        nxt = model.token_stats(ids, num_last_tokens=1)["argmax"][:, 0]
        assert nxt.shape == (3,)
