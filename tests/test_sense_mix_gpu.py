"""GPU parity of the fused sense-mix (bp_sense_lse_fwd + bp_sense_mix_fwd) against the oracle.

The reference has no test for this operator; the criterion is the one its fmha tests use:
    max|ours - fp32| <= 2 * max|reference eager in the same precision - fp32|
with the eager composition restated from backpack.py:116-122 + :313 and pinned by tests/golden/ops.npz.
"""
import numpy as np
import pytest
import torch

from oracle import backpack_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(b, s, nv, d, dtype, seed=0, transposed=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda", generator=g).to(dtype)
    if transposed:   # the layout the reference's content model returns (backpack.py:276)
        content = torch.randn(b, s, nv, d, device="cuda", generator=g).to(dtype).transpose(1, 2)
    else:
        content = torch.randn(b, nv, s, d, device="cuda", generator=g).to(dtype)
    return qk, content


def _check(qk, content):
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    out, lse = sense_mix(qk, content, return_lse=True)
    ref, lse_ref = O.sense_mix_fp32_ref(qk, content)
    eager = O.sense_mix_eager(qk, content)
    err, err_eager = O.max_abs(out, ref), O.max_abs(eager, ref)
    assert O.max_abs(lse, lse_ref) < 1e-3
    assert err <= 2 * err_eager + 1e-5, f"max err {err:.3e} vs eager {err_eager:.3e}"
    assert O.mean_abs(out, ref) <= 2 * O.mean_abs(eager, ref) + 1e-6
    return err, err_eager


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("s", [64, 97, 128, 200, 256, 257, 512, 1024])
@pytest.mark.parametrize("nv,d", [(16, 768), (16, 384), (4, 768), (8, 128), (1, 64)])
def test_sense_mix_matches_oracle(s, nv, d, dtype):
    qk, content = _inputs(2, s, nv, d, dtype, seed=s + nv)
    _check(qk, content)


def test_sense_mix_contiguous_content_and_determinism():
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qk, content = _inputs(3, 384, 16, 768, torch.bfloat16, seed=5, transposed=False)
    _check(qk, content)
    first = sense_mix(qk, content)
    for _ in range(5):
        assert torch.equal(first, sense_mix(qk, content))
    # same values through the transposed layout give bit-identical output
    c2 = content.transpose(1, 2).contiguous().transpose(1, 2)
    assert torch.equal(first, sense_mix(qk, c2))


def test_sense_mix_linearity_in_content():
    """Size-independent property: the operator is linear in `content` (what the intervention wrappers
    rely on, intervened_models.py:78-101): mix(qk, a*C1 + C2) == a*mix(qk, C1) + mix(qk, C2) up to rounding."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qk, c1 = _inputs(2, 512, 16, 768, torch.bfloat16, seed=11)
    _, c2 = _inputs(2, 512, 16, 768, torch.bfloat16, seed=12)
    lhs = sense_mix(qk, (2.0 * c1 + c2))
    rhs = 2.0 * sense_mix(qk, c1).float() + sense_mix(qk, c2).float()
    assert (lhs.float() - rhs).abs().max() < 0.25 and (lhs.float() - rhs).abs().mean() < 1e-2
    zero = sense_mix(qk, torch.zeros_like(c1))
    assert zero.abs().max() == 0


def test_sense_mix_rows_of_alpha_sum_to_one():
    """content == 1 everywhere  =>  out == nv (every alpha row sums to one, for every sense)."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qk, c = _inputs(2, 1024, 16, 768, torch.bfloat16, seed=13)
    out = sense_mix(qk, torch.ones_like(c))
    assert (out.float() - 16.0).abs().max() < 0.15


def test_sense_mix_golden_reference_fixture(golden_dir):
    """Fixture generated from the real reference (tests/golden/make_golden.py): ContextSelfAttn -> sense sum."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    g = np.load(f"{golden_dir}/ops.npz")
    h, w, b = (torch.from_numpy(g[k]).cuda() for k in ("ctx_h", "ctx_w", "ctx_b"))
    content = torch.from_numpy(g["ctx_content_bsnd"]).cuda().transpose(1, 2)
    qk = torch.nn.functional.linear(h, w, b).reshape(2, 64, 2, 8, 16)
    out = sense_mix(qk.bfloat16(), content.bfloat16())
    ref = torch.from_numpy(g["ctx_sense_sum"]).cuda()
    eager = O.sense_mix_eager(qk.bfloat16(), content.bfloat16())
    assert O.max_abs(out, ref) <= 2 * O.max_abs(eager, ref) + 1e-5


def test_sense_mix_config3_shape():
    """BASELINE config 3 operator shape (b=64 is exercised by bench.py; b=8 keeps the oracle cheap)."""
    qk, content = _inputs(8, 1024, 16, 768, torch.bfloat16, seed=3)
    err, err_eager = _check(qk, content)
    print(f"sense-mix s1024 k16 d768: max err {err:.3e} (eager bf16 reference {err_eager:.3e})")


def test_sense_mix_rejects_bad_arguments():
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qk, content = _inputs(1, 64, 16, 768, torch.bfloat16)
    with pytest.raises(RuntimeError, match="same dtype"):
        sense_mix(qk, content.half())
    with pytest.raises(RuntimeError, match="content must be"):
        sense_mix(qk, content[:, :8])
    with pytest.raises(RuntimeError, match="multiple of 64"):
        sense_mix(torch.zeros(1, 64, 2, 4, 8, device="cuda", dtype=torch.bfloat16),
                  torch.zeros(1, 4, 64, 32, device="cuda", dtype=torch.bfloat16))


@pytest.mark.parametrize("s", [128, 512, 1024])
def test_sense_mix_k64_odd_key_width(s):
    """BASELINE config 5: k = 64 senses of a 768-wide model => sense key width 12 (zero-padded to 16 for TMA)."""
    qk, content = _inputs(2, s, 64, 768, torch.bfloat16, seed=s)
    assert qk.shape[-1] == 12
    _check(qk, content)


def _per_batch_check(qk, content):
    """2x rule + LSE bound with the oracle evaluated one batch element at a time (alpha is (nv, s, s) fp32 per
    element: 4.3 GB at k = 64, s = 4096)."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    out, lse = sense_mix(qk, content, return_lse=True)
    for i in range(qk.shape[0]):
        ref, lse_ref = O.sense_mix_fp32_ref(qk[i:i + 1], content[i:i + 1])
        eager = O.sense_mix_eager(qk[i:i + 1], content[i:i + 1])
        err, err_eager = O.max_abs(out[i:i + 1], ref), O.max_abs(eager, ref)
        assert O.max_abs(lse[i:i + 1], lse_ref) < 1e-3
        assert err <= 2 * err_eager + 1e-5, f"batch {i}: max err {err:.3e} vs eager {err_eager:.3e}"
        assert O.mean_abs(out[i:i + 1], ref) <= 2 * O.mean_abs(eager, ref) + 1e-6
        del ref, lse_ref, eager


@pytest.mark.parametrize("nv", [4, 16, 64])
@pytest.mark.parametrize("s,b", [(2048, 8), (4096, 4), (1024, 16)])
def test_sense_mix_config5_cells(s, b, nv):
    """Every sense-mix cell of BASELINE config 5 that profiles/ reports a time for (seq x senses at 16384 tokens,
    d = 768); the s <= 1024 cells at k in {4, 16} are also covered by test_sense_mix_matches_oracle."""
    qk, content = _inputs(b, s, nv, 768, torch.bfloat16, seed=s + nv)
    _per_batch_check(qk, content)


@pytest.mark.parametrize("dtype,bound", [(torch.float16, 1e-3), (torch.bfloat16, 7e-3)])
@pytest.mark.parametrize("s,nv,d", [(512, 16, 768), (257, 4, 768), (1024, 8, 128)])
def test_sense_mix_fp32_output_mode(s, nv, d, dtype, bound):
    """T2 of SURVEY.md §8c for the sense-mix: accumulator stored before the final 16-bit rounding, against the fp32
    oracle on identical 16-bit inputs.  As in the attention kernel the remaining error is the rounding of P to the
    16-bit MMA operand type (rel. 2^-12 for fp16, 2^-9 for bf16), here summed over nv senses whose outputs add up
    (independent roundings of O(1) terms: the maximum grows like sqrt(nv)).  Measured on B200: fp16 2.4e-3 max /
    8.7e-5 mean at 16 senses (the north-star 1e-3 bar holds per sense: bound = 1e-3 * sqrt(nv)), bf16 1.8e-2 max /
    6.9e-4 mean (bound = 7e-3 * sqrt(nv)); the MEAN error is below 1e-3 for both types and the LSE is within 1e-3.
    The production output must be exactly the rounding of the test-mode output."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qk, content = _inputs(2, s, nv, d, dtype, seed=3 * s + nv)
    out32, lse = sense_mix(qk, content, return_lse=True, out_fp32=True)
    assert out32.dtype == torch.float32
    ref, lse_ref = O.sense_mix_fp32_ref(qk, content)
    err, mean = O.max_abs(out32, ref), O.mean_abs(out32, ref)
    print(f"sense-mix fp32-output mode s{s} k{nv} d{d} {dtype}: max|err| {err:.2e} mean {mean:.2e}")
    assert err < bound * nv ** 0.5, err
    assert mean < 1e-3
    assert O.max_abs(lse, lse_ref) < 1e-3
    assert torch.equal(sense_mix(qk, content), out32.to(dtype))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("s,nv,d,vocab", [(1024, 16, 768, 3000), (200, 16, 768, 517), (257, 4, 768, 64), (512, 8, 128, 1000),
                                          (128, 64, 768, 300)])
def test_sense_mix_table_gathers_inside_the_kernel(s, nv, d, vocab, dtype):
    """bp_sense_mix_table_fwd == bp_sense_mix_fwd on the gathered tensor, bit for bit, and within the 2x rule of the
    fp32 oracle evaluated on table[ids]."""
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix, sense_mix_table
    g = torch.Generator(device="cuda").manual_seed(s + vocab)
    b = 3
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda", generator=g).to(dtype)
    table = torch.randn(vocab, nv, d, device="cuda", generator=g).to(dtype)
    ids = torch.randint(0, vocab, (b, s), device="cuda", generator=g)
    ids[0, :4] = torch.tensor([0, vocab - 1, 0, vocab - 1], device="cuda")     # first / last table rows
    content = table[ids].transpose(1, 2)                                      # (b, nv, s, d) view, what the table replaces
    out, lse = sense_mix_table(qk, table, ids, return_lse=True)
    out_t, lse_t = sense_mix(qk, content, return_lse=True)
    assert torch.equal(lse, lse_t)
    assert torch.equal(out, out_t)
    ref, lse_ref = O.sense_mix_fp32_ref(qk, content)
    eager = O.sense_mix_eager(qk, content)
    assert O.max_abs(out, ref) <= 2 * O.max_abs(eager, ref) + 1e-5
    assert O.max_abs(lse, lse_ref) < 1e-3
    first = sense_mix_table(qk, table, ids)
    for _ in range(3):
        assert torch.equal(first, sense_mix_table(qk, table, ids))


def test_sense_mix_table_rejects_bad_arguments():
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix_table
    qk = torch.zeros(1, 64, 2, 16, 48, device="cuda", dtype=torch.bfloat16)
    table = torch.zeros(100, 16, 768, device="cuda", dtype=torch.bfloat16)
    ids = torch.zeros(1, 64, device="cuda", dtype=torch.int64)
    with pytest.raises(RuntimeError, match="int64"):
        sense_mix_table(qk, table, ids.int())
    with pytest.raises(RuntimeError, match="table must be"):
        sense_mix_table(qk, table[:, :8], ids)
    with pytest.raises(RuntimeError, match="same dtype"):
        sense_mix_table(qk, table.half(), ids)
    with pytest.raises(RuntimeError, match="CUDA"):
        sense_mix_table(qk, table, ids.cpu())
