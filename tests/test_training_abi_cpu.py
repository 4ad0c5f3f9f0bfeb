"""CPU checks of the training-side entry points: argument validation happens before any CUDA
call, the workspace queries are pure functions, the dropout mask restatement is well behaved, and the autograd operators
refuse host tensors (no CPU fallback)."""
import ctypes

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    from backpacks_flash_attn_b200 import _lib, build
    build.build_library()
    return _lib.load()


def _aligned(nbytes=256):
    buf = ctypes.create_string_buffer(nbytes + 16)
    return buf, (ctypes.addressof(buf) + 15) // 16 * 16


def test_workspace_queries(lib):
    # attention backward: two ticket counters (16 bytes), then per 64-row block 3 x 64 fp32 words (-lse, -delta, dropout
    # word), rows padded to 128
    assert lib.bp_fmha_bwd_workspace_bytes(2, 3, 1000) == 16 + 2 * 3 * (1024 // 64) * 192 * 4
    assert lib.bp_fmha_bwd_workspace_bytes(0, 3, 1000) == 0
    assert (lib.bp_fmha_bwd_dropout_workspace_bytes(2, 3, 1000, 130)
            == lib.bp_fmha_bwd_workspace_bytes(2, 3, 1000) + 2 * 3 * 256 * 4)
    assert lib.bp_fmha_fwd_dropout_workspace_bytes(2, 3, 130) == 2 * 3 * 256 * 4
    assert lib.bp_ln_bwd_workspace_bytes(768) == 1024 * 2 * 768 * 4
    assert lib.bp_bias_act_bwd_workspace_bytes(3072) == 256 * 3072 * 4


def test_backward_entry_points_validate_before_touching_the_gpu(lib):
    from backpacks_flash_attn_b200 import _lib
    keep, a = _aligned()
    strides = (ctypes.c_int64 * 16)(*([64, 64] * 8))
    sp = ctypes.addressof(strides)
    args = [a] * 11 + [1, 1, 64, 16, 16, 16, 16, sp, 16, 1.0, 1, 1]
    # workspace too small
    assert lib.bp_fmha_bwd(*args, a, 16, None) == -1 and "workspace" in _lib.last_error()
    # head dim not a multiple of 8
    bad = list(args)
    bad[13] = 12
    assert lib.bp_fmha_bwd(*bad, a, 1 << 20, None) == -1 and "multiple of 8" in _lib.last_error()
    # stride not a multiple of 8 elements
    odd = (ctypes.c_int64 * 16)(*([64, 60] * 8))
    bad = list(args)
    bad[18] = ctypes.addressof(odd)
    assert lib.bp_fmha_bwd(*bad, a, 1 << 20, None) == -1 and "strides" in _lib.last_error()
    # dropout probability out of range
    assert lib.bp_fmha_bwd_dropout(*args, 1.0, 7, a, 1 << 20, None) == -1 and "p_dropout" in _lib.last_error()
    fwd = [a] * 7 + [1, 1, 64, 16, 16, 16, 16] + [64] * 8 + [16, 1.0, 1, 1]
    assert lib.bp_fmha_fwd_dropout(*fwd, -0.1, 7, a, 1 << 20, None) == -1 and "p_dropout" in _lib.last_error()
    assert lib.bp_fmha_fwd_dropout(*fwd, 0.1, 7, None, 0, None) == -1 and "workspace" in _lib.last_error()
    # LayerNorm backward
    assert lib.bp_ln_residual_bwd(a, None, a, a, None, None, a, None, a, a, a, 16, 4, 64, 1e-5, 1, 2, 1, None) == -1
    assert "workspace" in _lib.last_error()
    assert lib.bp_ln_residual_bwd(a, None, a, a, None, None, a, None, a, a, a, 1 << 30, 4, 60, 1e-5, 1, 2, 1, None) == -1
    assert "multiple of 8" in _lib.last_error()
    assert lib.bp_ln_residual_bwd(None, None, a, a, None, None, a, None, a, a, a, 1 << 30, 4, 64, 1e-5, 1, 2, 1, None) == -1
    assert lib.bp_ln_residual_bwd(a, None, a, a, a, None, a, None, a, a, a, 1 << 30, 4, 64, 1e-5, 1, 2, 1, None) == -1
    assert "together" in _lib.last_error()
    # dgelu / bias gradient
    assert lib.bp_bias_act_bwd(a, None, None, a, a, 1 << 30, 8, 64, 1, 1, None) == -1 and "GELU" in _lib.last_error()
    assert lib.bp_bias_act_bwd(a, None, None, None, a, 1 << 30, 8, 64, 0, 1, None) == -1
    assert lib.bp_bias_act_bwd(a, a, a, a, a, 1 << 30, 8, 60, 1, 1, None) == -1 and "multiple of 8" in _lib.last_error()
    assert lib.bp_bias_act_bwd(a, a, a, a, a, 1 << 30, 8, 64, 1, 2, None) == -1 and "fp16 and bf16" in _lib.last_error()
    # cross entropy
    assert lib.bp_xentropy_fwd(a, a, a, a, 4, 64, 32, 0.0, -100, -1, 1, None) == -1 and "row stride" in _lib.last_error()
    assert lib.bp_xentropy_fwd(a, a, a, a, 4, 64, 64, 1.5, -100, -1, 1, None) == -1 and "smoothing" in _lib.last_error()
    assert lib.bp_xentropy_fwd(a, None, a, a, 4, 64, 64, 0.0, -100, -1, 1, None) == -1
    assert lib.bp_xentropy_bwd(a, a, a, a, a, 4, 64, 64, 32, 0.0, -100, -1, 1, None) == -1
    assert lib.bp_xentropy_bwd(a, a, a, a, a, 4, 64, 64, 64, 0.0, -100, 32, 1, None) == -1 and "total_classes" in _lib.last_error()
    del keep


def test_dropout_mask_restatement():
    """attention_dropout_mask == an independent numpy restatement of csrc/bp_common.cuh; drop fraction = thr / 256;
    deterministic in the seed; different per (batch, head)."""
    from backpacks_flash_attn_b200.flash_attn_interface import attention_dropout_mask, effective_dropout_p
    assert effective_dropout_p(0.0) == 0.0 and effective_dropout_p(0.1) == 26 / 256 and effective_dropout_p(1e-4) == 1 / 256
    assert effective_dropout_p(0.999) == 255 / 256
    M = 0xFFFFFFFF

    def mix32(x):
        x = x & M
        x ^= x >> 16
        x = (x * 0x7FEB352D) & M
        x ^= x >> 15
        x = (x * 0x846CA68B) & M
        return x ^ (x >> 16)

    seed, b, h, sq, sk, p = 0xDEADBEEF12345678, 2, 3, 200, 300, 0.17
    m = attention_dropout_mask(seed, b, h, sq, sk, p)
    assert m.shape == (b, h, sq, sk) and m.dtype == torch.bool
    thr = int(effective_dropout_p(p) * 256)
    for bh in (0, 4):
        base = mix32((seed & M) ^ mix32(((seed >> 32) + bh) & M))
        q, k = np.arange(sq, dtype=np.uint64), np.arange(sk, dtype=np.uint64)
        rw = mix32((base + q * 0x9E3779B1) & M)
        cw = mix32((((~np.uint64(base)) & M) + k * 0x85EBCA77) & M)
        ref = (((rw[:, None] ^ cw[None, :]) * 0x2C1B3C6D) & M) >= (thr << 24)
        assert (m[bh // h, bh % h].numpy() == ref).all()
    assert abs((1 - m.float().mean().item()) - thr / 256) < 5e-3
    assert torch.equal(m, attention_dropout_mask(seed, b, h, sq, sk, p))
    assert not torch.equal(m, attention_dropout_mask(seed + 1, b, h, sq, sk, p))
    assert not torch.equal(m[0, 0], m[0, 1])


def test_training_operators_refuse_cpu_tensors(lib):
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss, SoftmaxCrossEntropyLossFn
    from backpacks_flash_attn_b200.ops.fused_dense import bias_act_backward, linear_bias_act
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    qkv = torch.zeros(16, 3, 2, 64, dtype=torch.bfloat16, requires_grad=True)
    cu = torch.tensor([0, 16], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA"):
        flash_attn_unpadded_qkvpacked_func(qkv, cu, 16, 0.1, causal=True)
    x = torch.zeros(4, 64, dtype=torch.bfloat16, requires_grad=True)
    w = torch.ones(64, dtype=torch.bfloat16, requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        dropout_add_layer_norm(x, None, w, w, 0.0, 1e-5)
    with pytest.raises(RuntimeError, match="CUDA"):
        linear_bias_act(x, torch.zeros(8, 64, dtype=torch.bfloat16, requires_grad=True), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        bias_act_backward(x.detach(), x.detach(), "gelu_tanh", True)
    with pytest.raises(RuntimeError, match="CUDA"):
        sense_mix(torch.zeros(1, 8, 2, 2, 8, dtype=torch.bfloat16, requires_grad=True),
                  torch.zeros(1, 2, 8, 16, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="CUDA"):
        CrossEntropyLoss()(torch.zeros(4, 8), torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        SoftmaxCrossEntropyLossFn.apply(torch.zeros(4, 8), torch.zeros(4, dtype=torch.long))


def test_layer_norm_dropout_mask_restatement_and_argument_validation(lib):
    """The Python restatement of the LayerNorm dropout mask against a scalar re-derivation of bp_common.cuh's hash, and
    the argument checks of the two dropout entry points (before any CUDA call)."""
    from backpacks_flash_attn_b200 import _lib
    from backpacks_flash_attn_b200.flash_attn_interface import effective_dropout_p
    from backpacks_flash_attn_b200.ops.layer_norm import layer_norm_dropout_mask

    def mix(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16
        x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x846CA68B) & 0xFFFFFFFF
        return x ^ (x >> 16)

    seed, rows, cols, p = (123456789 << 20) + 17, 37, 24, 0.3
    thr = int(effective_dropout_p(p) * 256)
    base = mix((seed & 0xFFFFFFFF) ^ mix(((seed >> 32) & 0xFFFFFFFF) + 0x4C4E))
    m = layer_norm_dropout_mask(seed, rows, cols, p)
    for r in range(rows):
        rw = mix(base + r * 0x9E3779B1)
        for c in range(cols):
            cw = mix((~base & 0xFFFFFFFF) + c * 0x85EBCA77)
            assert bool(m[r, c]) == ((((rw ^ cw) * 0x2C1B3C6D) & 0xFFFFFFFF) >= (thr << 24))
    big = layer_norm_dropout_mask(seed, 4096, 768, p)
    assert abs((1 - big.float().mean().item()) - thr / 256) < 2e-3
    assert not torch.equal(big, layer_norm_dropout_mask(seed + 1, 4096, 768, p))
    buf = ctypes.create_string_buffer(256)
    a = (ctypes.addressof(buf) + 15) // 16 * 16
    st = lib.bp_ln_residual_fwd_dropout(a, a, a, a, a, a, None, None, 4, 8, 1e-5, 1, 2, 1, 1.0, 5, None)
    assert st == -1 and "dropout_p" in _lib.last_error()
    st = lib.bp_ln_residual_fwd_dropout(a, a, a, a, a, a, None, None, 4, 8, 1e-5, 1, 2, 1, -0.1, 5, None)
    assert st == -1 and "dropout_p" in _lib.last_error()
    st = lib.bp_ln_residual_bwd_dropout(a, a, a, a, None, None, a, a, a, a, a, 1 << 30, 4, 8, 1e-5, 1, 2, 1, 1.5, 5, None)
    assert st == -1 and "dropout_p" in _lib.last_error()
    st = lib.bp_ln_residual_fwd_dropout(a, a, a, a, a, a, None, None, 4, 12, 1e-5, 1, 2, 1, 0.1, 5, None)
    assert st == -1 and "multiple of 8" in _lib.last_error()
