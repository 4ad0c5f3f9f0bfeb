"""CPU checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every
symbol include/backpack_b200.h declares; the Python binding fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "backpack_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from backpacks_flash_attn_b200 import build
    return build.build_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("bp_fmha_fwd", "bp_sense_lse_fwd", "bp_sense_mix_fwd", "bp_ln_residual_fwd",
              "bp_linear_bias_act_fwd", "bp_linear_bias_residual_fwd", "bp_ln_fwd", "bp_rotary_qk_inplace", "bp_last_error", "bp_abi_version",
              "bp_check_device", "bp_sense_mix_table_fwd", "bp_lm_head_stats_fwd", "bp_decode_attn_fwd",
              "bp_sense_mix_decode_fwd"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/backpack_b200.h but not exported"
    assert lib.bp_abi_version() == 1


def test_binding_signatures_cover_the_header(lib_path):
    from backpacks_flash_attn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_invalid_arguments_are_rejected_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call, mirroring fmha_api.cpp:215-252."""
    from backpacks_flash_attn_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    a = ctypes.addressof(buf)
    a = (a + 15) // 16 * 16
    # head dim not a multiple of 8
    st = lib.bp_fmha_fwd(a, a, a, a, a, a, a, 1, 1, 12, 16, 16, 16, 16, 8, 8, 8, 8, 8, 8, 8, 8, 16, 1.0, 1, 1, None)
    assert st == -1 and "multiple of 8" in _lib.last_error()
    # unsupported dtype
    st = lib.bp_fmha_fwd(a, a, a, a, a, a, a, 1, 1, 64, 16, 16, 16, 16, 64, 64, 64, 64, 64, 64, 64, 64, 16, 1.0, 1, 2, None)
    assert st == -1 and "fp16 and bf16" in _lib.last_error()
    # null pointer
    st = lib.bp_ln_residual_fwd(None, None, a, a, a, None, None, None, 4, 64, 1e-5, 1, 2, 1, None)
    assert st == -1
    st = lib.bp_ln_residual_fwd(a, None, a, a, a, None, None, None, 4, 60, 1e-5, 1, 2, 1, None)
    assert st == -1 and "multiple of 8" in _lib.last_error()
    st = lib.bp_rotary_qk_inplace(a, a, a, None, None, 1, 4, 1, 64, 65, 1, None)
    assert st == -1
    # residual-epilogue GEMM: fewer rows than one CTA-pair tile, odd n, null residual
    st = lib.bp_linear_bias_residual_fwd(a, a, None, a, 64, 64, 64, 1, None)
    assert st != 0 and "m >= 256" in _lib.last_error()
    st = lib.bp_linear_bias_residual_fwd(a, a, None, a, 512, 60, 64, 1, None)
    assert st == -1 and "multiples of 8" in _lib.last_error()
    st = lib.bp_linear_bias_residual_fwd(a, a, None, None, 512, 64, 64, 1, None)
    assert st == -1
    # LayerNorm of the fp32 residual: unsupported dtype combination, bad width
    st = lib.bp_ln_fwd(a, a, a, a, None, None, 4, 64, 1e-5, 2, 1, 0, None)      # f32 -> bf16 with f16 weights
    assert st != 0 and "not built" in _lib.last_error()
    st = lib.bp_ln_fwd(a, a, a, a, None, None, 4, 60, 1e-5, 2, 1, 1, None)
    assert st == -1 and "multiple of 8" in _lib.last_error()


def test_operators_refuse_cpu_tensors(lib_path):
    """No CPU fallback: the product path raises on host tensors."""
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    qkv = torch.zeros(16, 3, 2, 64, dtype=torch.bfloat16)
    cu = torch.tensor([0, 16], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA"):
        flash_attn_unpadded_qkvpacked_func(qkv, cu, 16, 0.0, causal=True)


def test_graphed_forward_refuses_cpu_inputs():
    from backpacks_flash_attn_b200.utils.graph import GraphedForward
    with torch.inference_mode():
        with pytest.raises(RuntimeError, match="CUDA"):
            GraphedForward(torch.nn.Identity(), torch.zeros(1, 8, dtype=torch.long))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from backpacks_flash_attn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()
