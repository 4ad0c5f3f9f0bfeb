"""Hardware-contract probe: the UMMA shared-memory / instruction descriptor conventions and the TMA
128B-swizzle layout that every tcgen05 kernel in csrc/ relies on (tests/gpu_probe/umma_probe.cu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_umma_descriptor_probe():
    exe = os.path.join(ROOT, "tests", "gpu_probe", "umma_probe.bin")
    if not os.path.exists(exe):
        from backpacks_flash_attn_b200 import build
        exe = build.build_probe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PROBE OK" in r.stdout
