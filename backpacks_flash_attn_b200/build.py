"""Build libbackpack_b200.so (sm_100a only) in-tree with nvcc.

    python -m backpacks_flash_attn_b200.build [--force]

The shared library is the C-ABI of include/backpack_b200.h.  It is built next to this file so it travels
with the repository snapshot to the GPU box (a JIT cache would not).  cudart is linked statically and the
driver API is resolved at run time, so the .so loads (and its symbols can be checked) without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# BP_BUILD_TAG=trace builds a side-by-side debug variant (build_trace/, libbackpack_b200_trace.so) that the
# binding loads when BP_LIB_TAG=trace; the production library is never overwritten by a debug build.
_TAG = os.environ.get("BP_BUILD_TAG", "")
BUILD_DIR = os.path.join(PKG_DIR, "build" + (f"_{_TAG}" if _TAG else ""))
LIB_PATH = os.path.join(PKG_DIR, "libbackpack_b200" + (f"_{_TAG}" if _TAG else "") + ".so")
ROOT = os.path.dirname(PKG_DIR)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libbackpack_b200.so cannot be built")
    return exe


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> list[str]:
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hdr + [os.path.join(ROOT, "include", "backpack_b200.h")]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, obj: str) -> str:
    extra = os.environ.get("BP_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DBP_TRACE for timeline debugging
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = sources()
    objs = [os.path.join(BUILD_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    todo = [(s, o) for s, o in zip(srcs, objs) if force or _stale(o, [s] + _deps())]
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            logs = list(ex.map(lambda so: _compile(*so), todo))
        with open(os.path.join(BUILD_DIR, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    if todo or force or _stale(LIB_PATH, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


def build_probe(force: bool = False) -> str:
    """tests/gpu_probe/umma_probe: stand-alone hardware-contract check of the UMMA/TMA descriptors."""
    src = os.path.join(ROOT, "tests", "gpu_probe", "umma_probe.cu")
    out = os.path.join(ROOT, "tests", "gpu_probe", "umma_probe.bin")
    host = os.path.join(CSRC, "bp_host.cu")
    if force or _stale(out, [src, host] + _deps()):
        flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
        r = subprocess.run([_nvcc(), *flags, src, host, "-o", out], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"probe build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_library(force=force, verbose="-v" in sys.argv))
    print(build_probe(force=force))
