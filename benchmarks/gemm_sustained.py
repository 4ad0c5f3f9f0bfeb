"""Sustained (power-capped) throughput of this library's GEMM next to cuBLAS, per model shape.

The stand-alone timings of benchmarks/bench_kernels.py are bursts of ~20 launches at boost clocks; inside the 25 ms
Backpack step the GPU sits at its 1000 W cap (SM clock ~1.4 GHz) and what counts is throughput at that operating point.
Each shape is run back to back for `--seconds` per backend while SM clock and board power are sampled every 50 ms.

    python benchmarks/gemm_sustained.py [--seconds 2.0]
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act  # noqa: E402


class Sampler:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.clk, self.pw = [], []
        self._stop = threading.Event()

    def __enter__(self):
        def run():
            while not self._stop.is_set():
                self.clk.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self._stop.wait(0.05)
        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self.t.join()
        return False

    def tail_median(self, xs):
        xs = sorted(xs[len(xs) // 2:])            # second half of the run: steady state
        return xs[len(xs) // 2] if xs else None


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n, chunk = 0, 20
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Sampler() as s:
        t0 = time.perf_counter()
        a.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(chunk):
                fn()
            n += chunk
            torch.cuda.synchronize()          # keep the launch queue bounded; 20 launches >> sync latency
        e.record()
        e.synchronize()
    return a.elapsed_time(e) / n, s.tail_median(s.clk), s.tail_median(s.pw)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    args = ap.parse_args()
    m = 65536
    shapes = [("Wqkv", 2304, 768, True), ("out_proj", 768, 768, True), ("fc2", 768, 3072, True),
              ("final_mlp.fc2", 12288, 3072, True), ("lm_head", 50264, 768, False)]
    for name, n, k, has_bias in shapes:
        x = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
        bias = torch.randn(n, device="cuda").bfloat16() if has_bias else None
        flops = 2.0 * m * n * k
        rec = {"gemm": name, "m": m, "n": n, "k": k, "seconds_per_backend": args.seconds}
        for backend, fn in (("own", lambda: linear_bias_act(x, w, bias, "none")),
                            ("cublas", lambda: torch.nn.functional.linear(x, w, bias))):
            ms, clk, pw = sustained(fn, args.seconds)
            rec[backend] = {"us": ms * 1e3, "tflops": flops / ms / 1e9, "sm_mhz": clk, "power_w": pw,
                            "gflop_per_joule": flops / 1e9 / (ms / 1e3 * pw) if pw else None}
        rec["own_over_cublas_time"] = rec["own"]["us"] / rec["cublas"]["us"]
        print(json.dumps(rec), flush=True)
        del x, w, bias
