"""BASELINE config 5 sweep: fused attention over seq {512,1024,2048,4096} x head_dim {64,128} (b*s = 32768 tokens,
h = 768/d, causal, bf16) and the fused sense-mix over the same seq x k {4,16,64} senses (d = 768, b*s = 16384).
Prints one JSON line per point (kernel time = median of back-to-back launches bracketed by CUDA events)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.bench_kernels import peaks, time_fn  # noqa: E402


def main():
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    hbm, tf_burst, tf_sus, src = peaks()
    for d in (64, 128):
        for s in (512, 1024, 2048, 4096):
            b, h = 32768 // s, 768 // d
            qkvs = [torch.randn(b * s, 3, h, d, device="cuda").bfloat16() for _ in range(3)]
            cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
            t, _ = time_fn(lambda i: flash_attn_unpadded_qkvpacked_func(qkvs[i], cu, s, 0.0, causal=True), 3, 10)
            flops, nbytes = 4 * b * h * s * s * d / 2, 4 * b * s * h * d * 2 + 4 * b * h * s
            print(json.dumps({"op": "fmha_fwd", "seq": s, "head_dim": d, "batch": b, "heads": h, "us": t * 1e6,
                              "tflops": flops / t / 1e12, "frac_tensor_burst": flops / t / 1e12 / tf_burst,
                              "gbs": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm}))
            del qkvs
    for k in (4, 16, 64):
        for s in (512, 1024, 2048, 4096):
            b, d = 16384 // s, 768
            qk = torch.randn(b, s, 2, k, d // k, device="cuda").bfloat16()
            content = torch.randn(b, s, k, d, device="cuda").bfloat16().transpose(1, 2)
            t, _ = time_fn(lambda i: sense_mix(qk, content), 1, 6, inner=3)
            flops = b * s * s * d * (1 + k)
            nbytes = (2 * b * s * d + k * b * s * d + b * s * d) * 2
            print(json.dumps({"op": "sense_mix", "seq": s, "senses": k, "dk": d // k, "batch": b, "us": t * 1e6,
                              "tflops": flops / t / 1e12, "frac_tensor_burst": flops / t / 1e12 / tf_burst,
                              "gbs": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm}))
            del qk, content


if __name__ == "__main__":
    main()
