import sys
sys.path.insert(0, '/root/repo')
import benchmarks.sweep as sw
import torch, json
from benchmarks.bench_kernels import peaks, time_fn
from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
hbm, tf_burst, tf_sus, src = peaks()
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
                          "gbs": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm}), flush=True)
