import sys, torch
sys.path.insert(0, '/root/repo')
from benchmarks.bench_kernels import time_fn
from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
m, n, k = 65536, 3072, 768
x = torch.randn(m, k, device="cuda").bfloat16()
w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
b = torch.randn(n, device="cuda").bfloat16()
for name, fn in [("gelu+bias", lambda i: linear_bias_act(x, w, b, "gelu_tanh")),
                 ("bias only", lambda i: linear_bias_act(x, w, b, "none")),
                 ("plain", lambda i: linear_bias_act(x, w, None, "none")),
                 ("cublas", lambda i: torch.nn.functional.linear(x, w, b))]:
    t, _ = time_fn(fn, 1, 10)
    print(f"{name:10s} {t*1e6:8.1f} us  {2*m*n*k/t/1e12:7.1f} TF")
# K sweep: mainloop share
for kk in (256, 768, 3072):
    x2 = torch.randn(m, kk, device="cuda").bfloat16(); w2 = (torch.randn(n, kk, device="cuda") * kk ** -0.5).bfloat16()
    t, _ = time_fn(lambda i: linear_bias_act(x2, w2, None, "none"), 1, 6)
    t2, _ = time_fn(lambda i: torch.nn.functional.linear(x2, w2), 1, 6)
    print(f"k={kk}: ours {t*1e6:8.1f} us {2*m*n*kk/t/1e12:7.1f} TF | cublas {t2*1e6:8.1f} us {2*m*n*kk/t2/1e12:7.1f} TF")
