"""out_proj / fc2 shapes: cuBLAS F.linear vs this library's GEMM (16-bit out) vs the residual-epilogue GEMM."""
import sys, torch
sys.path.insert(0, '/root/repo')
from benchmarks.bench_kernels import time_fn
from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act, linear_bias_residual_
from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm, layer_norm_from_residual
m = 65536
for (n, k) in ((768, 768), (768, 3072)):
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    b = torch.randn(n, device="cuda").bfloat16()
    res = torch.randn(m, n, device="cuda")
    g = torch.ones(n, device="cuda").bfloat16(); be = torch.zeros(n, device="cuda").bfloat16()
    with torch.no_grad():
        for rep in range(2):
            t_cb, _ = time_fn(lambda i: torch.nn.functional.linear(x, w, b), 1, 8)
            t_pl, _ = time_fn(lambda i: linear_bias_act(x, w, b, "none"), 1, 8)
            t_rs, _ = time_fn(lambda i: linear_bias_residual_(x, w, b, res), 1, 8)
            y = torch.nn.functional.linear(x, w, b)
            t_ln, _ = time_fn(lambda i: dropout_add_layer_norm(y, res, g, be, 0.0, 1e-5, prenorm=True), 1, 8)
            t_lf, _ = time_fn(lambda i: layer_norm_from_residual(res, g, be, 1e-5), 1, 8)
            print(f"n={n} k={k}: cublas {t_cb*1e6:6.1f} | ours16 {t_pl*1e6:6.1f} | ours+res {t_rs*1e6:6.1f} | LN(add) {t_ln*1e6:6.1f} | LN(fp32) {t_lf*1e6:6.1f}"
                  f" || two-kernel {1e6*(t_cb+t_ln):6.1f} vs fused {1e6*(t_rs+t_lf):6.1f}", flush=True)
