"""One launch of this library's GEMM and of cuBLAS per model shape, for an ncu DRAM-traffic comparison:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
        --clock-control none --csv --log-file out.csv python benchmarks/gemm_traffic.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act  # noqa: E402

m = 65536
for name, n, k in [("Wqkv", 2304, 768), ("fc2", 768, 3072), ("final_mlp.fc2", 12288, 3072), ("lm_head", 50264, 768)]:
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    for _ in range(2):
        linear_bias_act(x, w, None, "none")
        torch.nn.functional.linear(x, w)
    torch.cuda.synchronize()
    del x, w
