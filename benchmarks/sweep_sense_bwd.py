"""Sense-mix backward over sequence length x number of senses (the BASELINE config-5 grid of the forward sweep), hand-derived
backward (batched GEMMs + bp_sense_softmax_bwd) next to autograd through the reference's eager composition.

    python benchmarks/sweep_sense_bwd.py > profiles/r02_sweep_sense_bwd.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.bench_kernels import time_fn  # noqa: E402
from backpacks_flash_attn_b200.ops.sense_mix import _sense_mix_backward, _sense_mix_backward_eager  # noqa: E402

d = 768
for s in (512, 1024, 2048, 4096):
    for nv in (4, 16, 64):
        b = max(1, 16384 // s)                       # 16 k tokens per call
        qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
        content = (torch.randn(b, s, nv, d, device="cuda") * 0.5).bfloat16().transpose(1, 2)
        dout = torch.randn(b, s, d, device="cuda").bfloat16()
        scale = (d // nv) ** -0.5
        t, _ = time_fn(lambda i: _sense_mix_backward(qk, content, dout, scale, True, True), 1, 5, warmup=2, inner=1)
        te, _ = time_fn(lambda i: _sense_mix_backward_eager(qk, content, dout, scale, True, True), 1, 3, warmup=1, inner=1)
        a = _sense_mix_backward(qk, content, dout, scale, True, True)
        e = _sense_mix_backward_eager(qk, content, dout, scale, True, True)
        rel = [float((x.float() - y.float()).abs().max() / y.float().abs().max()) for x, y in zip(a, e)]
        print(json.dumps({"op": "sense_mix_bwd", "seq": s, "senses": nv, "batch": b, "d": d, "ms": t * 1e3, "ms_eager_autograd": te * 1e3,
                          "speedup": te / t, "max_rel_diff_vs_eager_dqk_dcontent": rel}), flush=True)
        del qk, content, dout, a, e
        torch.cuda.empty_cache()
