"""Graph-level A/B of the linear backend, one GEMM shape at a time: the Backpack-Small forward (config 3) replayed as
one CUDA graph with every plain linear on cuBLAS except ONE shape on this library's GEMM (and the two extremes).
Per-kernel event timings do not decide this: the step runs under the power cap, so a kernel's effect on the clocks of
its neighbours counts.  Prints one JSON line per policy (median of `reps` timed blocks of `steps` replays)."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config  # noqa: E402
from backpacks_flash_attn_b200.ops import fused_dense as FD  # noqa: E402
from backpacks_flash_attn_b200.utils.graph import GraphedForward  # noqa: E402
from backpacks_flash_attn_b200.utils.weights import name_seeded_  # noqa: E402

steps, reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30, 3
cfg = flash_config(n_embd=768, n_head=12, n_layer=12, n_positions=1024)
model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)
ids = torch.randint(0, 50257, (64, 1024), generator=torch.Generator().manual_seed(1234)).cuda()
d, V = 768, cfg.vocab_size
shapes = {"Wqkv": (3 * d, d), "out_proj": (d, d), "fc2": (d, 4 * d), "ctx_Wqkv": (2 * d, d), "final_fc2": (16 * d, 4 * d),
          "lm_head": (V, d)}
policies = {"library": "library", "own": "own"}
for name, key in shapes.items():
    policies["own_only_" + name] = {k: ("own" if k == key else "library") for k in shapes.values()}
policies["auto"] = "auto"
policies["library_again"] = "library"
results = {}
with torch.inference_mode():
    for pname, pol in policies.items():
        FD.set_linear_backend(pol)
        for _ in range(2):
            model(ids)
        g = GraphedForward(model, ids)
        for _ in range(3):
            g()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                g()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / steps)
        results[pname] = statistics.median(ts)
        print(json.dumps({"policy": pname, "ms_per_step": results[pname], "all": ts}), flush=True)
        del g
        torch.cuda.empty_cache()
