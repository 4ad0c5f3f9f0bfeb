"""Top CUDA kernels of one Backpack-Small training step (torch profiler): where the time outside this library goes.

    python benchmarks/profile_training_step.py [batch]                 # torch profiler table of one step
    python benchmarks/profile_training_step.py 64 --plain 2             # two bare steps (ncu launch list)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import ProfilerActivity, profile
from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
from backpacks_flash_attn_b200.utils.weights import name_seeded_

b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = flash_config(n_embd=768, n_head=12, n_layer=12, n_positions=1024, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
model = name_seeded_(BackpackLMHeadModel(cfg)).to("cuda", torch.bfloat16).train()
ids = torch.randint(0, 50257, (b, 1024), device="cuda")
labels = torch.cat([ids[:, 1:], torch.full_like(ids[:, :1], -100)], dim=1).reshape(-1)
ce = CrossEntropyLoss(inplace_backward=True)


def step():
    model.zero_grad(set_to_none=True)
    logits = model(ids).logits
    ce(logits.view(-1, logits.shape[-1]), labels).backward()


if len(sys.argv) > 2 and sys.argv[2] == "--plain":      # N bare steps, for an ncu launch list of the training step
    for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
        step()
    torch.cuda.synchronize()
    raise SystemExit(0)
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"total CUDA time {total / 1e3:.1f} ms")
for e in rows[:28]:
    print(f"{e.device_time_total / 1e3:8.2f} ms {100 * e.device_time_total / total:5.1f} %  x{e.count:4d}  {e.key[:110]}")
