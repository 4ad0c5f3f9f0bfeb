"""Scratch: repeat a multi-item backward and report which gradient tensor is non-finite / non-reproducible."""
import sys
import torch
from backpacks_flash_attn_b200 import flash_attn_interface as F

b, h, s, d = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (37, 8, 384, 64)))
qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
out, lse = F._flash_attn_forward(qkv[:, 0], qkv[:, 1], qkv[:, 2], torch.empty_like(qkv[:, 0]), cu, cu, s, s, d ** -0.5, True)
g = torch.randn_like(out)
first = None
for it in range(20):
    dqkv = torch.full_like(qkv, float("nan"))
    F._flash_attn_backward(g, qkv[:, 0], qkv[:, 1], qkv[:, 2], out, lse, dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], cu, cu, s, s, d ** -0.5, True)
    torch.cuda.synchronize()
    msg = []
    for i, n in enumerate(("dQ", "dK", "dV")):
        t = dqkv[:, i].float()
        bad = ~torch.isfinite(t)
        if bad.any():
            rows = bad.any(-1).any(-1).nonzero()[:, 0]
            msg.append(f"{n}: {int(bad.sum())} non-finite, rows {rows[:6].tolist()} (seq pos {[int(r) % s for r in rows[:6]]}, batch {[int(r) // s for r in rows[:6]]})")
        if first is not None and not torch.equal(dqkv[:, i], first[:, i]):
            diff = (dqkv[:, i].float() - first[:, i].float()).abs()
            rows = (diff.amax((-1, -2)) > 0).nonzero()[:, 0]
            msg.append(f"{n}: differs from run 0 in {len(rows)} rows, e.g. seq pos {[int(r) % s for r in rows[:6]]}")
    if first is None:
        first = dqkv.clone()
    print(it, "; ".join(msg) if msg else "ok", flush=True)
