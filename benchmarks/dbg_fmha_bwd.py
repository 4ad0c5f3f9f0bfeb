"""Scratch check of bp_fmha_bwd: per-tensor errors against the fp32 oracle for a few shapes, then a timing."""
import sys
import torch
from oracle import backpack_oracle as O
from backpacks_flash_attn_b200 import flash_attn_interface as F


def run(b, s, h, d, causal, dtype=torch.bfloat16):
    g0 = torch.Generator(device="cuda").manual_seed(s + d)
    qkv = torch.randn(b, s, 3, h, d, device="cuda", generator=g0).to(dtype)
    g = torch.randn(b, s, h, d, device="cuda", generator=g0).to(dtype)
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    x = qkv.reshape(b * s, 3, h, d).clone().requires_grad_(True)
    out = F.flash_attn_unpadded_qkvpacked_func(x, cu, s, 0.0, causal=causal)
    dqkv, = torch.autograd.grad(out, x, g.reshape(b * s, h, d))
    torch.cuda.synchronize()
    dqkv = dqkv.reshape(b, s, 3, h, d).float()
    xr = qkv.float().requires_grad_(True)
    ref_out, _ = O.attention_fp32_ref(*xr.unbind(2), None, causal)
    ref, = torch.autograd.grad(ref_out, xr, g.float())
    y = qkv.clone().requires_grad_(True)
    pt, = torch.autograd.grad(O.self_attention_eager(y, None, causal), y, g)
    msg = f"b{b} s{s} h{h} d{d} causal={int(causal)}:"
    for i, n in enumerate(("dQ", "dK", "dV")):
        e = (dqkv[:, :, i] - ref[:, :, i]).abs().max().item()
        ep = (pt[:, :, i].float() - ref[:, :, i]).abs().max().item()
        msg += f"  {n} err {e:.3e} (eager {ep:.3e}, |ref| {ref[:, :, i].abs().max().item():.2f})"
    print(msg, flush=True)


for cfg in [(1, 128, 1, 64, False), (1, 128, 1, 64, True), (2, 256, 2, 64, True), (2, 200, 2, 64, True),
            (2, 1024, 4, 64, True), (2, 1024, 4, 64, False), (2, 512, 2, 128, True), (2, 333, 2, 40, True),
            (2, 97, 2, 80, False)]:
    run(*cfg)

if "--time" in sys.argv:
    b, s, h, d = 32, 1024, 12, 64
    qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16().requires_grad_(True)
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    out = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
    g = torch.randn_like(out)
    for _ in range(3):
        torch.autograd.grad(out, qkv, g, retain_graph=True)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        torch.autograd.grad(out, qkv, g, retain_graph=True)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    fl = 2.5 * 4 * b * h * s * s * d / 2
    print(f"config 2 backward: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s (2.5x forward FLOPs convention)")
