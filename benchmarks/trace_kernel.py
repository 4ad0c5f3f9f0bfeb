"""Debug: dump the per-warp-role timeline of CTA 0 of a kernel (library built with BP_EXTRA_NVCC_FLAGS=-DBP_TRACE).

    python benchmarks/trace_kernel.py fmha|sense|sense_table [max_lines]
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backpacks_flash_attn_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "fmha"
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
R, N = 8, 512
buf = torch.zeros(R * N * 2 + 2 * 148 + 4 * 4096, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.bp_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.bp_debug_set_trace.restype = None

if which == "fmha":
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    b, s, h, d = 32, 1024, 12, 64
    qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    run = lambda: flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
    names = ["prod", "mma0", "mma1", "sm0", "sm1", "-", "-", "-"]
elif which in ("bwd_dkdv", "bwd_dq"):
    os.environ["BP_TRACE_BWD"] = which[4:]
    from backpacks_flash_attn_b200.flash_attn_interface import _flash_attn_backward, _flash_attn_forward
    b, s, h, d = 32, 1024, 12, 64
    qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    out, lse = _flash_attn_forward(qkv[:, 0], qkv[:, 1], qkv[:, 2], torch.empty_like(qkv[:, 0]), cu, cu, s, s, d ** -0.5, True)
    g, dqkv = torch.randn_like(out), torch.empty_like(qkv)
    run = lambda: _flash_attn_backward(g, qkv[:, 0], qkv[:, 1], qkv[:, 2], out, lse, dqkv[:, 0], dqkv[:, 1], dqkv[:, 2],
                                       cu, cu, s, s, d ** -0.5, True)
    names = ["prod", "mma", "sm0", "sm1", "-", "-", "-", "-"]
else:
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix, sense_mix_table
    b, s, nv, d = 64, 1024, 16, 768
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
    if which == "sense_table":
        table = torch.randn(50264, nv, d, device="cuda").bfloat16()
        ids = torch.randint(0, 50257, (b, s), device="cuda")
        run = lambda: sense_mix_table(qk, table, ids)
    else:
        content = torch.randn(b, s, nv, d, device="cuda").bfloat16().transpose(1, 2)
        run = lambda: sense_mix(qk, content)
    names = ["prodC", "issue", "sm0", "sm1", "-", "-", "-", "-"]
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.bp_debug_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.bp_debug_set_trace(None)
cta = buf.cpu()[R * N * 2:R * N * 2 + 2 * 148].view(148, 2)
t = buf.cpu()[:R * N * 2].view(R, N, 2)
starts = [int(t[r, 0, 1]) for r in range(R) if int(t[r, 0, 1]) > 0]
t0 = min(starts)
events = []
for r in range(R):
    for i in range(N):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0:
            break
        events.append((clk - t0, names[r], tag >> 32, tag & 0xffffffff))
events.sort()
for e in events[:limit]:
    print(f"{e[0]:8d} {e[1]:5s} ev{e[2]} #{e[3]}")
if int(cta[:, 0].max()) > 0:
    t0 = int(cta[:, 0][cta[:, 0] > 0].min())
    ends = sorted(int(e) - t0 for e in cta[:, 1] if int(e) > 0)
    starts = sorted(int(b) - t0 for b in cta[:, 0] if int(b) > 0)
    print("CTA start ns: min %d max %d; end ns: min %d p25 %d median %d p75 %d max %d" % (
        starts[0], starts[-1], ends[0], ends[len(ends) // 4], ends[len(ends) // 2], ends[3 * len(ends) // 4], ends[-1]))
print("total events", len(events))

if which.startswith("bwd"):
    # per-CTA schedule: how much of each SM's time is covered by 0 / 1 / 2 resident CTAs, and the CTA durations
    rec = buf.cpu()[R * N * 2 + 2 * 148:].view(4096, 4)
    rec = rec[rec[:, 0] > 0]
    t0 = int(rec[:, 0].min())
    span = int(rec[:, 1].max()) - t0
    dur = (rec[:, 1] - rec[:, 0]).float()
    steps = rec[:, 3].float()
    print(f"{len(rec)} CTAs, kernel span {span} ns; CTA duration ns: mean {dur.mean():.0f} min {dur.min():.0f} max {dur.max():.0f}")
    for k in sorted(set(steps.tolist())):
        m = steps == k
        print(f"  steps {int(k):3d}: {int(m.sum()):5d} CTAs, mean duration {dur[m].mean():.0f} ns, min {dur[m].min():.0f}")
    busy = 0
    for sm in sorted(set(rec[:, 2].tolist())):
        r = rec[rec[:, 2] == sm]
        busy += int((r[:, 1] - r[:, 0]).sum())
    nsm = len(set(rec[:, 2].tolist()))
    print(f"SMs used {nsm}; mean resident CTAs per SM over the span: {busy / nsm / span:.2f}")
