"""Debug: dump the timeline of CTA 0 of the fmha kernel (timestamps of pipeline events per warp role)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backpacks_flash_attn_b200 import _lib
from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func

b, s, h, d = 32, 1024, 12, 64
qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
for _ in range(3):
    flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
R, N = 7, 512
buf = torch.zeros(R * N * 2, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.bp_debug_set_fmha_trace.argtypes = [ctypes.c_void_p]
lib.bp_debug_set_fmha_trace.restype = None
lib.bp_debug_set_fmha_trace(buf.data_ptr())
flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
torch.cuda.synchronize()
lib.bp_debug_set_fmha_trace(None)
t = buf.cpu().view(R, N, 2)
names = ["prod", "mma0", "mma1", "sm00", "sm01", "sm10", "sm11"]
t0 = min(int(t[r, 0, 1]) for r in range(R) if int(t[r, 0, 1]) > 0)
events = []
for r in range(R):
    for i in range(N):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0:
            break
        events.append((clk - t0, names[r], tag >> 32, tag & 0xffffffff))
events.sort()
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for e in events[:limit]:
    print(f"{e[0]:8d} {e[1]:5s} ev{e[2]} #{e[3]}")
print("total events", len(events), "last", events[-1])
