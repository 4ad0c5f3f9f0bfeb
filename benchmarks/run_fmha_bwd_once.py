"""One backward of config 2 (for ncu launch lists / captures of the three bp_fmha_bwd kernels)."""
import sys
import torch
from backpacks_flash_attn_b200 import flash_attn_interface as F

b, s, h, d = 32, 1024, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 64
qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16().requires_grad_(True)
cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
out = F.flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
g = torch.randn_like(out)
for _ in range(3):
    torch.autograd.grad(out, qkv, g, retain_graph=True)
torch.cuda.synchronize()
