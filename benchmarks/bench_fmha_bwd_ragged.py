"""bp_fmha_bwd on a ragged batch (32 sequences of 32..1024 tokens): the case that separates the dynamic ticket
scheduler of the backward kernels from a static per-CTA schedule."""
import torch
from backpacks_flash_attn_b200 import flash_attn_interface as F

torch.manual_seed(0)
lens = [1024, 64, 1024, 128, 900, 32, 1024, 256] * 4
h, d = 12, 64
qkv = torch.randn(sum(lens), 3, h, d, device="cuda").bfloat16()
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
out, lse = F._flash_attn_forward(qkv[:, 0], qkv[:, 1], qkv[:, 2], torch.empty_like(qkv[:, 0]), cu, cu, 1024, 1024, d ** -0.5, True)
g, dqkv = torch.randn_like(out), torch.empty_like(qkv)
f = lambda: F._flash_attn_backward(g, qkv[:, 0], qkv[:, 1], qkv[:, 2], out, lse, dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], cu, cu,
                                   1024, 1024, d ** -0.5, True)
for _ in range(3):
    f()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    f()
e.record()
torch.cuda.synchronize()
print('{"kernel": "fmha_bwd, ragged batch: 32 sequences of 32..1024 tokens, h12 d64 causal bf16", "us": %.1f}' % (a.elapsed_time(e) / 20 * 1e3))
