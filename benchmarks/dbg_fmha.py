import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
from oracle import backpack_oracle as O
import os as _os
CASES = [(1, 128, 1, 64, True), (1, 256, 1, 64, True), (2, 1024, 4, 64, True), (2, 200, 2, 128, False), (32, 1024, 12, 64, True)]
for (b, s, h, d, causal) in CASES[:int(_os.environ.get('NCASES', '99'))]:
    qkv = torch.randn(b, s, 3, h, d, device="cuda").bfloat16()
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    for rep in range(int(_os.environ.get("REPS", "3"))):
        out = flash_attn_unpadded_qkvpacked_func(qkv.reshape(b * s, 3, h, d), cu, s, 0.0, causal=causal)
        torch.cuda.synchronize()
    if b * h <= 16:
        ref, _ = O.attention_fp32_ref(*qkv.unbind(2), None, causal)
        print((b, s, h, d, causal), "ok, max err", (out.reshape(b, s, h, d).float() - ref).abs().max().item(), flush=True)
    else:
        print((b, s, h, d, causal), "ran", flush=True)
