"""Run ONE kernel of the library a few times at its BASELINE shape (for `ncu -k regex:...` captures and quick A/B timing).

    python benchmarks/run_one.py fmha|fmha128|sense|sense_table|lse|gemm|ln|dec_attn|dec_sense|sense_softmax_bwd [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1] if len(sys.argv) > 1 else "fmha"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3

if which in ("fmha", "fmha128"):
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    d = 64 if which == "fmha" else 128
    b, s, h = 32, 1024, 768 // d
    qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    run = lambda: flash_attn_unpadded_qkvpacked_func(qkv, cu, s, 0.0, causal=True)
elif which in ("sense", "sense_table", "lse"):
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
elif which == "dec_attn":
    from backpacks_flash_attn_b200.ops.decode import decode_attention
    b, s, h, d = 64, 1024, 12, 64
    cache = torch.randn(b, s, 2, h, d, device="cuda").bfloat16()
    q = torch.randn(b, 1, h, d, device="cuda").bfloat16()
    run = lambda: decode_attention(q, cache, s)
elif which == "dec_sense":
    from backpacks_flash_attn_b200.ops.decode import sense_mix_decode
    b, s, nv, d = 64, 1024, 16, 768
    table = torch.randn(50264, nv, d, device="cuda").bfloat16()
    kc = torch.randn(b, s, nv, d // nv, device="cuda").bfloat16()
    ids = torch.randint(0, 50257, (b, s), device="cuda")
    q = torch.randn(b, nv, d // nv, device="cuda").bfloat16()
    run = lambda: sense_mix_decode(q, kc, ids, table, s)
elif which == "sense_softmax_bwd":
    from backpacks_flash_attn_b200 import _lib
    b, s, nv = 64, 1024, 16
    S = torch.randn(nv, b, s, s, device="cuda").bfloat16()
    dA = torch.randn(nv, b, s, s, device="cuda").bfloat16()
    lib = _lib.load()
    run = lambda: _lib.check(lib.bp_sense_softmax_bwd(S.data_ptr(), dA.data_ptr(), nv * b * s, s, 48 ** -0.5, 1,
                                                      torch.cuda.current_stream().cuda_stream), "bp_sense_softmax_bwd")
elif which == "gemm":
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
    x = torch.randn(65536, 768, device="cuda").bfloat16()
    w = (torch.randn(3072, 768, device="cuda") * 768 ** -0.5).bfloat16()
    bias = torch.randn(3072, device="cuda").bfloat16()
    run = lambda: linear_bias_act(x, w, bias, "gelu_tanh")
else:
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    x0 = torch.randn(65536, 768, device="cuda").bfloat16()
    x1 = torch.randn(65536, 768, device="cuda")
    g, bt = torch.ones(768, device="cuda").bfloat16(), torch.zeros(768, device="cuda").bfloat16()
    run = lambda: dropout_add_layer_norm(x0, x1, g, bt, 0.0, 1e-5, prenorm=True)

for _ in range(reps):
    run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    run()
e.record()
torch.cuda.synchronize()
print(f"{which}: {a.elapsed_time(e) / reps * 1e3:.1f} us per call")
