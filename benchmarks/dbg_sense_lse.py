"""Scratch: bp_sense_lse_fwd against torch.logsumexp for a few shapes, then a timing at config 3."""
import sys, torch
from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
for b, s, nv, dk in [(1, 128, 4, 48), (2, 300, 16, 48), (2, 1024, 16, 48), (1, 777, 5, 16), (2, 512, 3, 192)]:
    qk = torch.randn(b, s, 2, nv, dk, device="cuda").bfloat16()
    c = torch.randn(b, nv, s, 64, device="cuda").bfloat16()
    out, lse = sense_mix(qk, c, return_lse=True)
    torch.cuda.synchronize()
    q, k = qk.float().unbind(2)
    sc = torch.einsum("bthd,bshd->bhts", q, k) * dk ** -0.5
    sc = sc.masked_fill(~torch.ones(s, s, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    ref = torch.logsumexp(sc, -1)
    print(b, s, nv, dk, "max |lse err|", (lse - ref).abs().max().item(), flush=True)
if "--time" in sys.argv:
    from backpacks_flash_attn_b200 import _lib
    b, s, nv, d = 64, 1024, 16, 768
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
    lse = torch.empty(b, nv, s, device="cuda")
    lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.bp_sense_lse_fwd(qk.data_ptr(), lse.data_ptr(), b, s, nv, d // nv, (d // nv) ** -0.5, 1, st)
    for _ in range(5): f()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    e.record(); torch.cuda.synchronize()
    print("lse pass config 3: %.1f us" % (a.elapsed_time(e) / 20 * 1e3))
