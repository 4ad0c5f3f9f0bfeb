"""First thing to run on a GPU box after touching a kernel: every kernel of the library on a handful of shapes, each
case in its OWN subprocess (a trapped kernel kills the CUDA context of its process only), compared with plain torch
fp32 math.  Prints one PASS/FAIL line per case and exits non-zero when any case failed.

    python benchmarks/sanity.py [--tag dbg]      # BP_LIB_TAG of a debug build to re-run failing cases with
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "fmha_d64_s1024": "fmha(2, 1024, 4, 64, True)",
    "fmha_d64_s200_nc": "fmha(3, 200, 2, 64, False)",
    "fmha_d128_s1024": "fmha(2, 1024, 2, 128, True)",
    "fmha_d40_s257": "fmha(2, 257, 3, 40, True)",
    "fmha_d64_b32": "fmha(32, 1024, 12, 64, True, check=False)",
    "fmha_f16_d64": "fmha(2, 512, 4, 64, True, dtype=torch.float16)",
    "sense_k16_s512": "sense(2, 512, 16, 768)",
    "sense_k16_s200": "sense(2, 200, 16, 768)",
    "sense_k4_s512": "sense(2, 512, 4, 768)",
    "sense_k64_s256": "sense(2, 256, 64, 768)",
    "sense_table_k16": "sense(2, 512, 16, 768, table=True)",
    "sense_table_k4": "sense(2, 300, 4, 768, table=True)",
    "sense_k16_b64": "sense(64, 1024, 16, 768, check=False)",
    "lm_head_stats": "lmstats(2000, 50264, 768)",
    "lm_head_stats_small": "lmstats(300, 1000, 64)",
    "gemm_plain": "gemm(3000, 2304, 768)",
    "decode_attn_d64": "dec_attn(3, 300, 12, 64)",
    "decode_attn_d128": "dec_attn(2, 1025, 4, 128)",
    "decode_sense": "dec_sense(3, 300, 16, 768)",
    "decode_sense_k64": "dec_sense(2, 100, 64, 768)",
}

PRELUDE = r'''
import sys, torch, math
sys.path.insert(0, %r)
def fmha(b, s, h, d, causal, dtype=torch.bfloat16, check=True):
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_with_lse
    torch.manual_seed(0)
    qkv = torch.randn(b, s, 3, h, d, device="cuda").to(dtype)
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    q, k, v = (qkv[:, :, i].reshape(b * s, h, d) for i in range(3))
    out, lse = flash_attn_unpadded_with_lse(q, k, v, cu, cu, s, s, causal=causal)
    torch.cuda.synchronize()
    if not check:
        print("ran", float(out.float().abs().mean())); return
    qf, kf, vf = (qkv[:, :, i].float() for i in range(3))
    sc = torch.einsum("bthd,bshd->bhts", qf, kf) / math.sqrt(d)
    if causal:
        sc = sc.masked_fill(~torch.ones(s, s, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    ref = torch.einsum("bhts,bshd->bthd", torch.softmax(sc, -1), vf)
    err = (out.view(b, s, h, d).float() - ref).abs().max().item()
    lerr = (lse[:, :, :s] - torch.logsumexp(sc, -1)).abs().max().item()
    print(f"max|err| {err:.3e} lse err {lerr:.3e}")
    assert err < 3e-2 and lerr < 1e-3
def lmstats(m, n, k):
    from backpacks_flash_attn_b200.ops.lm_head import lm_head_stats
    torch.manual_seed(0)
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    t = torch.randint(0, n, (m,), device="cuda")
    out = lm_head_stats(x, w, t)
    torch.cuda.synchronize()
    logits = x.float() @ w.float().t()
    e1 = (out["lse"] - torch.logsumexp(logits, -1)).abs().max().item()
    e2 = (out["target_logit"] - logits.gather(-1, t[:, None])[:, 0]).abs().max().item()
    agree = (out["argmax"].long() == logits.argmax(-1)).float().mean().item()
    print(f"lse err {e1:.3e} target err {e2:.3e} argmax agreement {agree:.4f}")
    assert e1 < 2e-3 and e2 < 2e-3 and agree > 0.98
def gemm(m, n, k):
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
    torch.manual_seed(0)
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    b = torch.randn(n, device="cuda").bfloat16()
    for act in ("none", "gelu_tanh"):
        out = linear_bias_act(x, w, b, act)
        ref = x.float() @ w.float().t() + b.float()
        if act == "gelu_tanh": ref = torch.nn.functional.gelu(ref, approximate="tanh")
        err = (out.float() - ref).abs().max().item()
        print(act, f"max|err| {err:.3e}")
        assert err < 5e-2
def dec_attn(b, s, h, d):
    from backpacks_flash_attn_b200.ops.decode import decode_attention
    torch.manual_seed(0)
    cache = torch.randn(b, s + 5, 2, h, d, device="cuda").bfloat16()
    q = torch.randn(b, 1, h, d, device="cuda").bfloat16()
    out = decode_attention(q, cache, s)
    torch.cuda.synchronize()
    k, v = cache[:, :s, 0].float(), cache[:, :s, 1].float()
    p = torch.softmax(torch.einsum("bthd,bshd->bhts", q.float(), k) * d ** -0.5, -1)
    ref = torch.einsum("bhts,bshd->bthd", p, v)
    err = (out.float() - ref).abs().max().item()
    print(f"max|err| {err:.3e}")
    assert err < 2e-2
def dec_sense(b, s, nv, d):
    from backpacks_flash_attn_b200.ops.decode import sense_mix_decode
    torch.manual_seed(0)
    dk = (d // nv + 7) // 8 * 8
    kc = torch.randn(b, s + 5, nv, dk, device="cuda").bfloat16()
    q = torch.randn(b, nv, dk, device="cuda").bfloat16()
    tab = torch.randn(777, nv, d, device="cuda").bfloat16()
    ids = torch.randint(0, 777, (b, s + 5), device="cuda")
    out = sense_mix_decode(q, kc, ids, tab, s)
    torch.cuda.synchronize()
    sc = torch.einsum("bhd,bshd->bhs", q.float(), kc[:, :s].float()) * dk ** -0.5
    ref = torch.einsum("bhs,bshd->bd", torch.softmax(sc, -1), tab[ids[:, :s]].float())
    err = (out.float() - ref).abs().max().item()
    print(f"max|err| {err:.3e} (max|ref| {ref.abs().max().item():.2f})")
    assert err < 0.08
def sense(b, s, nv, d, table=False, check=True):
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix, sense_mix_table
    torch.manual_seed(0)
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
    if table:
        tab = torch.randn(777, nv, d, device="cuda").bfloat16()
        ids = torch.randint(0, 777, (b, s), device="cuda")
        content = tab[ids].transpose(1, 2)
        out, lse = sense_mix_table(qk, tab, ids, return_lse=True)
    else:
        content = torch.randn(b, s, nv, d, device="cuda").bfloat16().transpose(1, 2)
        out, lse = sense_mix(qk, content, return_lse=True)
    torch.cuda.synchronize()
    if not check:
        print("ran", float(out.float().abs().mean())); return
    q, k = qk.float().unbind(2)
    sc = torch.einsum("bthd,bshd->bhts", q, k) * (d // nv) ** -0.5
    sc = sc.masked_fill(~torch.ones(s, s, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    ref = (torch.softmax(sc, -1) @ content.float()).sum(1)
    err = (out.float() - ref).abs().max().item()
    lerr = (lse - torch.logsumexp(sc, -1)).abs().max().item()
    print(f"max|err| {err:.3e} (max|ref| {ref.abs().max().item():.2f}) lse err {lerr:.3e}")
    assert err < 0.12 and lerr < 1e-3
''' % ROOT


def run(name, expr, tag=None):
    env = dict(os.environ)
    if tag:
        env["BP_LIB_TAG"] = tag
    try:
        r = subprocess.run([sys.executable, "-c", PRELUDE + expr], capture_output=True, text=True, timeout=180, env=env)
        ok, out = r.returncode == 0, (r.stdout + r.stderr).strip().splitlines()
    except subprocess.TimeoutExpired:
        ok, out = False, ["TIMEOUT"]
    print(f"{'PASS' if ok else 'FAIL'} {name}{' [' + tag + ']' if tag else ''}: {' | '.join(out[-(2 if ok else 12):])}", flush=True)
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default=None, help="debug library (BP_LIB_TAG) to re-run failing cases with")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    failed = []
    for name, expr in CASES.items():
        if a.only and a.only not in name:
            continue
        if not run(name, expr):
            failed.append(name)
            if a.tag:
                run(name, expr, a.tag)
    print("sanity:", "all passed" if not failed else f"FAILED {failed}")
    sys.exit(1 if failed else 0)
