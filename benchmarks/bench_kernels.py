"""Kernel-level timings (CUDA events, rotating inputs larger than L2, median of N) with roofline fractions.

    python benchmarks/bench_kernels.py [--which fmha,sense,ln,gemm] [--iters 20]

Algorithmic work per launch follows SURVEY.md §8(d) / BASELINE.md §3.  Peaks come from MEASURED_PEAKS.json
when present (else the fallback stated in B200_PROFILING.md).  Prints one JSON line per kernel.
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def time_fn(fn, variants, iters, warmup=5, inner=8):
    """fn(i) runs variant i; variants rotate so that consecutive launches never see a warm L2.  Each timed
    sample brackets `inner` back-to-back launches with one pair of CUDA events (so the host-side launch
    preparation is hidden behind the previous kernel and does not leak into the measurement)."""
    for i in range(warmup):
        fn(i % variants)
    torch.cuda.synchronize()
    times = []
    k = 0
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(k % variants); k += 1          # keeps the GPU busy while the timed launches are being queued
        a.record()
        for _ in range(inner):
            fn(k % variants); k += 1
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b) * 1e-3 / inner)
    return statistics.median(times), min(times)


def report(name, t_med, t_min, flops, nbytes, extra=None):
    hbm, tf_burst, tf_sus, src = peaks()
    rec = {"kernel": name, "ms_median": t_med * 1e3, "ms_min": t_min * 1e3,
           "tflops": flops / t_med / 1e12, "gbs": nbytes / t_med / 1e9,
           "frac_tensor_burst": flops / t_med / 1e12 / tf_burst, "frac_hbm": nbytes / t_med / 1e9 / hbm,
           "peaks": src}
    if extra:
        rec.update(extra)
    print(json.dumps(rec))
    return rec


def bench_fmha(iters, b=32, s=1024, h=12, d=64, comparators=True):
    from backpacks_flash_attn_b200.flash_attn_interface import flash_attn_unpadded_qkvpacked_func
    nvar = 3  # 3 x 75 MB of qkv + outputs > 126 MB L2
    qkvs = [torch.randn(b * s, 3, h, d, device="cuda").bfloat16() for _ in range(nvar)]
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    fn = lambda i: flash_attn_unpadded_qkvpacked_func(qkvs[i], cu, s, 0.0, causal=True)
    t_med, t_min = time_fn(fn, nvar, iters)
    flops = 4 * b * h * s * s * d / 2
    nbytes = 4 * b * s * h * d * 2 + 4 * b * h * s
    rec = report(f"fmha_fwd b{b} h{h} s{s} d{d} causal bf16", t_med, t_min, flops, nbytes)
    if comparators:
        fmha_comparators(iters, qkvs, b, s, h, d, flops, nbytes)
    return rec


def fmha_comparators(iters, qkvs, b, s, h, d, flops, nbytes):
    """On-box yard-sticks SURVEY.md §8c allows as INFORMATIVE speed comparators (they are other projects' kernels,
    not the reference fork): torch SDPA with the cuDNN / flash backends, and the site-packages flash_attn 2.8 (FA-2,
    mma.sync).  Same inputs, same timing loop."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    views = [q.view(b, s, 3, h, d).permute(2, 0, 3, 1, 4) for q in qkvs]     # (3, b, h, s, d) strided views
    for name, backend in (("torch SDPA, cuDNN backend", SDPBackend.CUDNN_ATTENTION),
                          ("torch SDPA, flash backend", SDPBackend.FLASH_ATTENTION)):
        try:
            def fn(i, backend=backend):
                with sdpa_kernel(backend):
                    return F.scaled_dot_product_attention(views[i][0], views[i][1], views[i][2], is_causal=True)
            t, tm = time_fn(fn, len(qkvs), iters)
            report(f"  (comparator, informative) {name}", t, tm, flops, nbytes)
        except Exception as e:  # backend not available for this shape / build
            print(json.dumps({"kernel": f"(comparator) {name}", "unavailable": str(e)[:200]}))
    try:
        import importlib
        sys_path = list(sys.path)
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]   # the site-packages flash_attn, not a local dir
        fa = importlib.import_module("flash_attn")
        sys.path = sys_path
        packed = [q.view(b, s, 3, h, d) for q in qkvs]
        t, tm = time_fn(lambda i: fa.flash_attn_qkvpacked_func(packed[i], causal=True), len(qkvs), iters)
        report(f"  (comparator, informative) site-packages flash_attn {getattr(fa, '__version__', '?')} (FA-2)", t, tm, flops, nbytes)
    except Exception as e:
        print(json.dumps({"kernel": "(comparator) site-packages flash_attn", "unavailable": str(e)[:200]}))


def bench_fmha_bwd(iters, b=32, s=1024, h=12, d=64, comparators=True):
    """bp_fmha_bwd (three launches) at BASELINE config 2.  FLOPs by the usual convention: 2.5 x the forward
    (five tile products against two); bytes: q, k, v, o, dO read and dq, dk, dv written once."""
    from backpacks_flash_attn_b200.flash_attn_interface import _flash_attn_backward, _flash_attn_forward
    nvar = 3
    cu = torch.arange(0, (b + 1) * s, s, dtype=torch.int32, device="cuda")
    sets = []
    for _ in range(nvar):
        qkv = torch.randn(b * s, 3, h, d, device="cuda").bfloat16()
        out, lse = _flash_attn_forward(qkv[:, 0], qkv[:, 1], qkv[:, 2], torch.empty_like(qkv[:, 0]), cu, cu, s, s,
                                       d ** -0.5, True)
        sets.append((qkv, out, lse, torch.randn_like(out), torch.empty_like(qkv)))

    def fn(i):
        qkv, out, lse, g, dqkv = sets[i]
        _flash_attn_backward(g, qkv[:, 0], qkv[:, 1], qkv[:, 2], out, lse, dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], cu, cu,
                             s, s, d ** -0.5, True)
    t_med, t_min = time_fn(fn, nvar, iters)
    flops = 2.5 * 4 * b * h * s * s * d / 2
    nbytes = 8 * b * s * h * d * 2 + 4 * b * h * s
    rec = report(f"fmha_bwd (stats + dKdV + dQ) b{b} h{h} s{s} d{d} causal bf16", t_med, t_min, flops, nbytes)
    if comparators:
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for name, backend in (("torch SDPA backward, cuDNN backend", SDPBackend.CUDNN_ATTENTION),
                              ("torch SDPA backward, flash backend", SDPBackend.FLASH_ATTENTION)):
            try:
                graphs = []
                for qkv, _, _, g, _ in sets:
                    v = qkv.view(b, s, 3, h, d).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)
                    with sdpa_kernel(backend):
                        o = F.scaled_dot_product_attention(v[0], v[1], v[2], is_causal=True)
                    graphs.append((o, v, g.view(b, s, h, d).transpose(1, 2)))
                t, tm = time_fn(lambda i: torch.autograd.grad(graphs[i][0], graphs[i][1], graphs[i][2], retain_graph=True),
                                nvar, iters)
                report(f"  (comparator, informative) {name}", t, tm, flops, nbytes)
            except Exception as e:
                print(json.dumps({"kernel": f"(comparator) {name}", "unavailable": str(e)[:200]}))
        try:
            import importlib
            sys_path = list(sys.path)
            sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
            fa = importlib.import_module("flash_attn")
            sys.path = sys_path
            graphs = []
            for qkv, _, _, g, _ in sets:
                v = qkv.view(b, s, 3, h, d).detach().requires_grad_(True)
                graphs.append((fa.flash_attn_qkvpacked_func(v, causal=True), v, g.view(b, s, h, d)))
            t, tm = time_fn(lambda i: torch.autograd.grad(graphs[i][0], graphs[i][1], graphs[i][2], retain_graph=True),
                            nvar, iters)
            report(f"  (comparator, informative) site-packages flash_attn {getattr(fa, '__version__', '?')} backward (FA-2)",
                   t, tm, flops, nbytes)
        except Exception as e:
            print(json.dumps({"kernel": "(comparator) site-packages flash_attn backward", "unavailable": str(e)[:200]}))
    return rec


def bench_bwd_ops(iters, rows=65536, cols=768, inner=3072):
    """The HBM-bound backward passes at config-3 size: bp_ln_residual_bwd and bp_bias_act_bwd."""
    from backpacks_flash_attn_b200.ops.fused_dense import bias_act_backward
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    x0 = torch.randn(rows, cols, device="cuda").bfloat16().requires_grad_()
    x1 = torch.randn(rows, cols, device="cuda").requires_grad_()
    w = torch.ones(cols, device="cuda").bfloat16().requires_grad_()
    bb = torch.zeros(cols, device="cuda").bfloat16().requires_grad_()
    z, r = dropout_add_layer_norm(x0, x1, w, bb, 0.0, 1e-5, prenorm=True)
    g, g2 = torch.randn_like(z), torch.randn_like(r)
    t, tm = time_fn(lambda i: torch.autograd.grad((z, r), (x0, x1, w, bb), (g, g2), retain_graph=True), 1, iters)
    # dz (2) + x (4) + dx_residual (4) in, dx0 (2) + dx1 (4) out
    report(f"ln_residual_bwd {rows}x{cols} bf16 / fp32 residual", t, tm, 0, rows * cols * 16)
    # residual dropout inside the same kernels (training configuration of the reference: resid_pdrop = 0.1)
    zd, rd = dropout_add_layer_norm(x0, x1, w, bb, 0.1, 1e-5, prenorm=True, seed=1)
    t, tm = time_fn(lambda i: torch.autograd.grad((zd, rd), (x0, x1, w, bb), (g, g2), retain_graph=True), 1, iters)
    report(f"ln_residual_bwd with dropout 0.1 {rows}x{cols} bf16 / fp32 residual", t, tm, 0, rows * cols * 16)
    with torch.no_grad():
        t, tm = time_fn(lambda i: dropout_add_layer_norm(x0, x1, w, bb, 0.1, 1e-5, prenorm=True, seed=1), 1, iters)
        report(f"ln_residual_fwd with dropout 0.1 {rows}x{cols}", t, tm, 0, rows * cols * 12)
        t, tm = time_fn(lambda i: dropout_add_layer_norm(x0, x1, w, bb, 0.0, 1e-5, prenorm=True), 1, iters)
        report(f"ln_residual_fwd without dropout {rows}x{cols}", t, tm, 0, rows * cols * 12)
        t, tm = time_fn(lambda i: torch.nn.functional.dropout(x0, 0.1, training=True), 1, iters)
        report("  (for comparison: the separate F.dropout pass over x0 this replaces, forward only)", t, tm, 0, rows * cols * 4)
    d = torch.randn(rows, inner, device="cuda").bfloat16()
    pre = torch.randn(rows, inner, device="cuda").bfloat16()
    t, tm = time_fn(lambda i: bias_act_backward(d, pre, "gelu_tanh", True), 1, iters)
    report(f"bias_act_bwd (dgelu + bias grad) {rows}x{inner} bf16", t, tm, 0, rows * inner * 6)
    d2 = torch.randn(rows, cols, device="cuda").bfloat16()
    t, tm = time_fn(lambda i: bias_act_backward(d2, None, "none", True), 1, iters)
    report(f"bias_act_bwd (bias grad only) {rows}x{cols} bf16", t, tm, 0, rows * cols * 2)


def bench_xent(iters, rows=65536, vocab=50264):
    """bp_xentropy_fwd / bp_xentropy_bwd (in place) at config-3 size: 6.6 GB of bf16 logits."""
    from backpacks_flash_attn_b200 import _lib
    x = torch.randn(rows, vocab, device="cuda", dtype=torch.bfloat16)
    y = torch.randint(0, vocab - 7, (rows,), device="cuda")
    losses = torch.empty(rows, device="cuda")
    lse = torch.empty(rows, device="cuda")
    g = torch.full((rows,), 1.0 / rows, device="cuda")
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    fwd = lambda i: _lib.check(lib.bp_xentropy_fwd(x.data_ptr(), y.data_ptr(), losses.data_ptr(), lse.data_ptr(), rows,
                                                   vocab, vocab, 0.0, -100, -1, 1, st), "bp_xentropy_fwd")
    t, tm = time_fn(fwd, 1, iters, inner=2)
    report(f"xentropy_fwd {rows}x{vocab} bf16", t, tm, 0, rows * vocab * 2)
    grad = torch.empty_like(x)
    bwd = lambda i: _lib.check(lib.bp_xentropy_bwd(g.data_ptr(), x.data_ptr(), lse.data_ptr(), y.data_ptr(), grad.data_ptr(),
                                                   rows, vocab, vocab, vocab, 0.0, -100, -1, 1, st), "bp_xentropy_bwd")
    t, tm = time_fn(bwd, 1, iters, inner=2)
    report(f"xentropy_bwd {rows}x{vocab} bf16", t, tm, 0, rows * vocab * 4)
    ref = lambda i: torch.nn.functional.cross_entropy(x.float(), y)
    t, tm = time_fn(ref, 1, max(3, iters // 4), inner=1)
    report("  (comparator, informative) F.cross_entropy on the fp32 upcast, forward only", t, tm, 0, rows * vocab * 2)


def bench_sense(iters, b=64, s=1024, nv=16, d=768):
    from backpacks_flash_attn_b200.ops.sense_mix import sense_mix
    from backpacks_flash_attn_b200 import _lib
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
    content = torch.randn(b, s, nv, d, device="cuda").bfloat16().transpose(1, 2)   # 1.6 GB >> L2
    fn = lambda i: sense_mix(qk, content)
    t_med, t_min = time_fn(fn, 1, iters)
    flops = b * s * s * d * (1 + nv)
    nbytes = (2 * b * s * d + nv * b * s * d + b * s * d) * 2
    rec = report(f"sense_mix(lse+mix) b{b} s{s} k{nv} d{d} bf16", t_med, t_min, flops, nbytes)
    # split the two passes
    lib = _lib.load()
    lse = torch.empty(b, nv, s, device="cuda", dtype=torch.float32)
    out = torch.empty(b, s, d, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    scale = (d // nv) ** -0.5
    f1 = lambda i: lib.bp_sense_lse_fwd(qk.data_ptr(), lse.data_ptr(), b, s, nv, d // nv, scale, 1, st)
    f2 = lambda i: lib.bp_sense_mix_fwd(qk.data_ptr(), content.data_ptr(), lse.data_ptr(), out.data_ptr(), b, s, nv,
                                        d // nv, d, content.stride(0), content.stride(1), content.stride(2), scale, 1, st)
    t1, _ = time_fn(f1, 1, iters)
    t2, _ = time_fn(f2, 1, iters)
    report("  sense_lse pass", t1, t1, b * nv * s * s * (d // nv), 2 * b * s * d * 2)
    report("  sense_mix pass", t2, t2, b * s * s * d * (1 + nv), nbytes)
    # table mode: the sense vectors are gathered inside the kernel from a (vocab, nv, d) table (1.23 GB at Small size)
    vocab = 50264
    table = torch.randn(vocab, nv, d, device="cuda").bfloat16()
    ids = torch.randint(0, 50257, (b, s), device="cuda")
    f3 = lambda i: lib.bp_sense_mix_table_fwd(qk.data_ptr(), table.data_ptr(), ids.data_ptr(), lse.data_ptr(),
                                              out.data_ptr(), b, s, nv, d // nv, d, vocab, scale, 1, st)
    t3, _ = time_fn(f3, 1, iters)
    report("  sense_mix pass, table mode (gather inside the kernel)", t3, t3, b * s * s * d * (1 + nv), nbytes)
    f4 = lambda i: torch.nn.functional.embedding(ids, table.view(vocab, -1))
    t4, _ = time_fn(f4, 1, iters)
    report("  (for comparison: ATen gather that materialises (b, s, nv, d))", t4, t4, 0, 2 * nv * b * s * d * 2)
    return rec


def bench_sense_bwd(iters, b=64, s=1024, nv=16, d=768):
    """Sense-mix backward at config-3 size: the hand-derived backward (library GEMMs around bp_sense_softmax_bwd) next
    to autograd through the reference's eager composition (what the reference trains through), and the element-wise
    pass alone."""
    from backpacks_flash_attn_b200.ops.sense_mix import _sense_mix_backward, _sense_mix_backward_eager
    from backpacks_flash_attn_b200 import _lib
    qk = torch.randn(b, s, 2, nv, d // nv, device="cuda").bfloat16()
    content = (torch.randn(b, s, nv, d, device="cuda") * 0.5).bfloat16().transpose(1, 2)
    dout = torch.randn(b, s, d, device="cuda").bfloat16()
    scale = (d // nv) ** -0.5
    causal_flops = b * nv * s * s * (2 * d + 3 * (d // nv))      # five products, causal half of each (2 flops / MAC)
    it = max(3, iters // 4)
    t, tm = time_fn(lambda i: _sense_mix_backward(qk, content, dout, scale, True, True), 1, it, warmup=2, inner=1)
    rec = report(f"sense_mix backward (GEMMs + bp_sense_softmax_bwd) b{b} s{s} k{nv} d{d} bf16", t, tm, causal_flops, 0)
    te, tme = time_fn(lambda i: _sense_mix_backward_eager(qk, content, dout, scale, True, True), 1, it, warmup=2, inner=1)
    report("  (comparator) autograd through the eager composition, 1 GB chunks", te, tme, causal_flops, 0,
           {"speedup_of_ours": te / t})
    S = torch.randn(nv, b, s, s, device="cuda").bfloat16()
    dA = torch.randn(nv, b, s, s, device="cuda").bfloat16()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    f = lambda i: _lib.check(lib.bp_sense_softmax_bwd(S.data_ptr(), dA.data_ptr(), nv * b * s, s, scale, 1, st), "ssb")
    t1, t1m = time_fn(f, 1, iters, inner=2)
    report("  bp_sense_softmax_bwd alone (reads the causal half, writes full rows)", t1, t1m, 0, nv * b * s * s * 2 * 3)
    return rec


def bench_ln(iters, rows=65536, cols=768):
    from backpacks_flash_attn_b200.ops.layer_norm import dropout_add_layer_norm
    nvar = 2
    x0 = [torch.randn(rows, cols, device="cuda").bfloat16() for _ in range(nvar)]
    x1 = [torch.randn(rows, cols, device="cuda") for _ in range(nvar)]
    g = torch.ones(cols, device="cuda").bfloat16()
    bta = torch.zeros(cols, device="cuda").bfloat16()
    fn = lambda i: dropout_add_layer_norm(x0[i], x1[i], g, bta, 0.0, 1e-5, prenorm=True)
    t_med, t_min = time_fn(fn, nvar, iters)
    return report(f"ln_residual_fwd rows{rows} cols{cols}", t_med, t_min, 0, rows * cols * (2 + 4) * 2)


def bench_gemm(iters, m=65536, n=3072, k=768):
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    bias = torch.randn(n, device="cuda").bfloat16()
    fn = lambda i: linear_bias_act(x, w, bias, "gelu_tanh")
    t_med, t_min = time_fn(fn, 1, iters)
    rec = report(f"linear_bias_gelu m{m} n{n} k{k}", t_med, t_min, 2 * m * n * k, (m * k + n * k + m * n) * 2)
    lib_fn = lambda i: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, bias), approximate="tanh")
    t_lib, _ = time_fn(lib_fn, 1, iters)
    report("  (cuBLAS F.linear + F.gelu, for comparison)", t_lib, t_lib, 2 * m * n * k, (m * k + n * k + 3 * m * n) * 2)
    lin_fn = lambda i: torch.nn.functional.linear(x, w, bias)
    t_lin, _ = time_fn(lin_fn, 1, iters)
    report("  (cuBLAS F.linear only, for comparison)", t_lin, t_lin, 2 * m * n * k, (m * k + n * k + m * n) * 2)
    return rec


def bench_gemms(iters):
    """Every linear of the Backpack-Small forward at config 3 (m = 65536 rows): this library's tcgen05 GEMM (plain
    epilogue: + bias) next to the cuBLAS call `F.linear` makes.  One JSON line per shape."""
    from backpacks_flash_attn_b200.ops.fused_dense import linear_bias_act
    m = 65536
    shapes = [("Wqkv", 2304, 768), ("out_proj", 768, 768), ("fc2", 768, 3072), ("fc1 (plain)", 3072, 768),
              ("ctx Wqkv", 1536, 768), ("final_mlp.fc2", 12288, 3072), ("lm_head (no bias)", 50264, 768)]
    for name, n, k in shapes:
        x = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
        bias = None if "no bias" in name else torch.randn(n, device="cuda").bfloat16()
        it = max(4, iters // 4) if n > 8000 else iters
        t_own, _ = time_fn(lambda i: linear_bias_act(x, w, bias, "none"), 1, it, inner=4)
        t_lib, _ = time_fn(lambda i: torch.nn.functional.linear(x, w, bias), 1, it, inner=4)
        flops = 2.0 * m * n * k
        print(json.dumps({"gemm": name, "m": m, "n": n, "k": k, "own_us": t_own * 1e6, "cublas_us": t_lib * 1e6,
                          "own_tflops": flops / t_own / 1e12, "cublas_tflops": flops / t_lib / 1e12,
                          "own_over_cublas_time": t_own / t_lib}), flush=True)
        if "lm_head" in name:
            # the LM head with the softmax statistics in the epilogue (no logits in HBM) next to what it replaces in
            # evaluation: the GEMM above followed by a cross-entropy over the logits
            from backpacks_flash_attn_b200.ops.lm_head import lm_head_stats
            tgt = torch.randint(0, 50257, (m,), device="cuda")
            t_stats, _ = time_fn(lambda i: lm_head_stats(x, w, tgt, n_valid=50257), 1, it, inner=4)
            logits = torch.nn.functional.linear(x, w, bias)
            t_ce, _ = time_fn(lambda i: torch.nn.functional.cross_entropy(logits, tgt, reduction="none"), 1, 4, inner=1)
            print(json.dumps({"gemm": "lm_head + softmax statistics (bp_lm_head_stats_fwd: lse, argmax, target logit; no logits written)",
                              "m": m, "n": n, "k": k, "own_us": t_stats * 1e6, "own_tflops": flops / t_stats / 1e12,
                              "replaces_us": {"cublas_gemm": t_lib * 1e6, "torch_cross_entropy_on_bf16_logits": t_ce * 1e6},
                              "own_over_replaced_time": t_stats / (t_lib + t_ce)}), flush=True)
            del logits
        del x, w


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="fmha,sense,ln,gemm")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shape", default="", help="b,s,h,d for the attention benchmarks (default: BASELINE config 2)")
    ap.add_argument("--no-comparators", action="store_true")
    a = ap.parse_args()
    for w in a.which.split(","):
        fn = {"fmha": bench_fmha, "sense": bench_sense, "ln": bench_ln, "gemm": bench_gemm, "gemms": bench_gemms,
              "fmha_bwd": bench_fmha_bwd, "bwd_ops": bench_bwd_ops, "xent": bench_xent, "sense_bwd": bench_sense_bwd}[w]
        if w in ("fmha", "fmha_bwd") and (a.shape or a.no_comparators):
            b, s_, h, d = (int(x) for x in a.shape.split(",")) if a.shape else (32, 1024, 12, 64)
            fn(a.iters, b=b, s=s_, h=h, d=d, comparators=not a.no_comparators)
        else:
            fn(a.iters)
