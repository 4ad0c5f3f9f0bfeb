"""Incremental decoding: the two decode kernels against their HBM roofline, and generation
throughput of Backpack-Small -- incremental (KV / sense caches) next to the reference's prefix re-run loop
(training/src/utils/generation.py:62-72) with every fused kernel on.  One JSON object per line.

    python benchmarks/bench_decode.py [--which kernels,generate] [--batches 1,8,64] [--prompt 512] [--new 64]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, serving_config  # noqa: E402
from backpacks_flash_attn_b200.ops.decode import decode_attention, sense_mix_decode  # noqa: E402
from backpacks_flash_attn_b200.utils.generation import greedy_decode  # noqa: E402
from backpacks_flash_attn_b200.utils.weights import name_seeded_  # noqa: E402


def peak_hbm():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def time_ms(fn, iters=50, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total / iters


def kernels(peak, peak_key):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > L2 (126 MB): cold-cache launches
    dt = torch.bfloat16
    for b, s in [(1, 1024), (8, 1024), (64, 1024), (64, 512), (8, 4096)]:
        h, dh = 12, 64
        cache = torch.randn(b, s, 2, h, dh, device="cuda", dtype=dt)
        q = torch.randn(b, 1, h, dh, device="cuda", dtype=dt)
        ms = time_ms(lambda: decode_attention(q, cache, s), flush=flush)
        nbytes = cache.numel() * 2 + 2 * q.numel() * 2
        print(json.dumps({"kernel": "bp_decode_attn_fwd", "batch": b, "seqlen_k": s, "nheads": h, "headdim": dh, "ms": ms,
                          "algorithmic_bytes": nbytes, "gbps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                          "peak": peak, "peak_source": peak_key, "l2": "flushed"}), flush=True)
    nv, dk, d, vocab = 16, 48, 768, 50264
    table = torch.randn(vocab, nv, d, device="cuda", dtype=dt)
    for b, s in [(1, 1024), (8, 1024), (64, 1024), (64, 512), (8, 4096)]:
        k_cache = torch.randn(b, s, nv, dk, device="cuda", dtype=dt)
        ids = torch.randint(0, vocab, (b, s), device="cuda")
        q = torch.randn(b, nv, dk, device="cuda", dtype=dt)
        ms = time_ms(lambda: sense_mix_decode(q, k_cache, ids, table, s), flush=flush)
        # every (token, sense) row of the context is read once (rows of repeated tokens may hit L2), K once, ids once
        nbytes = b * s * nv * d * 2 + k_cache.numel() * 2 + ids.numel() * 8 + b * d * 2
        print(json.dumps({"kernel": "bp_sense_mix_decode_fwd", "batch": b, "seqlen": s, "nv": nv, "d": d, "ms": ms,
                          "algorithmic_bytes": nbytes, "gbps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                          "peak": peak, "peak_source": peak_key, "l2": "flushed"}), flush=True)


def generate(batches, prompt, new):
    """Whole-call times for `new` tokens and, from a second call with 4x as many, the steady-state cost per token (the
    slope: prompt pass, graph capture and warm-up cancel)."""
    cfg = serving_config(n_embd=768, n_head=12, n_layer=12, n_positions=1024)
    model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to("cuda", torch.bfloat16)
    model.transformer.build_sense_table()
    modes = (("incremental_graph", True, True), ("incremental", True, False), ("prefix_rerun", False, False))

    def timed(ids, n, inc, graph):
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = greedy_decode(ids, model, prompt + n, incremental=inc, cuda_graph=graph, output_scores=False)
        e.record()
        e.synchronize()
        return a.elapsed_time(e), out.sequences

    for b in batches:
        ids = torch.randint(0, 50257, (b, prompt), device="cuda", generator=torch.Generator("cuda").manual_seed(b))
        res, seqs = {}, {}
        for name, inc, graph in modes:
            greedy_decode(ids, model, prompt + 4, incremental=inc, cuda_graph=graph)   # warm-up (allocator, GEMM heuristics)
            ms1, seq = timed(ids, new, inc, graph)
            ms4, _ = timed(ids, 4 * new, inc, graph)
            per_tok = (ms4 - ms1) / (3 * new)
            res[name] = {"ms_total": ms1, "ms_per_token_steady": per_tok, "tokens_per_s_steady": b / per_tok * 1e3,
                         "fixed_ms": ms1 - per_tok * new}
            seqs[name] = seq
        print(json.dumps({"generate": "backpack-small bf16 (sense table), greedy", "batch": b, "prompt": prompt,
                          "new_tokens": new, **res,
                          "steady_speedup_vs_rerun": {k: res["prefix_rerun"]["ms_per_token_steady"] / res[k]["ms_per_token_steady"]
                                                      for k in ("incremental", "incremental_graph")},
                          "token_agreement_with_rerun": {k: (seqs[k] == seqs["prefix_rerun"]).float().mean().item()
                                                         for k in ("incremental", "incremental_graph")},
                          "note": "prefix_rerun is the reference's loop (generation.py:62-72) on this library's fused "
                                  "kernels, LM head on the last position only; steady = slope between new_tokens and "
                                  "4 x new_tokens (mean context prompt + 2.5 x new_tokens)"}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="kernels,generate")
    ap.add_argument("--batches", default="1,8,64")
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--new", type=int, default=64)   # 4 x new + prompt must fit n_positions = 1024
    args = ap.parse_args()
    peak, key = peak_hbm()
    if "kernels" in args.which:
        kernels(peak, key)
    if "generate" in args.which:
        generate([int(x) for x in args.batches.split(",")], args.prompt, args.new)
