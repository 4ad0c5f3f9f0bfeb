#!/usr/bin/env python
"""Headline benchmark: forward tokens/s of Backpack-Small (d=768, 12 layers, 12 heads, k=16 senses), bf16,
seq 1024, batch 64 per GPU (BASELINE.json configs[2]; configs[3] = 8 GPUs x 64), synthetic ids and
name-seeded random weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--seqlen S]

N > 1 is launched by torchrun (one process per GPU); the batch is sharded with no data-path collective
(weak scaling: every rank runs B sequences).  Rank 0 prints ONE JSON line:
  value    whole-job tokens/s with the ids resident in HBM (CUDA events, barrier + synchronize on both sides,
           max over ranks).  The step is replayed as one CUDA graph (utils/graph.py; `--no-graph` launches every
           kernel from Python instead) -- the same kernels on the same buffers, without the gaps between launches;
  e2e      the same forward through the public API with HOST inputs: per step a pinned-host -> device copy of
           the ids and a device -> host read of the last-position logits (what the reference's generation loop
           consumes, training/src/utils/generation.py:34-44); the host reads step i while step i+1 runs;
  roofline the kernel of this library with the largest share of the step, `kernels` all of them: average launch
           durations from CUDA events recorded on the launching stream around every C-ABI call during an eager
           pass of the same K steps (events cannot be recorded inside a graph replay), against the measured peaks;
  cpu_baseline  the oracle port of the reference's pure-PyTorch path on the host cores (bounded sample).
`--impl reference` times that CPU path alone (the reference's CUDA attention cannot run on sm_100, and its
Python cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SMALL = dict(n_embd=768, n_head=12, n_layer=12, n_positions=1024)
METRIC = "tokens/sec fwd Backpack-Small seq1024 k=16"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"],
                "tf_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def make_ids(global_batch: int, seqlen: int) -> torch.Tensor:
    return torch.randint(0, 50257, (global_batch, seqlen), generator=torch.Generator().manual_seed(1234))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's pure-PyTorch path (test/bench infrastructure only)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(seqlen: int, steps: int, warmup: int, sample_batch: int = 1):
    from oracle import backpack_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(**{**SMALL, "n_positions": max(1024, seqlen)})
    w = O.name_seeded_weights(cfg)
    ids = make_ids(sample_batch, seqlen)
    with torch.inference_mode():
        for _ in range(warmup):
            O.backpack_logits(ids, w, cfg)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.backpack_logits(ids, w, cfg)
        dt = time.perf_counter() - t0
    return {"value": sample_batch * seqlen * steps / dt, "unit": "tokens/s", "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"Backpack-Small fp32 eager forward, ids ({sample_batch},{seqlen}), {steps} timed steps after "
                      f"{warmup} warm-up, {dt / steps * 1e3:.0f} ms/step"}, dt / steps


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    base, ms = cpu_reference_run(args.seqlen, args.steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Backpack-Small forward, seq 1024, k=16, d=768; CPU arm runs a (1,1024) sample per step",
                       "seq_len": args.seqlen, "per_step_batch": 1},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's pure-PyTorch CPU path restated in oracle/ (its CUDA fmha refuses sm_100, "
                    "csrc/flash_attn/fmha_api.cpp:206-210; its Python cannot travel to the GPU box)"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from backpacks_flash_attn_b200 import _lib, parallel
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
    from backpacks_flash_attn_b200.utils.weights import name_seeded_

    rank, local_rank, world = parallel.init_distributed()
    if world != args.gpus and rank == 0:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.load().bp_check_device(), "bp_check_device", launched=False)

    B, S = args.batch, args.seqlen
    cfg = flash_config(**{**SMALL, "n_positions": max(1024, S)})
    model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to(dev, torch.bfloat16)
    parallel.assert_replicas_match(model)
    ids_host = parallel.shard_batch(make_ids(B * world, S), rank, world).contiguous().pin_memory()
    ids_dev = ids_host.to(dev)
    last_host = torch.empty((2, B, cfg.vocab_size), dtype=torch.bfloat16).pin_memory()   # double-buffered results

    graphed = None   # set after warm-up (capture needs inference_mode)

    def step_resident():
        return graphed() if graphed is not None else model(ids_dev).logits

    copy_stream = torch.cuda.Stream(device=dev)

    def run_e2e(steps):
        """Serving loop through the public API with host buffers.  Every step copies its ids from pinned host
        memory and its last-position logits back to pinned host memory; the host waits for (consumes) the result
        of step i while step i+1 is already running, so the device never idles behind the host's launch loop."""
        pending = None
        checksum = 0.0
        main = torch.cuda.current_stream()
        for i in range(steps):
            if graphed is not None:
                logits = graphed(ids_host)               # pinned host -> static device input, graph replay
            else:
                logits = model(ids_host.to(dev, non_blocking=True)).logits
            last_dev = logits[:, -1].contiguous()
            del logits
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(copy_stream):      # the read-back leaves the compute stream free for step i+1
                copy_stream.wait_event(ready)
                last_host[i % 2].copy_(last_dev, non_blocking=True)
                last_dev.record_stream(copy_stream)
                done = torch.cuda.Event()
                done.record(copy_stream)
            if pending is not None:
                pending[0].synchronize()
                checksum += float(last_host[pending[1], 0, 0])   # the host touches the result
            pending = (done, i % 2)
        pending[0].synchronize()
        checksum += float(last_host[pending[1], 0, 0])
        return checksum

    with torch.inference_mode():
        for _ in range(args.warmup):
            step_resident()
        torch.cuda.synchronize()
        if args.graph:
            from backpacks_flash_attn_b200.utils.graph import GraphedForward
            graphed = GraphedForward(model, ids_dev)
            for _ in range(args.warmup):
                step_resident()
        # ---- timed region 1: inputs resident in HBM ----
        with ClockSampler(physical_gpu_index(local_rank)) as clocks:
            parallel.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step_resident()
            e1.record()
            torch.cuda.synchronize()
            parallel.barrier()
        dt = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
        tokens = parallel.sum_over_ranks(float(B * S * args.steps), dev)
        # ---- kernel region: the same K steps launched from Python, every kernel of this library bracketed by CUDA
        #      events on its stream (the per-kernel durations behind `roofline` / `kernels`) ----
        graph_on, graphed = graphed, None                      # this region launches from Python
        launches0 = _lib.total_launches()
        per_kernel = {}
        timers = [_lib.KernelTimer(n) for n in ("bp_fmha_fwd", "bp_sense_lse_fwd", "bp_sense_mix_fwd",
                                                "bp_linear_bias_act_fwd", "bp_linear_bias_residual_fwd",
                                                "bp_ln_residual_fwd", "bp_ln_fwd")]
        for t in timers:
            t.__enter__()
        parallel.barrier()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.steps):
            step_resident()
        k1.record()
        torch.cuda.synchronize()
        for t in reversed(timers):
            t.__exit__(None, None, None)
        for t in timers:
            per_kernel[t.name] = t.mean_ms()
            per_kernel[t.name + ":n"] = len(t.events) / args.steps
        launches = _lib.total_launches() - launches0          # per K steps; a graph replay launches the same kernels
        dt_eager = parallel.max_over_ranks(k0.elapsed_time(k1) * 1e-3, dev)
        graphed = graph_on
        # ---- timed region 2: end to end through the public API with host buffers ----
        run_e2e(2)
        parallel.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        run_e2e(args.steps)
        g1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        parallel.barrier()
        dt_e2e = parallel.max_over_ranks(max(g0.elapsed_time(g1) * 1e-3, wall), dev)

        # ---- secondary variant: sense vectors gathered from a precomputed (vocab, nv, d) table ----
        table_ms = None
        if args.sense_table:
            graphed = None                                    # the variant runs the eager launch sequence
            model.transformer.build_sense_table()
            for _ in range(2):
                step_resident()
            parallel.barrier()
            torch.cuda.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(args.steps):
                step_resident()
            h1.record()
            torch.cuda.synchronize()
            parallel.barrier()
            table_ms = parallel.max_over_ranks(h0.elapsed_time(h1) * 1e-3, dev)
            model.transformer.drop_sense_table()

    if rank != 0:
        return
    peaks = load_peaks()
    h, dh, nv, d = cfg.n_head, cfg.n_embd // cfg.n_head, cfg.num_content_vectors, cfg.n_embd
    # algorithmic work per launch
    fmha_flops = 4 * B * h * S * S * dh / 2
    fmha_bytes = 4 * B * S * h * dh * 2 + 4 * B * h * S
    mix_flops = B * S * S * d * (1 + nv)
    mix_bytes = (2 * B * S * d + nv * B * S * d + B * S * d) * 2
    fmha_t = per_kernel["bp_fmha_fwd"] * 1e-3
    mix_t = (per_kernel["bp_sense_lse_fwd"] + per_kernel["bp_sense_mix_fwd"]) * 1e-3
    gemm_t = per_kernel["bp_linear_bias_act_fwd"] * 1e-3
    ln_t = per_kernel["bp_ln_residual_fwd"] * 1e-3
    peak_tf = peaks["tf_sustained"]
    inner = cfg.n_inner or 4 * d
    M = B * S
    gemm_flops = 2.0 * M * inner * d                          # every fused GEMM+GELU launch is (B*S, 4d, d)
    gemm_bytes = (M * d + inner * d + M * inner) * 2
    ln_bytes = M * d * (2 + 4) * 2                            # x0 bf16 + residual fp32 in, z bf16 + residual fp32 out
    # GEMMs with the residual add in the epilogue: per layer out_proj (d x d) and fc2 (d x 4d), + the content
    # block's fc2; the mean over the launches of a step is what the live timer measures
    n_res = per_kernel["bp_linear_bias_residual_fwd:n"]
    n_fc2 = (n_res + 1) // 2 if n_res else 0
    n_out = n_res - n_fc2
    res_flops = (n_out * 2.0 * M * d * d + n_fc2 * 2.0 * M * d * inner) / max(n_res, 1)
    res_bytes = (n_out * (M * d * 2 + d * d * 2) + n_fc2 * (M * inner * 2 + inner * d * 2)) / max(n_res, 1) + M * d * 8
    res_t = per_kernel["bp_linear_bias_residual_fwd"] * 1e-3
    lnf_t = per_kernel["bp_ln_fwd"] * 1e-3
    lnf_bytes = M * d * (4 + 2)                               # fp32 residual in, bf16 z out

    def roof_tensor(flops, t):
        a = flops / t / 1e12
        return {"bound": "tensor", "achieved": a, "peak": peak_tf, "unit": "TFLOP/s", "frac": a / peak_tf}

    kernels = {}
    k = roof_tensor(fmha_flops, fmha_t)
    k.update({"kernel": "fmha_fwd_kernel<64,bf16> (bp_fmha_fwd)", "launches_per_step": per_kernel["bp_fmha_fwd:n"],
              "ms_per_launch": fmha_t * 1e3, "algorithmic_gflop_per_launch": fmha_flops / 1e9,
              "algorithmic_mb_per_launch": fmha_bytes / 1e6, "hbm_gbs": fmha_bytes / fmha_t / 1e9,
              "hbm_frac": fmha_bytes / fmha_t / 1e9 / peaks["hbm_gbs"],
              "traffic": ncu_traffic("fmha_fwd_kernel") if (B, S) == (64, 1024) else None})
    kernels["fmha"] = k
    k = roof_tensor(mix_flops, mix_t)
    k.update({"kernel": "sense_lse_kernel + sense_mix_kernel (bp_sense_lse_fwd, bp_sense_mix_fwd)",
              "launches_per_step": 1, "ms_per_launch": mix_t * 1e3, "ms_lse": per_kernel["bp_sense_lse_fwd"],
              "ms_mix": per_kernel["bp_sense_mix_fwd"], "algorithmic_gflop_per_launch": mix_flops / 1e9,
              "algorithmic_mb_per_launch": mix_bytes / 1e6,
              "traffic": ncu_traffic("sense_mix_kernel") if (B, S) == (64, 1024) else None})
    kernels["sense_mix"] = k
    k = roof_tensor(gemm_flops, gemm_t)
    k.update({"kernel": "gemm_bias_act_pair_kernel<bf16> (bp_linear_bias_act_fwd), fc1 + bias + tanh-GELU",
              "launches_per_step": per_kernel["bp_linear_bias_act_fwd:n"], "ms_per_launch": gemm_t * 1e3,
              "algorithmic_gflop_per_launch": gemm_flops / 1e9, "algorithmic_mb_per_launch": gemm_bytes / 1e6,
              "traffic": ncu_traffic("gemm_bias_act_pair_kernel") if (B, S) == (64, 1024) else None})
    kernels["gemm_bias_gelu"] = k
    a = ln_bytes / ln_t / 1e9
    kernels["ln_residual"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": a / peaks["hbm_gbs"], "kernel": "ln_residual_fwd_kernel (bp_ln_residual_fwd)",
                              "launches_per_step": per_kernel["bp_ln_residual_fwd:n"], "ms_per_launch": ln_t * 1e3,
                              "algorithmic_mb_per_launch": ln_bytes / 1e6,
                              "traffic": ncu_traffic("ln_residual_fwd_kernel") if (B, S) == (64, 1024) else None}
    if n_res:
        k = roof_tensor(res_flops, res_t)
        k.update({"kernel": "gemm_bias_act_pair_kernel<bf16>, residual epilogue (bp_linear_bias_residual_fwd): out_proj "
                            "and fc2 with the fp32 residual add in the epilogue",
                  "launches_per_step": n_res, "ms_per_launch": res_t * 1e3,
                  "algorithmic_gflop_per_launch": res_flops / 1e9, "algorithmic_mb_per_launch": res_bytes / 1e6,
                  "hbm_frac": res_bytes / res_t / 1e9 / peaks["hbm_gbs"], "traffic": None})
        kernels["gemm_bias_residual"] = k
    if per_kernel["bp_ln_fwd:n"]:
        a = lnf_bytes / lnf_t / 1e9
        kernels["ln_from_residual"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                       "frac": a / peaks["hbm_gbs"], "kernel": "ln_residual_fwd_kernel<float -> bf16> (bp_ln_fwd)",
                                       "launches_per_step": per_kernel["bp_ln_fwd:n"], "ms_per_launch": lnf_t * 1e3,
                                       "algorithmic_mb_per_launch": lnf_bytes / 1e6, "traffic": None}
    for v in kernels.values():
        v["ms_per_step"] = v["ms_per_launch"] * v["launches_per_step"]
        v["traffic_unit"] = "DRAM bytes per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum (profiles/)"
    # `roofline` = the kernel of this library with the largest share of the step
    dominant = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
    roofline = dict(kernels[dominant])
    roofline["dominant_of"] = {n: round(v["ms_per_step"], 3) for n, v in kernels.items()}
    roofline["peak_source"] = (f"MEASURED_PEAKS.json ({peaks['source']}): bf16_tflops_sustained for tensor-bound kernels "
                               "(timed inside a long step), hbm_gbs for HBM-bound ones")
    model_flops_per_token = 371.3e6   # SURVEY.md §8d, s = 1024
    cpu_base, _ = cpu_reference_run(S, steps=3, warmup=1)
    value = tokens / dt
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Backpack-Small forward (BASELINE configs[2]; configs[3] when n_gpus=8)",
                   "batch_per_gpu": B, "global_batch": B * world, "seq_len": S, "d_model": d, "n_layer": cfg.n_layer,
                   "n_head": h, "num_content_vectors": nv, "vocab": cfg.vocab_size, "parallelism": f"dp{world}",
                   "weights": "name-seeded random (SURVEY.md §8c recipe), bf16",
                   "execution": ("one CUDA graph per step (utils/graph.py); the eager launch sequence of the same "
                                 f"kernels, with an event pair recorded around each of them, takes {dt_eager / args.steps * 1e3:.3f} ms per step" if args.graph
                                 else "eager launches from Python"),
                   "kernel_timings": "per-kernel CUDA events around an eager pass of the same K steps on the same inputs",
                   "l2": "no flush: each step streams > 10 GB of activations (6.6 GB logits, 1.6 GB sense vectors) "
                         "through a 126 MB L2"},
        "e2e": {"value": tokens / dt_e2e, "unit": "tokens/s", "h2d_bytes_per_step": ids_host.numel() * 8 * world,
                "d2h_bytes_per_step": last_host[0].numel() * 2 * world, "ms_per_step": dt_e2e / args.steps * 1e3,
                "result": "last-position logits (batch, vocab) bf16 copied to pinned host memory every step; the host reads step i while step i+1 runs (two pinned result buffers)"},
        "gpu_launches": launches * world,
        "gpu_launches_per_step_per_gpu": launches / args.steps,
        "gpu_launches_note": "kernels of libbackpack_b200.so inside the timed region, all ranks (per GPU and step: "
                             + ", ".join(f"{per_kernel[n + ':n']:g} {n}" for n in (
                                 "bp_fmha_fwd", "bp_linear_bias_act_fwd", "bp_linear_bias_residual_fwd", "bp_ln_fwd",
                                 "bp_ln_residual_fwd", "bp_sense_lse_fwd", "bp_sense_mix_fwd"))
                             + "); library GEMMs / gathers not counted",
        "roofline": roofline,
        "kernels": kernels,
        "model_mfu": {"achieved_tflops": value / world * model_flops_per_token / 1e12,
                      "frac_of_sustained_peak": value / world * model_flops_per_token / 1e12 / peak_tf},
        "cpu_baseline": cpu_base,
        "clocks": clocks.summary(),
    }
    if table_ms is not None:
        line["variants"] = {"sense_table": {
            "value": tokens / table_ms, "unit": "tokens/s", "ms_per_step": table_ms / args.steps * 1e3,
            "note": "inference-only: content model replaced by a gather from a precomputed (vocab, nv, d) table of "
                    "sense vectors (context-free by construction, backpack.py:258); NOT the headline value"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU")
    ap.add_argument("--seqlen", type=int, default=1024)
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--no-sense-table", dest="sense_table", action="store_false",
                    help="skip the secondary sense-vector-table variant")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")))
        return
    try:
        run_ours(args)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
