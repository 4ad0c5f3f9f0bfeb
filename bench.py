#!/usr/bin/env python
"""Headline benchmark: forward tokens/s of Backpack-Small (d=768, 12 layers, 12 heads, k=16 senses), bf16,
seq 1024, batch 64 per GPU (BASELINE.json configs[2]; configs[3] = 8 GPUs x 64), synthetic ids and
name-seeded random weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--seqlen S]

N > 1 is launched by torchrun (one process per GPU); the batch is sharded with no data-path collective
(weak scaling: every rank runs B sequences).  Rank 0 prints ONE JSON line:
  value      whole-job tokens/s of the FULL forward (content model evaluated for every token, all 6.6 GB of logits
             written) with the ids resident in HBM: CUDA events, barrier + synchronize on both sides, max over ranks.
             The step is replayed as one CUDA graph (utils/graph.py; `--no-graph` launches from Python instead);
  e2e        the same forward through the public API with HOST inputs: per step a pinned-host -> device copy of
             the ids and a device -> host read of the last-position logits (what the reference's generation loop
             consumes, training/src/utils/generation.py:34-44); the host reads step i while step i+1 runs;
             `e2e_full_logits` is the same loop returning the API's WHOLE result (6.6 GB of logits per step over PCIe);
  roofline   the kernel of this library with the largest share of the step (all launches of that kernel);
  north_star the two kernels BASELINE.json's north_star names (fused attention, sense-mix) with both their tensor
             and HBM fractions;  kernels: every kernel of the library, the GEMM split by shape;
  own_kernel_share  time inside kernels of libbackpack_b200.so / step time;
  variants.sense_table  the serving configuration (`serving_config()`: sense vectors gathered inside the sense-mix
             kernel from a precomputed (vocab, nv, d) table instead of running the content model per token);
  variants.training_step  forward + backward with the backward operators of this library timed;
  cpu_baseline  the oracle port of the reference's pure-PyTorch path on the host cores (bounded sample).
Per-kernel durations come from CUDA events recorded on the launching stream around every C-ABI call during an eager
pass of the same K steps (events cannot be recorded inside a graph replay), against the measured peaks.
`--impl reference` times the CPU path alone (the reference's CUDA attention cannot run on sm_100, and its
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
    steps = min(args.steps, 30)     # ~0.3 s per (1,1024) forward on 16 cores: a bounded sample of the workload
    base, ms = cpu_reference_run(args.seqlen, steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True,
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
OWN_ENTRY_POINTS = ("bp_fmha_fwd", "bp_sense_lse_fwd", "bp_sense_mix_fwd", "bp_sense_mix_table_fwd",
                    "bp_linear_bias_act_fwd", "bp_linear_bias_residual_fwd", "bp_ln_residual_fwd", "bp_ln_fwd")


def timed_steps(fn, steps, parallel, dev):
    """K calls of fn bracketed by barrier + synchronize on both sides and one CUDA event pair; max over ranks (s)."""
    parallel.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    return parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)


def run_ours(args):
    from backpacks_flash_attn_b200 import _lib, parallel
    from backpacks_flash_attn_b200.models.backpack import BackpackLMHeadModel, flash_config
    from backpacks_flash_attn_b200.utils.graph import GraphedForward
    from backpacks_flash_attn_b200.utils.weights import name_seeded_

    rank, local_rank, world = parallel.init_distributed()
    if world != args.gpus and rank == 0:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.load().bp_check_device(), "bp_check_device", launched=False)

    B, S, K = args.batch, args.seqlen, args.steps
    cfg = flash_config(**{**SMALL, "n_positions": max(1024, S)})
    model = name_seeded_(BackpackLMHeadModel(cfg).eval()).to(dev, torch.bfloat16)
    parallel.assert_replicas_match(model)
    ids_host = parallel.shard_batch(make_ids(B * world, S), rank, world).contiguous().pin_memory()
    ids_dev = ids_host.to(dev)
    last_host = torch.empty((2, B, cfg.vocab_size), dtype=torch.bfloat16).pin_memory()   # double-buffered results
    copy_stream = torch.cuda.Stream(device=dev)

    def make_forward(use_graph):
        """resident(): forward on the ids already in HBM;  from_host(): same with a pinned-host -> device copy first."""
        if use_graph:
            g = GraphedForward(model, ids_dev)
            return (lambda: g()), (lambda: g(ids_host))
        return (lambda: model(ids_dev).logits), (lambda: model(ids_host.to(dev, non_blocking=True)).logits)

    def run_e2e(from_host, steps):
        """Serving loop through the public API with host buffers.  Every step copies its ids from pinned host
        memory and its last-position logits back to pinned host memory; the host waits for (consumes) the result
        of step i while step i+1 is already running, so the device never idles behind the host's launch loop."""
        pending, checksum, main = None, 0.0, torch.cuda.current_stream()
        for i in range(steps):
            logits = from_host()
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
        return checksum + float(last_host[pending[1], 0, 0])

    def time_e2e(from_host, steps):
        run_e2e(from_host, 2)
        parallel.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        run_e2e(from_host, steps)
        g1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        parallel.barrier()
        return parallel.max_over_ranks(max(g0.elapsed_time(g1) * 1e-3, wall), dev)

    with torch.inference_mode():
        for _ in range(args.warmup):
            model(ids_dev)
        torch.cuda.synchronize()
        resident, from_host = make_forward(args.graph)
        for _ in range(args.warmup):
            resident()
        # ---- timed region 1: inputs resident in HBM ----
        with ClockSampler(physical_gpu_index(local_rank)) as clocks:
            dt = timed_steps(resident, K, parallel, dev)
        tokens = parallel.sum_over_ranks(float(B * S * K), dev)

        # ---- kernel region: the same K steps launched from Python, every kernel of this library bracketed by CUDA
        #      events on its stream (the per-kernel durations behind `roofline` / `kernels` / `own_kernel_share`) ----
        launches0 = _lib.total_launches()
        gemm_key = lambda a: (int(a[5]), int(a[6]), int(a[7]))          # (n, k, activation) of bp_linear_bias_act_fwd
        timers = {n: _lib.KernelTimer(n, gemm_key if n == "bp_linear_bias_act_fwd" else None) for n in OWN_ENTRY_POINTS}
        from backpacks_flash_attn_b200.ops import fused_dense as FD
        lib_events = {}

        class _Bracket:
            """CUDA events around an F.linear call (a GEMM the default policy leaves to cuBLAS: the tied LM head)."""

            def __init__(self, tag, n, k):
                self.key = (n, k)

            def __enter__(self):
                self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.a.record()

            def __exit__(self, *exc):
                self.b.record()
                lib_events.setdefault(self.key, []).append((self.a, self.b))
                return False

        for t in timers.values():
            t.__enter__()
        FD._timing_hook = _Bracket
        dt_eager = timed_steps(lambda: model(ids_dev), K, parallel, dev)
        FD._timing_hook = None
        for t in reversed(list(timers.values())):
            t.__exit__(None, None, None)
        torch.cuda.synchronize()
        default_lib_gemms = {k: (len(v) / K, sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in lib_events.items()}
        launches = _lib.total_launches() - launches0          # per K steps; a graph replay launches the same kernels
        per_kernel = {n: (len(t.events) / K, t.mean_ms()) for n, t in timers.items()}      # (launches per step, ms)
        gemm_by_shape = {k: (n / K, ms / max(n, 1)) for k, (n, ms) in timers["bp_linear_bias_act_fwd"].by_key().items()}

        # ---- timed region 2: end to end through the public API with host buffers ----
        dt_e2e = time_e2e(from_host, K)

        # ---- variant: the plain linears on the library GEMM (F.linear -> cuBLAS, what the reference calls), with every
        #      library GEMM of an eager pass bracketed by events like the own kernels above ----
        lib_linear = None
        policy_ms = {}
        if args.library_linears:
            lib_events = {}
            default_backend = FD.get_linear_backend()
            for pname, policy in (("library", "library"), ("own", "own"), ("auto", "auto")):
                FD.set_linear_backend(policy)
                for _ in range(2):
                    model(ids_dev)
                l_resident, _ = make_forward(args.graph)
                for _ in range(2):
                    l_resident()
                policy_ms[pname] = timed_steps(l_resident, K, parallel, dev) / K * 1e3
                if pname == "library":
                    FD._timing_hook = _Bracket
                    kk = min(K, 20)
                    timed_steps(lambda: model(ids_dev), kk, parallel, dev)
                    FD._timing_hook = None
                    torch.cuda.synchronize()
                    lib_linear = {"per_shape": {k: (len(v) / kk, sum(a.elapsed_time(b) for a, b in v) / len(v))
                                                for k, v in lib_events.items()}}
                del l_resident
            FD.set_linear_backend(default_backend)

        # ---- the API's WHOLE result to the host (few steps: 6.6 GB over PCIe each) ----
        full_ms = None
        if args.full_logits_steps > 0 and world == 1:
            full_host = torch.empty((B, S, cfg.vocab_size), dtype=torch.bfloat16).pin_memory()
            for i in range(1 + args.full_logits_steps):
                if i == 1:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                full_host.copy_(from_host(), non_blocking=True)
                torch.cuda.synchronize()
            full_ms = (time.perf_counter() - t0) / args.full_logits_steps * 1e3
            del full_host

        # ---- variant: the serving configuration (sense vectors gathered inside the kernel from a table) ----
        table = None
        if args.sense_table:
            model.transformer.use_sense_table = True
            model.transformer.build_sense_table()
            for _ in range(2):
                model(ids_dev)
            t_resident, t_from_host = make_forward(args.graph)
            for _ in range(2):
                t_resident()
            t_dt = timed_steps(t_resident, K, parallel, dev)
            tt = {n: _lib.KernelTimer(n) for n in ("bp_sense_lse_fwd", "bp_sense_mix_table_fwd")}
            for t in tt.values():
                t.__enter__()
            timed_steps(lambda: model(ids_dev), min(K, 20), parallel, dev)
            for t in reversed(list(tt.values())):
                t.__exit__(None, None, None)
            t_e2e = time_e2e(t_from_host, K)
            table = {"dt": t_dt, "e2e": t_e2e, "lse_ms": tt["bp_sense_lse_fwd"].mean_ms(),
                     "mix_ms": tt["bp_sense_mix_table_fwd"].mean_ms()}
            model.transformer.use_sense_table = False
            model.transformer.drop_sense_table()

        # ---- variant: evaluation (per-token cross-entropy) through the LM head with the softmax statistics in its
        #      epilogue: the forward without the 6.6 GB of logits (bp_lm_head_stats_fwd) ----
        fused_loss = None
        if args.fused_loss_steps > 0:
            labels = torch.randint(0, cfg.vocab_size - 7, (B, S), device=dev, generator=torch.Generator(device=dev).manual_seed(7))
            loss_timer = _lib.KernelTimer("bp_lm_head_stats_fwd")
            for _ in range(3):
                model.loss(ids_dev, labels)
            kk = min(K, args.fused_loss_steps)
            with loss_timer:
                fl_dt = timed_steps(lambda: model.loss(ids_dev, labels), kk, parallel, dev)
            fused_loss = {"dt": fl_dt, "steps": kk, "stats_ms": loss_timer.mean_ms()}

    # ---- variant: a training step (forward + backward with a next-token cross-entropy), SURVEY.md §8f rank 4.  Outside
    #      inference mode, same batch as the headline (the activations of 12 layers are kept: ~60 GB), dropouts 0 ----
    training = None
    training_error = None
    if args.training_steps > 0:
        try:
            tb = min(B, args.training_batch)
            tcfg = flash_config(**{**SMALL, "n_positions": max(1024, S), "resid_pdrop": 0.0, "embd_pdrop": 0.0,
                                   "attn_pdrop": 0.0})
            tmodel = BackpackLMHeadModel(tcfg).to(dev, torch.bfloat16).train()
            tmodel.load_state_dict(model.state_dict())
            del model
            torch.cuda.empty_cache()
            tids = ids_dev[:tb]

            from backpacks_flash_attn_b200.losses.cross_entropy import CrossEntropyLoss
            ce = CrossEntropyLoss(inplace_backward=True)     # bp_xentropy_fwd / _bwd: one pass each, in place over the logits
            # next-token targets for every position; the last one of a sequence is ignored (no copy of the logits)
            tlabels = torch.cat([tids[:, 1:], torch.full_like(tids[:, :1], -100)], dim=1).reshape(-1)

            def train_step():
                tmodel.zero_grad(set_to_none=True)
                logits = tmodel(tids).logits
                loss = ce(logits.view(-1, logits.shape[-1]), tlabels)
                loss.backward()
                parallel.allreduce_gradients(tmodel)   # data-parallel training: the one exchange step (no-op at N = 1)
                return loss

            for _ in range(2):
                train_step()
            bwd_names = ("bp_fmha_fwd", "bp_fmha_bwd", "bp_ln_residual_fwd", "bp_ln_residual_bwd", "bp_bias_act_bwd",
                         "bp_linear_bias_act_fwd", "bp_linear_bias_act_aux_fwd", "bp_sense_lse_fwd", "bp_sense_mix_fwd",
                         "bp_sense_softmax_bwd", "bp_xentropy_fwd",
                         "bp_xentropy_bwd")
            tt = {n: _lib.KernelTimer(n) for n in bwd_names}
            for t in tt.values():
                t.__enter__()
            tr_dt = timed_steps(train_step, args.training_steps, parallel, dev)
            for t in reversed(list(tt.values())):
                t.__exit__(None, None, None)
            training = {"dt": tr_dt, "steps": args.training_steps, "batch": tb,
                        "kernels": {n: (len(t.events) / args.training_steps, t.mean_ms()) for n, t in tt.items() if t.events}}
            # the same step captured once and replayed as ONE CUDA graph (single GPU: no collective inside the capture)
            if world == 1:
                try:
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        train_step()
                    torch.cuda.current_stream().wait_stream(side)
                    tmodel.zero_grad(set_to_none=True)
                    tg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(tg):
                        graph_loss = train_step()
                    for _ in range(2):
                        tg.replay()
                    training["graph_dt"] = timed_steps(tg.replay, args.training_steps, parallel, dev)
                    training["graph_loss"] = float(graph_loss.detach())
                    del tg
                except Exception as exc:
                    training["graph_error"] = f"{type(exc).__name__}: {exc}"[:200]
            # the reference's training configuration: residual / embedding / attention dropout 0.1, every mask drawn
            # INSIDE the LayerNorm and attention kernels (bp_ln_residual_*_dropout, bp_fmha_*_dropout)
            try:
                for m in tmodel.modules():
                    if isinstance(m, torch.nn.Dropout):
                        m.p = 0.1
                    if hasattr(m, "dropout_p"):
                        m.dropout_p = 0.1
                before = dict(_lib.launch_counts)
                for _ in range(2):
                    train_step()
                training["dropout_dt"] = timed_steps(train_step, args.training_steps, parallel, dev)
                training["dropout_launches"] = {n: _lib.launch_counts.get(n, 0) - before.get(n, 0) for n in
                                                ("bp_ln_residual_fwd_dropout", "bp_ln_residual_bwd_dropout",
                                                 "bp_fmha_fwd_dropout", "bp_fmha_bwd")}
            except Exception as exc:
                training["dropout_error"] = f"{type(exc).__name__}: {exc}"[:200]
            del tmodel
        except Exception as exc:   # a variant must never take the headline line down with it
            training, training_error = None, f"{type(exc).__name__}: {exc}"[:300]

    if rank != 0:
        return
    peaks = load_peaks()
    peak_tf, peak_hbm = peaks["tf_sustained"], peaks["hbm_gbs"]
    h, dh, nv, d = cfg.n_head, cfg.n_embd // cfg.n_head, cfg.num_content_vectors, cfg.n_embd
    M = B * S
    step_ms = dt / K * 1e3
    full_shape = (B, S) == (64, 1024)

    def tensor_rec(name, flops, nbytes, n_per_step, ms, extra=None):
        t = ms * 1e-3
        a = flops / t / 1e12
        rec = {"bound": "tensor", "achieved": a, "peak": peak_tf, "unit": "TFLOP/s", "frac": a / peak_tf, "kernel": name,
               "launches_per_step": n_per_step, "ms_per_launch": ms, "ms_per_step": ms * n_per_step,
               "algorithmic_gflop_per_launch": flops / 1e9, "algorithmic_mb_per_launch": nbytes / 1e6,
               "hbm_gbs": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / peak_hbm}
        rec.update(extra or {})
        return rec

    def hbm_rec(name, nbytes, n_per_step, ms):
        a = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak_hbm, "unit": "GB/s", "frac": a / peak_hbm, "kernel": name,
                "launches_per_step": n_per_step, "ms_per_launch": ms, "ms_per_step": ms * n_per_step,
                "algorithmic_mb_per_launch": nbytes / 1e6}

    # algorithmic work per launch
    kernels = {}
    n_f, ms_f = per_kernel["bp_fmha_fwd"]
    kernels["fmha"] = tensor_rec("fmha_fwd_kernel<64,bf16> (bp_fmha_fwd)", 4 * B * h * S * S * dh / 2,
                                 4 * B * S * h * dh * 2 + 4 * B * h * S, n_f, ms_f,
                                 {"traffic": ncu_traffic("fmha_fwd_kernel") if full_shape else None})
    ms_lse, ms_mix = per_kernel["bp_sense_lse_fwd"][1], per_kernel["bp_sense_mix_fwd"][1]
    mix_flops, mix_bytes = B * S * S * d * (1 + nv), (2 * B * S * d + nv * B * S * d + B * S * d) * 2
    kernels["sense_mix"] = tensor_rec("sense_lse_kernel + sense_mix_kernel (bp_sense_lse_fwd, bp_sense_mix_fwd)",
                                      mix_flops, mix_bytes, 1, ms_lse + ms_mix,
                                      {"ms_lse": ms_lse, "ms_mix": ms_mix,
                                       "traffic": ncu_traffic("sense_mix_kernel") if full_shape else None})
    gemm_names = {(3 * d, d, 0): "Wqkv", (d, d, 0): "out_proj", (d, 4 * d, 0): "fc2", (4 * d, d, 1): "fc1 + bias + tanh-GELU",
                  (2 * d, d, 0): "ctx Wqkv", (nv * d, 4 * d, 0): "content final_mlp.fc2", (cfg.vocab_size, d, 0): "tied LM head"}
    gemm_total_ms = gemm_total_flops = gemm_total_bytes = gemm_total_n = 0.0
    for (n, k, act), (cnt, ms) in sorted(gemm_by_shape.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
        flops, nbytes = 2.0 * M * n * k, (M * k + n * k + M * n) * 2.0
        name = gemm_names.get((n, k, act), f"n{n} k{k} act{act}")
        kernels[f"gemm[{name}]"] = tensor_rec(f"gemm_bias_act_pair_kernel<bf16> (bp_linear_bias_act_fwd): {name}, "
                                              f"(m, n, k) = ({M}, {n}, {k})", flops, nbytes, cnt, ms)
        gemm_total_ms += cnt * ms
        gemm_total_flops += cnt * flops
        gemm_total_bytes += cnt * nbytes
        gemm_total_n += cnt
    n_ln, ms_ln = per_kernel["bp_ln_residual_fwd"]
    kernels["ln_residual"] = hbm_rec("ln_residual_fwd_kernel (bp_ln_residual_fwd)", M * d * (2 + 4) * 2, n_ln, ms_ln)
    kernels["ln_residual"]["traffic"] = ncu_traffic("ln_residual_fwd_kernel") if full_shape else None
    for nm in ("bp_linear_bias_residual_fwd", "bp_ln_fwd"):
        if per_kernel[nm][0]:
            kernels[nm] = {"launches_per_step": per_kernel[nm][0], "ms_per_launch": per_kernel[nm][1],
                           "ms_per_step": per_kernel[nm][0] * per_kernel[nm][1]}
    # all launches of the GEMM kernel as one record (it is ONE kernel; the shapes above are its breakdown)
    gemm_all = tensor_rec("gemm_bias_act_pair_kernel<bf16> (bp_linear_bias_act_fwd), all launches of a step: every linear "
                          "of the model incl. fc1+GELU, the content model's 3072->12288 projection and the tied LM head",
                          gemm_total_flops / max(gemm_total_n, 1), gemm_total_bytes / max(gemm_total_n, 1), gemm_total_n,
                          gemm_total_ms / max(gemm_total_n, 1),
                          {"traffic": None, "traffic_note": "per-shape ncu captures are summarised under profiles/"})
    own_ms = sum(v["ms_per_step"] for kname, v in kernels.items())
    for v in kernels.values():
        v.setdefault("traffic_unit", "DRAM bytes per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum (profiles/)")
    candidates = {"gemm": gemm_all, "fmha": kernels["fmha"], "sense_mix": kernels["sense_mix"],
                  "ln_residual": kernels["ln_residual"]}
    dominant = max(candidates, key=lambda n: candidates[n]["ms_per_step"])
    roofline = dict(candidates[dominant])
    roofline["dominant_of"] = {n: round(v["ms_per_step"], 3) for n, v in candidates.items()}
    roofline["peak_source"] = (f"MEASURED_PEAKS.json ({peaks['source']}): bf16_tflops_sustained for tensor-bound kernels "
                               "(timed inside a long step), hbm_gbs for HBM-bound ones")
    north_star = {k: {kk: kernels[k][kk] for kk in ("kernel", "frac", "achieved", "unit", "hbm_frac", "hbm_gbs",
                                                    "ms_per_launch", "launches_per_step", "ms_per_step")}
                  for k in ("fmha", "sense_mix")}
    model_flops_per_token = 371.3e6   # SURVEY.md §8d, s = 1024
    cpu_base, _ = cpu_reference_run(S, steps=3, warmup=1)
    value = tokens / dt
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Backpack-Small forward (BASELINE configs[2]; configs[3] when n_gpus=8): full model, "
                               "content model evaluated per token, all logits written",
                   "batch_per_gpu": B, "global_batch": B * world, "seq_len": S, "d_model": d, "n_layer": cfg.n_layer,
                   "n_head": h, "num_content_vectors": nv, "vocab": cfg.vocab_size, "parallelism": f"dp{world}",
                   "weights": "name-seeded random (SURVEY.md §8c recipe), bf16",
                   "execution": ("one CUDA graph per step (utils/graph.py); the eager launch sequence of the same "
                                 f"kernels, with an event pair recorded around each of them, takes {dt_eager / K * 1e3:.3f} ms per step"
                                 if args.graph else "eager launches from Python"),
                   "kernel_timings": "per-kernel CUDA events around an eager pass of the same K steps on the same inputs",
                   "l2": "no flush: each step streams > 10 GB of activations (6.6 GB logits, 1.6 GB sense vectors) "
                         "through a 126 MB L2"},
        "e2e": {"value": tokens / dt_e2e, "unit": "tokens/s", "h2d_bytes_per_step": ids_host.numel() * 8 * world,
                "d2h_bytes_per_step": last_host[0].numel() * 2 * world, "ms_per_step": dt_e2e / K * 1e3,
                "result": "last-position logits (batch, vocab) bf16 copied to pinned host memory every step; the host reads "
                          "step i while step i+1 runs (two pinned result buffers)"},
        "gpu_launches": launches * world,
        "gpu_launches_per_step_per_gpu": launches / K,
        "gpu_launches_note": "kernels of libbackpack_b200.so inside the timed region, all ranks (per GPU and step: "
                             + ", ".join(f"{per_kernel[n][0]:g} {n}" for n in OWN_ENTRY_POINTS if per_kernel[n][0])
                             + "); embedding gathers / element-wise ATen kernels not counted",
        "own_kernel_share": own_ms / step_ms,
        "own_kernel_ms_per_step": own_ms,
        "library_gemms": {f"n{n} k{k}": {"kernel": "cuBLAS via F.linear (default linear policy, ops/fused_dense.py: the tied "
                                                   "LM head stays on the library)",
                                          "launches_per_step": c, "ms_per_launch": ms, "ms_per_step": c * ms,
                                          "frac": 2.0 * M * n * k / (ms * 1e-3) / 1e12 / peak_tf}
                          for (n, k), (c, ms) in default_lib_gemms.items()},
        "roofline": roofline,
        "north_star": north_star,
        "kernels": kernels,
        "model_mfu": {"achieved_tflops": value / world * model_flops_per_token / 1e12,
                      "frac_of_sustained_peak": value / world * model_flops_per_token / 1e12 / peak_tf},
        "cpu_baseline": cpu_base,
        "clocks": clocks.summary(),
    }
    if full_ms is not None:
        line["e2e_full_logits"] = {
            "value": B * S / (full_ms * 1e-3), "unit": "tokens/s", "ms_per_step": full_ms,
            "h2d_bytes_per_step": ids_host.numel() * 8, "d2h_bytes_per_step": B * S * cfg.vocab_size * 2,
            "steps": args.full_logits_steps,
            "note": "the API's whole result (batch, seq, vocab) bf16 copied to pinned host memory every step: bound by the "
                    "host link (6.6 GB per step), not by the GPU; no caller of the reference consumes it on the host"}
    line["variants"] = {}
    if lib_linear is not None:
        line["variants"]["linear_backends"] = {
            "ms_per_step": policy_ms,
            "per_shape_ms": {f"n{n} k{k}": {"launches_per_step": c, "cublas_ms": ms,
                                             "own_ms": gemm_by_shape.get((n, k, 0), (0, None))[1]}
                             for (n, k), (c, ms) in lib_linear["per_shape"].items()},
            "note": "the same forward (one CUDA graph per step) with the plain linears (Wqkv / out_proj / fc2 / content "
                    "projection / LM head) on F.linear = cuBLAS, as the reference's FusedDense.forward does "
                    "(flash_attn/ops/fused_dense.py:52,112), on this library's GEMM, or mixed; fc1+GELU is always fused. "
                    "per_shape_ms: CUDA events around every GEMM of an eager pass"}
    if table is not None:
        line["variants"]["sense_table"] = {
            "value": tokens / table["dt"], "unit": "tokens/s", "ms_per_step": table["dt"] / K * 1e3,
            "e2e": {"value": tokens / table["e2e"], "unit": "tokens/s", "ms_per_step": table["e2e"] / K * 1e3},
            "sense_mix_ms": {"lse": table["lse_ms"], "mix_table": table["mix_ms"]},
            "sense_mix_frac_of_tensor_peak": mix_flops / ((table["lse_ms"] + table["mix_ms"]) * 1e-3) / 1e12 / peak_tf,
            "note": "serving_config(): identical logits; the content model (24 % of the model's FLOPs) runs once per "
                    "vocabulary item instead of once per token and bp_sense_mix_table_fwd gathers the rows inside the "
                    "kernel (no (b, s, nv, d) tensor in HBM).  Reported as a variant: `value` above is the full forward"}
    if fused_loss is not None:
        fl_ms = fused_loss["dt"] / fused_loss["steps"] * 1e3
        line["variants"]["fused_loss"] = {
            "value": B * S * world / (fl_ms * 1e-3), "unit": "tokens/s", "ms_per_step": fl_ms, "steps": fused_loss["steps"],
            "lm_head_stats_ms": fused_loss["stats_ms"],
            "lm_head_stats_frac_of_tensor_peak": 2.0 * M * cfg.vocab_size * d / (fused_loss["stats_ms"] * 1e-3) / 1e12 / peak_tf,
            "note": "BackpackLMHeadModel.loss(ids, labels): the same forward with the tied LM head run by "
                    "bp_lm_head_stats_fwd (log-sum-exp, arg-max and target logit per row in the GEMM epilogue; no "
                    "(b, s, vocab) tensor), eager launches.  What perplexity evaluation needs; a different result than "
                    "`value` (per-token loss instead of logits), hence a variant"}
    if training_error is not None:
        line["variants"]["training_step"] = {"unavailable": training_error}
    if training is not None:
        tb, tms = training["batch"], training["dt"] / training["steps"] * 1e3
        tk = training["kernels"]
        fb_flops = 2.5 * 4 * tb * h * S * S * dh / 2
        line["variants"]["training_step"] = {
            "value": tb * S * world / (tms * 1e-3), "unit": "tokens/s", "ms_per_step": tms, "steps": training["steps"],
            "batch_per_gpu": tb,
            "cuda_graph": ({"ms_per_step": training["graph_dt"] / training["steps"] * 1e3,
                            "value": tb * S * world / (training["graph_dt"] / training["steps"]), "unit": "tokens/s",
                            "note": "forward + backward captured once, replayed as one CUDA graph"}
                           if "graph_dt" in training else {"unavailable": training.get("graph_error", "multi-GPU run")}),
            "with_dropout": ({"ms_per_step": training["dropout_dt"] / training["steps"] * 1e3,
                              "value": tb * S * world / (training["dropout_dt"] / training["steps"]), "unit": "tokens/s",
                              "launches": training["dropout_launches"],
                              "note": "the reference's training configuration (residual / embedding / attention dropout "
                                      "0.1): every mask is drawn inside the LayerNorm and attention kernels from a seed"}
                             if "dropout_dt" in training else {"unavailable": training.get("dropout_error", "not run")}),
            "own_kernel_share": sum(c * ms for c, ms in tk.values()) / tms,
            "kernels": {n: {"calls_per_step": c, "ms_per_call": ms, "ms_per_step": c * ms} for n, (c, ms) in tk.items()},
            "fmha_bwd": {"kernel": "bp_fmha_bwd: bwd_stats_kernel + fmha_bwd_kernel<64, keys own> + fmha_bwd_kernel<64, "
                                   "queries own> (three launches per call)",
                         "ms_per_call": tk["bp_fmha_bwd"][1],
                         "tflops": fb_flops / (tk["bp_fmha_bwd"][1] * 1e-3) / 1e12,
                         "frac": fb_flops / (tk["bp_fmha_bwd"][1] * 1e-3) / 1e12 / peak_tf,
                         "flops_convention": "2.5 x the forward (five tile products against two); the two kernels execute "
                                             "seven (S and dP are recomputed in both)"},
            "ln_residual_bwd": {"ms_per_call": tk["bp_ln_residual_bwd"][1],
                                "hbm_frac": tb * S * d * 16 / (tk["bp_ln_residual_bwd"][1] * 1e-3) / 1e9 / peak_hbm},
            "note": "forward + backward of the same model in train mode (dropouts 0), next-token cross-entropy by "
                    "bp_xentropy_fwd / _bwd in place over the bf16 logits, eager launches, no optimizer step: attention "
                    "backward = bp_fmha_bwd, LayerNorm backward = "
                    "bp_ln_residual_bwd, dgelu + bias gradients = bp_bias_act_bwd (the fc1 GEMM stores its pre-activation: no recompute), dgrad GEMMs = this library's GEMM, wgrad "
                    "GEMMs = cuBLAS through PyTorch, sense-mix backward = batched cuBLAS GEMMs around bp_sense_softmax_bwd; at N > 1 the batch is sharded "
                    "and the gradients are averaged with bucketed NCCL all-reduces (parallel.allreduce_gradients).  A "
                    "variant: the headline metric is the forward"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU")
    ap.add_argument("--seqlen", type=int, default=1024)
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--no-sense-table", dest="sense_table", action="store_false",
                    help="skip the serving-configuration variant")
    ap.add_argument("--no-library-linears", dest="library_linears", action="store_false",
                    help="skip the variant that runs the plain linears on cuBLAS")
    ap.add_argument("--fused-loss-steps", type=int, default=30,
                    help="steps of the evaluation-loss variant (0 = skip)")
    ap.add_argument("--training-steps", type=int, default=5, help="steps of the training-step variant (0 = skip)")
    ap.add_argument("--training-batch", type=int, default=64, help="sequences per GPU in the training-step variant")
    ap.add_argument("--full-logits-steps", type=int, default=3,
                    help="steps of the whole-logits-to-host loop (0 = skip; single GPU only)")
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
