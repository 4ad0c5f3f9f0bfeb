#!/bin/bash
# Collect the ncu evidence of one round on the GPU box (run through gpurun from the repo root):
#   bash profiles/collect.sh r02
# writes gpurun_out/launches_<round>.csv (per-launch durations of `bench.py --steps 2 --warmup 3`) and one
# `--set full` report per kernel of libbackpack_b200.so (gpurun_out/prof_<round>_<kernel>.ncu-rep).
# Summarise here with profiles/summarize_ncu.py / profiles/summarize_launches.py.
set -u
R=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-sense-table --no-graph --no-library-linears --full-logits-steps 0 --fused-loss-steps 0 > gpurun_out/launches_$R.log 2>&1
# kernel:launches during the 3 warm-up steps (skipped, so that the profiled launch is one of the timed step)
# (per step: 12 fmha, 1 + 1 sense, 53 own GEMMs -- the first of a step is a Wqkv, +1 = out_proj, +2 = fc1+GELU -- 28 LN)
for ks in fmha_fwd_kernel:36 sense_mix_kernel:3 sense_lse_kernel:3 gemm_bias_act_pair_kernel:161 ln_residual_fwd_kernel:89; do
  k=${ks%%:*}; skip=${ks##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${R}_$k \
      python bench.py --steps 1 --warmup 3 --no-sense-table --no-graph --no-library-linears --full-logits-steps 0 --fused-loss-steps 0 > gpurun_out/prof_${R}_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
