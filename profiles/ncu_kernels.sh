#!/bin/bash
# ncu --set full capture of one launch of each hot kernel at its BASELINE shape (stand-alone, benchmarks/run_one.py):
#   bash profiles/ncu_kernels.sh <round> [which...]      -> gpurun_out/prof_<round>_<which>.ncu-rep
set -u
R=${1:-r02}; shift
O=gpurun_out; mkdir -p $O
for w in "${@:-fmha sense sense_table}"; do for which in $w; do
  case $which in
    fmha|fmha128) k=fmha_fwd_kernel;; sense|sense_table) k=sense_mix_kernel;; lse) k=sense_lse_kernel; which_run=sense;;
    gemm) k=gemm_bias_act_pair_kernel;; ln) k=ln_residual_fwd_kernel;;
    dec_attn) k=decode_attn_kernel;; dec_sense) k=sense_mix_decode_kernel;; sense_softmax_bwd) k=sense_softmax_bwd_kernel;;
  esac
  run=$which; [ $which = lse ] && run=sense
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/prof_${R}_$which \
      python benchmarks/run_one.py $run 3 > $O/prof_${R}_$which.log 2>&1
  tail -2 $O/prof_${R}_$which.log
done; done
ls -la $O/*${R}*.ncu-rep
