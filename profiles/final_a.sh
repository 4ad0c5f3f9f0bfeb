set -u
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
python benchmarks/sweep_sense.py > gpurun_out/sweep_sense.jsonl 2>&1; head -4 gpurun_out/sweep_sense.jsonl | cut -c1-110
python benchmarks/bench_kernels.py --which sense | cut -c1-110
for ks in sense_lse_kernel:3 ln_residual_fwd_kernel:89; do
  k=${ks%%:*}; skip=${ks##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_r01_$k \
      python bench.py --steps 1 --warmup 3 --no-sense-table --no-graph > gpurun_out/prof_r01_$k.log 2>&1
done
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_n1.json
