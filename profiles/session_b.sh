# Round 2, second half: backward kernels (stand-alone timings with comparators, sweep, ncu captures).  bash profiles/session_b.sh
set -u
O=gpurun_out; mkdir -p $O
python benchmarks/bench_kernels.py --which fmha_bwd,bwd_ops,xent > $O/r02w_kernel_bench_bwd.jsonl 2>$O/r02w_kb.err
for sh in 64,512,12,64 16,2048,12,64 8,4096,12,64 32,1024,6,128 8,4096,6,128; do
  python benchmarks/bench_kernels.py --which fmha_bwd --shape $sh --no-comparators >> $O/r02w_kernel_bench_bwd.jsonl 2>>$O/r02w_kb.err
done

cut -c1-200 $O/r02w_kernel_bench_bwd.jsonl
for n in 0 1; do timeout 200 ncu --set full --clock-control none --import-source on -k regex:fmha_bwd_kernel -s $((2+n)) -c 1 -f -o $O/prof_r02w_fmha_bwd_$n python -m benchmarks.run_fmha_bwd_once > $O/prof_r02w_fmha_bwd_$n.log 2>&1; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02w_bwd_launches.csv python -m benchmarks.run_fmha_bwd_once > /dev/null 2>&1
grep -E "bwd" $O/r02w_bwd_launches.csv | cut -d, -f5,12- | tail -3
