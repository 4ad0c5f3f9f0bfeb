# Round 2: sense-mix backward (bp_sense_softmax_bwd): tests, sanitizer, timing, training step.  bash profiles/session_e.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_sense_mix_bwd_gpu.py tests/test_training_gpu.py -q -x --timeout 300 -s > $O/r02e_pytest.log 2>&1; tail -5 $O/r02e_pytest.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -2 $O/r02_sanitizer_memcheck.log
timeout 300 python benchmarks/bench_kernels.py --which sense_bwd --iters 12 > $O/r02e_sense_bwd.jsonl 2> $O/r02e_sense_bwd.err; cat $O/r02e_sense_bwd.jsonl; tail -3 $O/r02e_sense_bwd.err
timeout 300 python benchmarks/profile_training_step.py > $O/r02e_training_profile.txt 2>&1; head -40 $O/r02e_training_profile.txt
