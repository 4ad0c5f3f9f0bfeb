# Round 2: residual dropout inside the LayerNorm kernels: tests, sanitizer, timings.  bash profiles/session_h.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ln_dropout_gpu.py tests/test_ln_rotary_gpu.py tests/test_bwd_ops_gpu.py tests/test_training_gpu.py tests/test_fmha_dropout_gpu.py -q -x --timeout 300 > $O/r02h_pytest.log 2>&1; tail -15 $O/r02h_pytest.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -2 $O/r02_sanitizer_memcheck.log
timeout 300 python benchmarks/bench_kernels.py --which bwd_ops --iters 12 > $O/r02h_bwd_ops.jsonl 2> $O/r02h_bwd_ops.err; cut -c1-175 $O/r02h_bwd_ops.jsonl; tail -3 $O/r02h_bwd_ops.err
