# Round 2: pre-activation output of the forward GEMM (bp_linear_bias_act_aux_fwd): tests, sanitizer, GEMM timing, training.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_bwd_ops_gpu.py tests/test_fused_dense_gpu.py tests/test_training_gpu.py -q -x --timeout 300 > $O/r02k_pytest.log 2>&1; tail -4 $O/r02k_pytest.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -2 $O/r02_sanitizer_memcheck.log
timeout 300 python - <<'PY'
import torch, sys, json
sys.path.insert(0, ".")
from benchmarks.bench_kernels import time_fn
from backpacks_flash_attn_b200.ops.fused_dense import _linear_bias_act_aux, linear_bias_act
x = torch.randn(65536, 768, device="cuda").bfloat16()
w = (torch.randn(3072, 768, device="cuda") * 768 ** -0.5).bfloat16()
b = torch.randn(3072, device="cuda").bfloat16()
t1, _ = time_fn(lambda i: linear_bias_act(x, w, b, "gelu_tanh"), 1, 20)
t2, _ = time_fn(lambda i: _linear_bias_act_aux(x, w, b, "gelu_tanh"), 1, 20)
t3, _ = time_fn(lambda i: linear_bias_act(x, w, b, "none"), 1, 20)
print(json.dumps({"kernel": "fc1 + GELU 65536 x 3072 x 768 bf16", "ms_plain": t1 * 1e3, "ms_with_pre_activation_output": t2 * 1e3,
                  "ms_recompute_gemm_it_replaces": t3 * 1e3}))
PY
timeout 600 python bench.py --steps 20 > $O/r02k_bench.json 2> $O/r02k_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench.json")); t = d["variants"]["training_step"]
print("fwd", round(d["ms_per_step"], 2), "training", round(t["ms_per_step"], 2), "graph", t["cuda_graph"].get("ms_per_step"), "dropout", t["with_dropout"].get("ms_per_step"), "share", round(t["own_kernel_share"], 3), d["clocks"])
PY
