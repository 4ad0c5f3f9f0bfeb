#!/bin/bash
# GPU session A of a round (run through gpurun from the repo root):  bash profiles/session_a.sh <round>
# sanity of every kernel in isolated processes -> GPU test suite -> stand-alone kernel timings of the production
# library and of the tagged variant builds present -> config-5 sweep -> timeline traces -> the bench line.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/${R}_gpu.txt 2>&1
echo "== sanity"; timeout 900 python benchmarks/sanity.py --tag dbg > $O/${R}_sanity.log 2>&1; tail -16 $O/${R}_sanity.log | cut -c1-230
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -rf --tb=short > $O/${R}_pytest.log 2>&1; tail -25 $O/${R}_pytest.log | cut -c1-200
echo "== kernels"
for tag in ""; do
  [ -n "$tag" ] && [ ! -f backpacks_flash_attn_b200/libbackpack_b200_$tag.so ] && continue
  which="fmha,sense"; [ -z "$tag" ] && which="fmha,sense,ln,gemm"
  BP_LIB_TAG=$tag timeout 600 python benchmarks/bench_kernels.py --which $which > $O/${R}_kernels_${tag:-prod}.jsonl 2>&1
  echo "-- ${tag:-prod}"; cut -c1-150 $O/${R}_kernels_${tag:-prod}.jsonl | grep -v "^Traceback" | head -16
done
echo "== gemms"; timeout 600 python benchmarks/bench_kernels.py --which gemms > $O/${R}_gemms.jsonl 2>&1; cut -c1-200 $O/${R}_gemms.jsonl
echo "== gemm sustained"; timeout 300 python benchmarks/gemm_sustained.py > $O/${R}_gemm_sustained.jsonl 2>&1; cut -c1-330 $O/${R}_gemm_sustained.jsonl
echo "== linear policy A/B"; timeout 400 python benchmarks/linear_policy_ab.py > $O/${R}_linear_policy_ab.jsonl 2>&1; cut -c1-160 $O/${R}_linear_policy_ab.jsonl
echo "== decode"; timeout 400 python benchmarks/bench_decode.py > $O/${R}_decode.jsonl 2>&1; cut -c1-260 $O/${R}_decode.jsonl
echo "== sweep"; timeout 900 python benchmarks/sweep.py > $O/${R}_sweep.jsonl 2>&1; cut -c1-120 $O/${R}_sweep.jsonl | head -24
if [ -f backpacks_flash_attn_b200/libbackpack_b200_trace.so ]; then
  echo "== traces"
  BP_LIB_TAG=trace timeout 300 python benchmarks/trace_kernel.py fmha > $O/${R}_trace_fmha.txt 2>&1; tail -2 $O/${R}_trace_fmha.txt
  BP_LIB_TAG=trace timeout 300 python benchmarks/trace_kernel.py sense > $O/${R}_trace_sense.txt 2>&1; tail -2 $O/${R}_trace_sense.txt
  BP_LIB_TAG=trace timeout 300 python benchmarks/trace_kernel.py sense_table > $O/${R}_trace_sense_table.txt 2>&1; tail -2 $O/${R}_trace_sense_table.txt
fi
echo "== bench"; timeout 1200 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench.err; tail -3 $O/${R}_bench.err; cut -c1-300 $O/${R}_bench_n1.json; python - <<PY
import json
d=json.load(open("$O/${R}_bench_n1.json"))
print("linear backends", d["variants"].get("linear_backends",{}).get("ms_per_step")); print("library gemms", d.get("library_gemms")); print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "own share", d["own_kernel_share"], "table", d.get("variants",{}).get("sense_table",{}).get("ms_per_step"), "full", d.get("e2e_full_logits",{}).get("ms_per_step"))
for k,v in d["kernels"].items(): print(f"  {k:40s} n={v['launches_per_step']:5.1f} ms={v['ms_per_launch']:.4f} step={v['ms_per_step']:.3f} frac={v.get('frac')}")
PY
