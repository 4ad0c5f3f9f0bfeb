# Round 2: compute-sanitizer over small forward + backward cases of every kernel.  bash profiles/session_d.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 120 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_plain.log 2>&1; tail -3 $O/r02_sanitizer_plain.log
for tool in memcheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "PASS|FAIL|ERROR SUMMARY|Invalid|Barrier error|all finite" $O/r02_sanitizer_$tool.log | tail -24
done
