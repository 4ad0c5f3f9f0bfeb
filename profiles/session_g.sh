# Round 2: racecheck of the small cases, ncu capture of the sense softmax-backward pass.  bash profiles/session_g.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python benchmarks/sanitizer_cases.py > $O/r02_sanitizer_racecheck.log 2>&1
echo "== racecheck rc=$?"; grep -E "FAIL|ERROR SUMMARY|RACECHECK SUMMARY|hazard|all finite" $O/r02_sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -20
bash profiles/ncu_kernels.sh r02g sense_softmax_bwd
