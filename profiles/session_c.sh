# Round 2, final verification: full GPU test suite, smoke, the default bench line.  bash profiles/session_c.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r02x_pytest.log 2>&1; tail -4 $O/r02x_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r02x_smoke.log 2>&1; tail -3 $O/r02x_smoke.log
timeout 900 python bench.py > $O/r02x_bench_n1.json 2> $O/r02x_bench.err; tail -2 $O/r02x_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02x_bench_n1.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "own share", round(d["own_kernel_share"], 3))
print("north_star", {k: (round(v["frac"], 3), round(v["ms_per_launch"], 3)) for k, v in d["north_star"].items()})
t = d["variants"]["training_step"]
print("training", round(t["value"]), "tok/s", round(t["ms_per_step"], 2), "ms; fmha_bwd", round(t["fmha_bwd"]["ms_per_call"], 3), "ms", round(t["fmha_bwd"]["frac"], 3))
print("sense_table", round(d["variants"]["sense_table"]["ms_per_step"], 2), "clocks", d["clocks"])
PY
