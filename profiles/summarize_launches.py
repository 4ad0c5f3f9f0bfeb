"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches per step, ms per
step, share of the step.  Kernels of libbackpack_b200.so are marked with *.

    python profiles/summarize_launches.py gpurun_out/launches_r02.csv STEPS > profiles/r02_launch_list_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

OURS = ("fmha::", "sense::", "gemm::", "ln::", "rotary::", "decode::", "fmha_bwd::", "lnb::", "bab::", "xent::", "smb::")


def main(path, steps, title=None, top=14):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for k, v in rows:
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    fm = [k for k in cnt if "fmha_fwd_kernel" in k]
    if fm:   # 12 attention launches per forward pass: count the passes instead of trusting the argument
        steps = max(1, round(sum(cnt[k] for k in fm) / 12))
    if title:
        print(f"ncu --metrics gpu__time_duration.sum --clock-control none, `{title}` ({steps} steps captured; per-launch "
              f"times are cold-cache and serialised: compare SHARES)")
    else:
        print(f"ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3 --no-sense-table --no-graph --no-library-linears --full-logits-steps 0 --fused-loss-steps 0` (profiles/collect.sh) "
              f"({steps} forward passes captured; per-launch times are cold-cache and serialised: compare SHARES)")
    ours = sum(v for k, v in tot.items() if any(o in k for o in OURS))
    for k in sorted(tot, key=lambda k: -tot[k])[:top]:
        mark = "*" if any(o in k for o in OURS) else " "
        short = re.sub(r"\(.*", "", k)[:90]
        print(f"{mark} {short:92s} n/step {cnt[k] / steps:6.1f}  ms/step {tot[k] / steps:8.3f}  share {100 * tot[k] / total:5.1f}%")
    print(f"total ms/step under ncu {total / steps:.3f};  kernels of libbackpack_b200.so (*) {ours / steps:.3f} ms = "
          f"{100 * ours / total:.1f}% of the step")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 9, sys.argv[3] if len(sys.argv) > 3 else None,
         int(sys.argv[4]) if len(sys.argv) > 4 else 14)
