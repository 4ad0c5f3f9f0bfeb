"""Turn `ncu --set full` reports (gpurun_out/prof_<round>_<label>.ncu-rep) into the committed text summaries
profiles/<round>_ncu_<label>.txt and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py).

    python profiles/summarize_ncu.py gpurun_out/prof_r02_*.ncu-rep

<label> is the kernel name for captures taken inside the bench step (profiles/collect.sh) or the case name of a
stand-alone capture (profiles/ncu_kernels.sh: sense_table, dec_attn, dec_sense, ...).  A trailing letter of the round
tag (r02p -> r02) names the GPU session and is dropped.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.avg|sm__cycles_elapsed\.avg\.per_second|"
    r"sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_alu\.avg\.pct_of_peak_sustained_active|sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|"
    r"launch__grid_size|launch__block_size|launch__cluster_size|smsp__inst_executed\.sum|lts__t_sector_hit_rate\.pct|l1tex__m_xbar2l1tex_read_bytes\.sum|"
    r"l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio)$|tensor_cycles_active")


def main(paths):
    traffic_path = os.path.join(HERE, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("no data in", path)
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = dict(zip(hdr, zip(units, vals)))
        kname = rec.get("Kernel Name", ("", "?"))[1]
        short = re.sub(r"^void |<.*", "", kname).split("::")[-1]
        m = re.match(r"prof_(r\d+)[a-z]?_(.+)\.ncu-rep$", os.path.basename(path))
        rnd, label = (m.group(1), m.group(2)) if m else ("rXX", short)
        in_bench = label.endswith("_kernel")
        lines = [("ncu --set full --clock-control none, one launch inside `python bench.py --steps 1 --warmup 3 --no-graph` "
                  "(profiles/collect.sh)" if in_bench else
                  f"ncu --set full --clock-control none, one stand-alone launch: `python benchmarks/run_one.py {label}` "
                  "(profiles/ncu_kernels.sh)"),
                 f"kernel: {kname}", f"report: {os.path.basename(path)} (not committed; regenerate with the command above)", ""]
        for h in hdr:
            if KEEP.search(h):
                u, v = rec[h]
                lines.append(f"{h:100s} {u:16s} {v}")

        def num(key):
            u, v = rec[key]
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        try:
            tot = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            traffic[short if in_bench else label] = {"dram_bytes_per_launch": tot, "dram_read": num("dram__bytes_read.sum"),
                              "dram_write": num("dram__bytes_write.sum"), "source": os.path.basename(path)}
            lines.append(f"\nDRAM traffic per launch: {tot / 1e6:.1f} MB")
        except Exception as e:  # noqa
            lines.append(f"traffic unavailable: {e}")
        with open(os.path.join(HERE, f"{rnd}_ncu_{label}.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", f"{rnd}_ncu_{label}.txt")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
