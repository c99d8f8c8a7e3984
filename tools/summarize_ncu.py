#!/usr/bin/env python
"""Summarises an .ncu-rep (or an ncu --csv launch list) into the short text files kept under profiles/.
   python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt
   python tools/summarize_ncu.py --launches gpurun_out/launches.csv profiles/r01_launches.txt"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def summarize_report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d.get('Kernel Name', '?')}\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"  {k:85s} {d[k]:>18s} {units[hdr.index(k)]}\n")
            try:
                tr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
                u = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                f.write(f"  traffic (dram read + write)                                                           {tr[0]} {u[0]} + {tr[1]} {u[1]}\n")
            except Exception:
                pass


def summarize_launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg[name][r[mi]] += v
        if r[mi] == "gpu__time_duration.sum":
            cnt[name] += 1
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list summary of {path}\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            t = a["gpu__time_duration.sum"]
            f.write(f"{k[:90]:90s} launches={cnt[k]:4d} total={t / 1e3:10.1f} us avg={t / cnt[k] / 1e3:9.1f} us share={t / tot * 100:5.1f}%\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        summarize_launches(sys.argv[2], sys.argv[3])
    else:
        summarize_report(sys.argv[1], sys.argv[2])
