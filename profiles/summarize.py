#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launch lists into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1a.csv  > profiles/r1a_launches.md
    python profiles/summarize.py full     gpurun_out/prof_r1a.ncu-rep  > profiles/r1a_ncu_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            try:
                agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    ours = sum(sum(v) for k, v in agg.items() if "ivc::" in k)
    print(f"# launch list {path}: {sum(len(v) for v in agg.values())} launches, {tot/1e6:.3f} ms total "
          f"(ivc:: kernels {ours/1e6:.3f} ms = {ours/tot*100:.1f} %; the rest is torch input synthesis / e2e statistics)\n")
    print("| kernel | launches | avg us | share of all | share of ivc:: |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        so = f"{sum(v)/ours*100:.1f} %" if "ivc::" in k else ""
        print(f"| `{k[:90]}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot*100:.1f} % | {so} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path}\n")
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        print(f"## {name}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
