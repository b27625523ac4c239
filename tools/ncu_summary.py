#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.txt
  python tools/ncu_summary.py raw gpurun_out/prof_X.ncu-rep profiles/X_metrics.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_dim_x",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
           "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "sm__cycles_active.avg"]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s\n%-64s %6s %12s %10s %7s\n" % (src, "kernel", "n", "total_us", "avg_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %6d %12.1f %10.2f %7.3f\n" % (k[:64], n, t, t / n, t / tot))
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none ; source: %s\n" % src)
        for r in rows[2:]:
            f.write("\n== %s\n" % r[h.index("Kernel Name")][:100])
            for m in METRICS:
                if m in h:
                    f.write("  %-70s %s %s\n" % (m, r[h.index(m)], units[h.index(m)]))
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
