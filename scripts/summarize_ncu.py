#!/usr/bin/env python
"""Summarise ncu output for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum)
-> per-kernel totals and shares; (2) a --set full .ncu-rep -> the handful of counters the design
argues from.  Usage:
    python scripts/summarize_ncu.py launches gpurun_out/launches.csv
    python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep [kernel-substring]
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpc__cycles_elapsed.avg.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_tma_ld.sum",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("%-72s %6s %11s %7s %10s" % ("kernel", "n", "total_ms", "share", "avg_ms"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %11.3f %6.1f%% %10.4f" % (k[:72], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    print("%-72s %6d %11.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        if pat and pat not in name:
            continue
        print("=== %s  (id %s)" % (name.split("(")[0], row[0]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-82s %18s %s" % (k, row[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
