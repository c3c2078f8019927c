#!/usr/bin/env python
"""Turns an `ncu --set full` report (gpurun_out/*.ncu-rep) into the metric table committed under profiles/, and an ncu launch
list (--metrics gpu__time_duration.sum --csv) into per-kernel shares.

    python tools/ncu_summary.py rep gpurun_out/r02_gram_full.ncu-rep            -> markdown table on stdout
    python tools/ncu_summary.py launches gpurun_out/r02_launches.csv            -> markdown table on stdout
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_membar",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_selected",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        get = dict(zip(hdr, zip(units, vals)))
        print(f"Kernel: `{get['Kernel Name'][1]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in get:
                print(f"| `{m}` | {get[m][1]} | {get[m][0]} |")
        print()


def launches(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"{n} launches, {total:.1f} ms of device time.\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.2f}% |")


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2])
