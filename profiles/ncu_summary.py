#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/ into text for profiles/.
    python profiles/ncu_summary.py launches gpurun_out/launches.csv
    python profiles/ncu_summary.py full gpurun_out/prof.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum"]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:9.2f} {v[1] / tot:6.3f}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in FULL if m in hdr]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""))
        for m, i in cols:
            print(f"    {m:66s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
