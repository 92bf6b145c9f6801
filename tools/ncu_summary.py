"""Summarise ncu output for profiles/: (1) a `--page raw --csv` dump of an .ncu-rep ->
key metrics per captured launch; (2) a gpu__time_duration launch list -> time share per kernel.

    python tools/ncu_summary.py rep <file.ncu-rep> [...]
    python tools/ncu_summary.py launches <launches.csv>
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"## {path}")
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        print(f"kernel: {name}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                print(f"  {k:78s} {r[i]:>14s} {units[i]}")
        try:
            t_us = float(vals["gpu__time_duration.sum"][0])
            rd, wr = float(vals["dram__bytes_read.sum"][0]), float(vals["dram__bytes_write.sum"][0])
            unit = vals["dram__bytes_read.sum"][1]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
            tot = (rd + wr) * scale
            print(f"  -> DRAM traffic {tot / 1e6:.1f} MB per launch, {tot / (t_us * 1e-6) / 1e9:.0f} GB/s under ncu (cold cache, serialised)")
        except Exception:
            pass
    print()


def launches(path):
    agg = defaultdict(lambda: [0, 0.0])
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"## {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e6:.3f} ms of kernel time (ncu: cold cache, serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:70s} {n:8d} {ns / 1e6:10.3f} {ns / n / 1e3:9.2f} {100 * ns / total:6.1f}%")
    print()


if __name__ == "__main__":
    mode, files = sys.argv[1], sys.argv[2:]
    for f in files:
        (rep if mode == "rep" else launches)(f)
