#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_rXX.csv  > profiles/launches_rXX.md
    python profiles/summarize.py kernels  gpurun_out/prof_rXX.ncu-rep  > profiles/kernels_rXX.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("memory_l1_wavefronts_shared", "smem wavefronts"),
    ("memory_l1_wavefronts_shared_ideal", "smem wavefronts (ideal)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "mio_throttle", "lg_throttle", "math_pipe_throttle",
          "branch_resolving", "not_selected", "selected", "no_instructions", "drain", "dispatch_stall"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("velvet::", "").replace("<unnamed>::", "").replace("unnamed>::", "")


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1000.0 if unit == "ns" else v * 1000.0 if unit == "ms" else v
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


def kernels(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = short(r[ix["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        print(f"### `{name}`\n\n| metric | value |\n|---|---:|")
        for key, label in KEYS:
            if key in ix:
                print(f"| {label} (`{key}`) | {r[ix[key]]} {units[ix[key]]} |")
        tot = sum(float(r[ix[f'smsp__pcsamp_warps_issue_stalled_{s}']].replace(",", "")) for s in STALLS
                  if f"smsp__pcsamp_warps_issue_stalled_{s}" in ix)
        if tot:
            parts = []
            for s in STALLS:
                k = f"smsp__pcsamp_warps_issue_stalled_{s}"
                if k in ix:
                    v = float(r[ix[k]].replace(",", ""))
                    if v / tot >= 0.02:
                        parts.append((v / tot, s))
            print("| warp-stall breakdown (pc sampling) | " + ", ".join(f"{s} {100 * f:.0f}%" for f, s in sorted(parts, reverse=True)) + " |")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
