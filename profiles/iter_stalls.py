"""Summarise an ncu report of iterate_tile_kernel: time, instructions, issue utilisation, stall mix, top stall sites.
Usage: python profiles/iter_stalls.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]:
    print(f"{k}: {d.get(k)}")
st = {k.split("stalled_")[1].split("_per")[0]: float(d[k]) for k in rows[0]
      if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k}
print("stall cycles per issue:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v >= 0.05))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, body = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
for r in body:
    try:
        n = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1].strip())
    ops[m.group(2).split(".")[0] if m else "?"] += n
tot = sum(ops.values())
print("opcode mix:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(14)))
for col in ("stall_long_sb", "stall_barrier", "stall_short_sb"):
    lst = sorted(((int(r[ix[col]]), n, r[1].strip()[:60]) for n, r in enumerate(body) if r[ix[col]].isdigit() and int(r[ix[col]]) > 0),
                 reverse=True)
    print(col, "samples", sum(v for v, _, _ in lst), "top:", "; ".join(f"{v}@{n} {s}" for v, n, s in lst[:6]))
