"""Per-source-line totals of an ncu report (needs -lineinfo + --import-source on):
    python profiles/src_lines.py gpurun_out/x.ncu-rep [min_pct]
Prints, per file:line, warp instructions executed, stall samples, and the dominant stall reasons."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--kernel-name", sys.argv[3]] if len(sys.argv) > 3 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
acc = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # 'Source' appears twice (cuda line text, sass); the dict keeps the last; first is index 1
        continue
    if hdr is None or len(r) < 10 or not r[0].isdigit():
        continue
    if r[hdr["Address"]] != "-":
        continue  # SASS rows nested under the line; the line row carries the totals
    key = (cur_file, int(r[0]))
    def g(name):
        try:
            return int(r[hdr[name]])
        except (ValueError, KeyError):
            return 0
    a = acc.setdefault(key, {"text": r[1].strip()[:90], "inst": 0, "samples": 0, "st": collections.Counter()})
    a["inst"] += g("Instructions Executed")
    a["samples"] += g("# Samples")
    for name in hdr:
        if name.startswith("stall_") and "(Not Issued)" not in name:
            a["st"][name[6:]] += g(name)
ti = sum(a["inst"] for a in acc.values()) or 1
ts = sum(a["samples"] for a in acc.values()) or 1
print(f"total warp instructions {ti}, stall samples {ts}")
print("| file:line | inst % | samples % | top stalls | source |\n|---|---:|---:|---|---|")
for (f, ln), a in acc.items():
    pi, ps = 100 * a["inst"] / ti, 100 * a["samples"] / ts
    if pi < minpct and ps < minpct:
        continue
    top = ", ".join(f"{k} {v}" for k, v in a["st"].most_common(3) if v)
    print(f"| {f}:{ln} | {pi:.1f} | {ps:.1f} | {top} | `{a['text']}` |")
