"""Splits the SASS of one kernel of an ncu report at its BAR.SYNC instructions and prints, per phase, the executed warp
instructions, stall samples and opcode mix:   python profiles/phase_split.py <rep> <kernel regex> [tiles] [warps per CTA]"""
import collections, csv, io, re, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 4761
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
heads = [n for n, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[heads[0]]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))] if len(r) >= len(hdr)]
num = lambda r, k: int(r[ix[k]]) if r[ix[k]].isdigit() else 0
ti, ts = sum(num(r, "Instructions Executed") for r in body), sum(num(r, "# Samples") for r in body)
bars = [n for n, r in enumerate(body) if "BAR.SYNC" in r[1]]
for a, b in zip([0] + bars, bars + [len(body)]):
    i = sum(num(r, "Instructions Executed") for r in body[a:b])
    s = sum(num(r, "# Samples") for r in body[a:b])
    ops = collections.Counter()
    for r in body[a:b]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1].strip())
        ops[m.group(2).split(".")[0] if m else "?"] += num(r, "Instructions Executed")
    print(f"sass {a}-{b}: inst {100 * i / ti:.1f}% ({i / (tiles * warps):.0f} per warp and tile), samples {100 * s / max(ts, 1):.1f}%:",
          ", ".join(f"{k} {100 * v / max(i, 1):.0f}%" for k, v in ops.most_common(8)))
