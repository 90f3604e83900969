#!/usr/bin/env python
"""(no GPU) dynamic view of one ncu capture taken with --import-source on: executed warp instructions by opcode, and a listing
of the SASS with per-instruction executed counts, average active threads and stall samples.
usage: python tools/ncu_hot.py <rep> [--listing out.txt] [--top N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
listing = sys.argv[sys.argv.index("--listing") + 1] if "--listing" in sys.argv else None
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
insts = rows[2:]
tot = sum(int(r[ix["Instructions Executed"]]) for r in insts)
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in insts)
by_op = collections.defaultdict(lambda: [0, 0, 0])
for r in insts:
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.split(".")[0]
    a = by_op[op]
    a[0] += int(r[ix["Instructions Executed"]])
    a[1] += int(r[ix["Thread Instructions Executed"]])
    a[2] += int(r[ix["# Samples"]] or 0)
print(f"kernel: {rows[0][1]}")
print(f"executed warp instructions {tot}, stall samples {tot_s}, static instructions {len(insts)}")
print(f"{'opcode':12s} {'warp-inst':>14s} {'share':>7s} {'avg thr':>8s} {'samples':>8s}")
for op, a in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:12s} {a[0]:14d} {100 * a[0] / tot:6.2f}% {a[1] / max(a[0], 1):8.1f} {100 * a[2] / max(tot_s, 1):7.2f}%")
if listing:
    with open(listing, "w") as f:
        for i, r in enumerate(insts):
            ie = int(r[ix["Instructions Executed"]])
            f.write(f"{i:5d} {100 * ie / tot:6.3f}% thr {r[ix['Avg. Threads Executed']]:>5s} smp {100 * int(r[ix['# Samples']] or 0) / max(tot_s, 1):6.3f}%  {r[ix['Source']].strip()}\n")
