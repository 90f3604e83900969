#!/usr/bin/env python
"""Join an ncu SASS source page (per-instruction executed counts) with nvdisasm line info and aggregate by source line.
usage: python tools/ncu_lines.py <src_page.csv> <disasm.txt> <kernel-symbol-substring> [top]"""
import csv, re, sys, collections
src_csv, disasm, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
insts = rows[2:]
# disasm: sequence of instructions for the kernel with current (file, line); inlined-at chains appear as extra comments
lines = open(disasm).read().split("\n")
cur = None
seq = []
active = False
for ln in lines:
    if ln.startswith(".text."):
        active = sym in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        seq.append(cur)
print("sass insts in report", len(insts), "in disasm", len(seq))
n = min(len(insts), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_t = 0
for k in range(n):
    r = insts[k]
    ie = int(r[ix["Instructions Executed"]]); te = int(r[ix["Thread Instructions Executed"]])
    sm = int(r[ix["# Samples"]]) if r[ix["# Samples"]] else 0
    a = agg[seq[k]]
    a[0] += ie; a[1] += te; a[2] += sm
    tot_i += ie; tot_t += te
print("total warp insts", tot_i, "thread insts", tot_t, "avg threads", tot_t / max(tot_i, 1))
tot_s = sum(a[2] for a in agg.values())
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):45s} warp-inst {a[0]:10d} ({100*a[0]/tot_i:5.1f}%)  avg-thr {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/max(tot_s,1):5.1f}%")

# grouped view: consecutive source-line ranges of device_core.cuh / kernels.cuh
if len(sys.argv) > 5:
    groups = collections.defaultdict(lambda: [0, 0, 0])
    bounds = [tuple(x.split(":")) for x in sys.argv[5].split(",")]  # name:file:lo:hi
    for key, a in agg.items():
        if key is None: continue
        g = "other"
        for name, fn, lo, hi in bounds:
            if key[0] == fn and int(lo) <= key[1] <= int(hi):
                g = name; break
        else:
            g = key[0] if key[0] not in ("device_core.cuh", "kernels.cuh") else "other:" + key[0]
        for i in range(3): groups[g][i] += a[i]
    print("---- grouped")
    for g, a in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print(f"{g:28s} warp-inst {a[0]:10d} ({100*a[0]/tot_i:5.1f}%)  avg-thr {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/max(tot_s,1):5.1f}%")
