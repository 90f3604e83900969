#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu source page, for a range of lines of one file.
usage: python tools/ncu_perline.py <src_page.csv> <disasm.txt> <kernel-substring> <file> <lo> <hi> [per=passes]"""
import collections, csv, re, sys
src_csv, disasm, sym, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
per = float(sys.argv[7]) if len(sys.argv) > 7 else 1e6
rows = list(csv.reader(open(src_csv)))
ix = {h: i for i, h in enumerate(rows[1])}
insts = rows[2:]
active, cur, k = False, None, 0
agg = collections.defaultdict(lambda: [0, 0, 0])
for ln in open(disasm).read().split("\n"):
    if ln.startswith(".text."):
        active = sym in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        r = insts[k]
        k += 1
        a = agg[cur]
        a[0] += int(r[ix["Instructions Executed"]])
        a[1] += int(r[ix["Thread Instructions Executed"]])
        a[2] += 1
import os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
text = open(os.path.join(root, "psim_b200", "csrc", fname)).read().split("\n")
tot = 0
for (f, l), (ie, te, n) in sorted(agg.items(), key=lambda kv: (str(kv[0][0]), kv[0][1])):
    if f == fname and lo <= l <= hi:
        tot += ie
        print(f"{l:4d} {ie / per:7.1f} thr {te / max(ie, 1):5.1f} sass {n:3d}  {text[l - 1].strip()[:100]}")
print("sum", tot / per)
