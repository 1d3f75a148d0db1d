#!/usr/bin/env python
"""Per-region stall-reason breakdown of an ncu SASS source-page CSV (see ncu_lines.py for the inputs).
usage: ncu_stalls.py src.csv cubin.asm kernel "name:lo-hi,..." """
import csv
import re
import sys
from collections import defaultdict

src_csv, asm, kname, regions = sys.argv[1:5]
lines = []
cur = None
infn = False
for ln in open(asm, errors="replace"):
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        infn = kname in ln
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ii = hdr.index("Instructions Executed")
stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) > ii]
regs = []
for spec in regions.split(","):
    name, rng = spec.split(":")
    lo, hi2 = [int(x) for x in rng.split("-")]
    regs.append((name, lo, hi2))
agg = defaultdict(lambda: defaultdict(int))
inst = defaultdict(int)
for k, r in enumerate(data):
    key = lines[k] if k < len(lines) else None
    name = "other"
    if key and key[0].startswith("vsmpc_qp"):
        for n, lo, hi2 in regs:
            if lo <= key[1] <= hi2:
                name = n
                break
    inst[name] += int(r[ii])
    for h, c in stall_cols:
        try:
            agg[name][h] += int(r[c])
        except ValueError:
            pass
tot = sum(sum(v.values()) for v in agg.values())
print(f"{'region':12s} {'inst%':>6s} {'samp%':>6s} {'cyc/inst':>8s}  top stalls")
ti = sum(inst.values())
for name in [n for n, _, _ in regs] + ["other"]:
    s = sum(agg[name].values())
    top = sorted(agg[name].items(), key=lambda kv: -kv[1])[:5]
    tops = ", ".join(f"{h[6:]} {100*v/max(s,1):.0f}%" for h, v in top)
    print(f"{name:12s} {100*inst[name]/ti:6.1f} {100*s/max(tot,1):6.1f} {s/max(inst[name],1)*1e3:8.2f}  {tops}")
