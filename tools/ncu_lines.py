#!/usr/bin/env python
"""Aggregate an ncu SASS source-page CSV by CUDA source line using nvdisasm line info.

usage: ncu_lines.py <src.csv from `ncu -i rep --page source --csv`> <cubin.asm from `nvdisasm -g -c`> <kernel substring>
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, asm, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# 1) instruction order with line info from nvdisasm
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
ii, sa, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = [r for r in rows[hi + 1:] if len(r) > ii]
print("sass rows", len(data), "asm instr", len(lines))
agg = defaultdict(lambda: [0, 0])
tot = tots = 0
for k, r in enumerate(data):
    n, s = int(r[ii]), int(r[sa])
    key = lines[k] if k < len(lines) else None
    agg[key][0] += n
    agg[key][1] += s
    tot += n
    tots += s
print("total warp-inst", tot, "samples", tots)
srcs = {}
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    text = ""
    if key:
        f, l = key
        if f not in srcs:
            import glob
            c = glob.glob(f"/root/repo/**/{f}", recursive=True)
            srcs[f] = open(c[0]).read().split("\n") if c else []
        if srcs[f] and l - 1 < len(srcs[f]):
            text = srcs[f][l - 1].strip()[:90]
    print(f"{100*s/max(tots,1):5.1f}% samp {100*n/tot:5.1f}% inst {n:9d}  {key}  {text}")

# ---- optional: shares by line ranges "name:lo-hi,..." in argv[5] --------------------------------------
if len(sys.argv) > 5:
    print("\nregion shares")
    for spec in sys.argv[5].split(","):
        name, rng = spec.split(":")
        lo, hi = [int(x) for x in rng.split("-")]
        n = sum(v[0] for k, v in agg.items() if k and k[0].startswith("vsmpc_qp") and lo <= k[1] <= hi)
        s = sum(v[1] for k, v in agg.items() if k and k[0].startswith("vsmpc_qp") and lo <= k[1] <= hi)
        print(f"  {name:14s} inst {100*n/tot:5.1f}%  samples {100*s/max(tots,1):5.1f}%")
    n = sum(v[0] for k, v in agg.items() if not (k and k[0].startswith("vsmpc_qp")))
    s = sum(v[1] for k, v in agg.items() if not (k and k[0].startswith("vsmpc_qp")))
    print(f"  {'other files':14s} inst {100*n/tot:5.1f}%  samples {100*s/max(tots,1):5.1f}%")
