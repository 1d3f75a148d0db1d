#!/usr/bin/env python
"""Development: per-source-line warp-stall samples and shared-memory wavefronts of one kernel from an ncu report captured with
--import-source on.   usage: python tools/ncu_source_lines.py <file.ncu-rep> [top_n]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, "", collections.Counter()])
tot = toti = totw = 0
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0] == "Function Name" or hdr is None or r[2] != "-":
        continue
    try:
        ln = int(r[0])
        s, ie = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        w = int(r[hdr.index("L1 Wavefronts Shared")])
    except ValueError:
        continue
    a = agg[(cur, ln)]
    a[0] += s; a[1] += ie; a[2] += w; a[3] = r[1].strip()
    for st in stalls:
        a[4][st] += int(r[hdr.index(st)])
    tot += s; toti += ie; totw += w
allst = collections.Counter()
for a in agg.values():
    allst.update(a[4])
print("samples", tot, "warp instructions", toti, "shared wavefronts", totw)
print("stall reasons:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(tot, 1)) for k, v in allst.most_common(12)))
print("\n-- by samples")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    t3 = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in a[4].most_common(3) if v)
    print("%-26s %4d  smp %5.2f%%  inst %5.2f%%  wf %5.2f%%  [%s]  %s" % (f, l, 100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(toti, 1),
                                                                      100.0 * a[2] / max(totw, 1), t3, a[3][:80]))
print("\n-- by shared-memory wavefronts")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][2])[:top // 2]:
    print("%-26s %4d  wf %5.2f%%  inst %5.2f%%  %s" % (f, l, 100.0 * a[2] / max(totw, 1), 100.0 * a[1] / max(toti, 1), a[3][:90]))
