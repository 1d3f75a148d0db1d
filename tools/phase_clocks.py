#!/usr/bin/env python
"""Development: per-phase cycle counts of the condensed kernel (library built with VSMPC_PHASE_CLOCKS=1)."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bat, L = bench.pkg("batched"), bench.pkg("_lib")
nom_pack, jp, packs = bench.make_workload(B, 20251002, 2)
mpc = bat.BatchedVSMPC(B, None, bench.load_traj())
mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
for j in range(4):
    mpc.update_pack(packs[j % 2]); mpc.solveMPC()
n = min(B, 4096)
clk = np.zeros((n, 16), dtype=np.int64)
rc = L.load().vsmpc_debug_phase_clocks(clk.ctypes.data, n)
assert rc == 0, rc
t0 = clk[:, 0].min()
names = ["head (stage data)", "factorisation loop (A)", "AS wait + ftheta + forward (A)"]
d = np.stack([clk[:, 1] - clk[:, 0], clk[:, 2] - clk[:, 1], clk[:, 3] - clk[:, 2]], 1)
print("instances", n, "kernel span (cycles, same-SM clocks only comparable per SM):")
for i, nm in enumerate(names):
    print(f"  {nm:34s} mean {d[:, i].mean():9.0f}  p50 {np.median(d[:, i]):9.0f}  max {d[:, i].max():9.0f}")
b = np.stack([clk[:, 5] - clk[:, 4], clk[:, 6] - clk[:, 5]], 1)
print(f"  {'reduced QP + active set (B)':34s} mean {b[:, 0].mean():9.0f}  p50 {np.median(b[:, 0]):9.0f}  max {b[:, 0].max():9.0f}")
print(f"  {'F theta (both)':34s} mean {b[:, 1].mean():9.0f}")
if clk[:, 13].any():
    print(f"    of which: Omega down-date (both warps) mean {(clk[:, 13] - clk[:, 4]).mean():.0f}, reduced QP build + inverse mean {(clk[:, 14] - clk[:, 13]).mean():.0f}, "
          f"active set mean {(clk[:, 5] - clk[:, 14]).mean():.0f} max {(clk[:, 5] - clk[:, 14]).max():.0f}")
tot = clk[:, 3] - clk[:, 0]
print(f"  total per instance mean {tot.mean():.0f} p50 {np.median(tot):.0f} max {tot.max():.0f}")

it, nw = clk[:, 7] // 100, clk[:, 7] % 100
as_t = clk[:, 5] - clk[:, 4]
print("GI iterations: mean %.1f p50 %d p90 %d max %d ; final |W| mean %.1f max %d" % (it.mean(), np.median(it), np.percentile(it, 90), it.max(), nw.mean(), nw.max()))
for lo_, hi_ in ((0, 0), (1, 3), (4, 6), (7, 10), (11, 15), (16, 99)):
    m = (it >= lo_) & (it <= hi_)
    if m.any():
        print(f"  iters {lo_:2d}-{hi_:2d}: {m.sum():4d} instances, AS phase mean {as_t[m].mean():8.0f} cycles")

print("warp A per launch: a_prop %.0f  a_eliminate %.0f ; warp B: b_prop %.0f  b_downdate %.0f (cycles summed over knots, mean over instances)" %
      (clk[:, 8].mean(), clk[:, 9].mean(), clk[:, 10].mean(), clk[:, 11].mean()))

sm = clk[:, 12]
loop = clk[:, 2] - clk[:, 1]
per_sm = {}
for s_, l_, tt in zip(sm, loop, clk[:, 3] - clk[:, 0]):
    per_sm.setdefault(int(s_), []).append((l_, tt))
cnt = np.array([len(v) for v in per_sm.values()])
lm = np.array([np.mean([a for a, _ in v]) for v in per_sm.values()])
tm = np.array([np.max([b for _, b in v]) for v in per_sm.values()])
print("SMs used", len(per_sm), "instances per SM min/max", cnt.min(), cnt.max())
for c in sorted(set(cnt)):
    m = cnt == c
    print(f"  SMs with {c} CTAs: {m.sum():3d}; mean loop {lm[m].mean():8.0f} (min {lm[m].min():.0f}, max {lm[m].max():.0f}); slowest instance total: mean {tm[m].mean():.0f}, max {tm[m].max():.0f}")
