#!/usr/bin/env python
"""Development: per-phase cycle counts of the long-horizon condensed kernel (library built with VSMPC_PHASE_CLOCKS=1).
usage: phase_clocks_wide.py B N,Ns,Nc"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
B = int(sys.argv[1]); N, Ns, Nc = [int(x) for x in sys.argv[2].split(",")]
bat, L = bench.pkg("batched"), bench.pkg("_lib")
nom_pack, jp, packs = bench.make_workload(B, 20251002, 2)
mpc = bat.BatchedVSMPC(B, dict(nIter=N, nIterSmall=Ns, controlHorizon=Nc), bench.load_traj())
mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
for j in range(4):
    mpc.update_pack(packs[j % 2]); mpc.solveMPC()
n = min(B, 4096)
clk = np.zeros((n, 16), dtype=np.int64)
rc = L.load().vsmpc_debug_phase_clocks(clk.ctypes.data, n)
assert rc == 0, rc
names = ["stage data", "recursion", "Omega down-date (DMMA)", "reduced QP + inverse", "active set", "F theta", "forward"]
print(f"horizon ({N},{Ns},{Nc}) B={B}: cycles per instance")
for i, nm in enumerate(names):
    d = clk[:, i + 1] - clk[:, i]
    print(f"  {nm:26s} mean {d.mean():10.0f}  p50 {np.median(d):10.0f}  max {d.max():10.0f}")
tot = clk[:, 7] - clk[:, 0]
it = clk[:, 8]
print(f"  total mean {tot.mean():.0f} max {tot.max():.0f}; active-set iterations mean {it.mean():.1f} p90 {np.percentile(it, 90):.0f} max {it.max()}")
print(f"  recursion work summed over knots: warp 0 a_prop {clk[:, 9].mean():.0f} + a_eliminate {clk[:, 11].mean():.0f}; first column warp "
      f"w_prop {clk[:, 13].mean():.0f}; last column warp w_prop {clk[:, 10].mean():.0f} + w_downdate {clk[:, 12].mean():.0f}")
print(f"  final working set: mean {clk[:, 14].mean():.1f} p10 {np.percentile(clk[:, 14], 10):.0f} p90 {np.percentile(clk[:, 14], 90):.0f} max {clk[:, 14].max()}")
print(f"  cycles per active-set iteration: {((clk[:, 5] - clk[:, 4]).sum() / max(it.sum(), 1)):.0f}")
