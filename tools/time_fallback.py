#!/usr/bin/env python
"""Development: time of the fallback kernel with every instance forced through it (vsmpc_set_fallback 2)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
bat = bench.pkg("batched")
for B in (int(a) for a in (sys.argv[1:] or ["64", "512"])):
    nom_pack, jp, packs = bench.make_workload(B, 20251002, 2)
    mpc = bat.BatchedVSMPC(B, None, bench.load_traj())
    mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
    for mode in (1, 2):
        mpc.set_fallback(mode)
        for j in range(2):
            mpc.update_pack(packs[j % 2]); mpc.solveMPC()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for j in range(4):
            mpc.update_pack(packs[j % 2]); mpc.solveMPC()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
        print(f"B={B} fallback mode {mode}: {dt*1e3:.3f} ms per tick")
    mpc.close()
