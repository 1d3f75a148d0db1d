#!/usr/bin/env python
"""Development: joint-limit rows on the long horizons — ticks with the tight boxes of tests/test_gpu_joint_limits.py, the
long-horizon kernel's own working set (fallback off) against the hand-over to the fallback kernel (VSMPC_WIDE_NO_JL=1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
bat, syn, pack = bench.pkg("batched"), bench.pkg("synthetic"), bench.pkg("pack")
JMIN = [-30.0, 5.0, 20.0, 5.0, -30.0, 5.0, 20.0, 5.0]
JMAX = [-8.0, 30.0, 42.0, 28.0, -8.0, 30.0, 42.0, 28.0]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for hz in ((34, 14, 24),):
    nom = syn.make_states(B, perturbed=False)
    pers = [syn.make_states(B, seed=300 + j, perturbed=True, near_bound_fraction=0.3) for j in range(2)]
    nom_pack, packs = pack.build_pack(nom), [pack.build_pack(p) for p in pers]
    jp = np.ascontiguousarray(nom["joint_pos"][:, pack.DEFAULT_JOINT_SELECTOR].T)
    for name, lim, nsets in (("no joint-limit rows", None, 2), ("tight boxes, new states", dict(jointPos_min=JMIN, jointPos_max=JMAX), 2),
                             ("tight boxes, same state", dict(jointPos_min=JMIN, jointPos_max=JMAX), 1)):
        params = dict(nIter=hz[0], nIterSmall=hz[1], controlHorizon=hz[2])
        params.update(lim or {})
        mpc = bat.BatchedVSMPC(B, params, bench.load_traj())
        mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
        for j in range(3):
            mpc.update_pack(packs[j % nsets]); mpc.solveMPC()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 6
        for j in range(n):
            mpc.update_pack(packs[j % nsets]); mpc.solveMPC()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        nf, ns = mpc.get_counts()
        _, status = mpc.get_output()
        print(f"knots {hz[0]} B={B} {name:26s}: {dt*1e3:8.3f} ms per tick ({B/dt/1e6:.3f} M solves/s), solved {np.mean(status == 0):.4f}, "
              f"factorisations per solve: mean {nf.mean():.2f} max {nf.max()}")
        mpc.close()
