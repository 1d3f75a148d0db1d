#!/usr/bin/env python
"""Development: device time of one tick for a horizon variant (BASELINE configs[3])."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
B = int(sys.argv[1]); N, Ns, Nc = [int(x) for x in sys.argv[2].split(",")]; solver = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bat = bench.pkg("batched")
nom_pack, jp, packs = bench.make_workload(B, 20251002, 2)
mpc = bat.BatchedVSMPC(B, dict(nIter=N, nIterSmall=Ns, controlHorizon=Nc), bench.load_traj(), solver=solver)
mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
d = [torch.from_numpy(p).cuda() for p in packs]
for j in range(3):
    mpc.update_device_ptr(d[j % 2].data_ptr()); mpc.solve_async()
mpc.wait()
t0 = time.perf_counter(); K = 10
for j in range(K):
    mpc.update_device_ptr(d[j % 2].data_ptr()); mpc.solve_async()
mpc.wait()
dt = (time.perf_counter() - t0) / K
_, st = mpc.get_output()
print(f"horizon ({N},{Ns},{Nc}) solver {solver} B={B}: {dt*1e3:.2f} ms/tick, {B/dt/1e3:.1f} k solves/s, solved {float((st==0).mean()):.3f}, n_var {mpc.n_var}")
