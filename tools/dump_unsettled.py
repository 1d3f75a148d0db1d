#!/usr/bin/env python
"""Development: run the parameter sweep with joint boxes (fallback off) and dump the QP data of the instances whose joint-box
working set did not settle, for analysis with tools/condensed_model.py on the CPU (tools/analyse_unsettled.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/unsettled.npz"
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
bat = bench.pkg("batched")
orig, orig_close = bat.BatchedVSMPC.__init__, bat.BatchedVSMPC.close
def init(self, *a, **k):
    orig(self, *a, **k)
    self.set_fallback(0)
def close(self):
    nf, ns = self.get_counts()
    _, status = self.get_output()
    bad = np.flatnonzero((status != 0) & (nf > 2))[:6]
    good = np.flatnonzero((status == 0) & (nf >= 3))[:2]
    idx = np.concatenate([bad, good])
    print("unsettled", np.flatnonzero((status != 0) & (nf > 2)).size, "dumping", idx.tolist(), "nf", nf[idx].tolist())
    A, BJ, BT, c, dt = self.get_dynamics()
    q, l, u = self.get_qp_vectors()
    H = self.getHessian(int(idx[0]) if idx.size else 0)
    np.savez_compressed(out, idx=idx, nf=nf[idx], status=status[idx], A=A[idx], BJ=BJ[idx], BT=BT[idx], c=c[idx], dt=dt, q=q[idx], l=l[idx], u=u[idx],
                        Hdiag=np.diag(H).copy(), H_tt=H[-24:, -24:].copy())
    orig_close(self)
bat.BatchedVSMPC.__init__, bat.BatchedVSMPC.close = init, close
r = bench.monte_carlo_leg(bat, n, ticks, 0, 1, 0, stream, dev, joint_boxes=True)
print(r["ms_per_tick"], r["solved_fraction_last_tick"])
