#!/usr/bin/env python
"""Development: the parameter sweep with per-instance joint boxes (bench.monte_carlo_leg, joint_boxes=True) with the fallback
kernel on / off, and who asks for the fallback: factorisation counts and statuses of the last tick."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
bat = bench.pkg("batched")
orig, orig_close = bat.BatchedVSMPC.__init__, bat.BatchedVSMPC.close
for mode in (1, 0):
    def init(self, *a, _m=mode, **k):
        orig(self, *a, **k)
        self.set_fallback(_m)
    def close(self):
        nf, ns = self.get_counts()
        _, status = self.get_output()
        print("   last tick: status counts", dict(zip(*np.unique(status, return_counts=True))), "factorisations histogram 0..9",
              np.bincount(nf, minlength=10)[:10].tolist(), "| of the non-solved:", np.bincount(nf[status != 0], minlength=10)[:10].tolist())
        orig_close(self)
    bat.BatchedVSMPC.__init__, bat.BatchedVSMPC.close = init, close
    r = bench.monte_carlo_leg(bat, n, ticks, 0, 1, 0, stream, dev, joint_boxes=True)
    print(f"instances {n} fallback mode {mode}: {r['ms_per_tick']:.3f} ms/tick, {r['value']/1e6:.3f} M closed-loop solves/s, solved {r['solved_fraction_last_tick']:.5f}, "
          f"factorisations per solve {r['factorisations_per_solve_last_tick']:.2f}")
bat.BatchedVSMPC.__init__, bat.BatchedVSMPC.close = orig, orig_close
