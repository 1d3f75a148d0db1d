#!/usr/bin/env python
"""Development: cost of the fallback kernel in the Monte Carlo sweep (bench.monte_carlo_leg with the fallback on / off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
bat = bench.pkg("batched")
orig = bat.BatchedVSMPC.__init__
for mode in (1, 0):
    def init(self, *a, _m=mode, **k):
        orig(self, *a, **k)
        self.set_fallback(_m)
    bat.BatchedVSMPC.__init__ = init
    r = bench.monte_carlo_leg(bat, n, 200, 0, 1, 0, stream, dev)
    print(f"instances {n} fallback mode {mode}: {r['ms_per_tick']:.3f} ms/tick, {r['value']/1e6:.2f} M closed-loop solves/s, solved {r['solved_fraction_last_tick']:.5f}")
bat.BatchedVSMPC.__init__ = orig
