#!/usr/bin/env python
"""Development: where the end-to-end step time goes (host buffers through the C-ABI), B = 1024."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
B = 1024
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SMI = len(sys.argv) > 3 and sys.argv[3] == 'smi'
bat, L = bench.pkg("batched"), bench.pkg("_lib")
nom_pack, jp, packs = bench.make_workload(B, 20251002, NP)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
mpc = bat.BatchedVSMPC(B, None, bench.load_traj()); mpc.set_stream(stream.cuda_stream)
mpc.configure_pack(nom_pack, jp, (np.arange(B) % 20).astype(np.int32))
hp = [torch.from_numpy(p).pin_memory() for p in packs]
dp = [torch.from_numpy(p).to(dev) for p in packs]
ho = [torch.empty((B, 54), dtype=torch.float64).pin_memory() for _ in range(2)]
hs = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]

smi = None
if SMI:
    import subprocess
    smi = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL)
print(f"K={K} host pack sets={NP} nvidia-smi sampler={'on' if SMI else 'off'}")
def run(name, body):
    for j in range(10): body(j)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for j in range(K): body(j)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"{name:58s} {dt*1e6:7.1f} us/step  {B/dt/1e6:.3f} M solves/s")

def dev_only(j): mpc.update_device_ptr(dp[j % NP].data_ptr()); mpc.solve_async()
def h2d_only(j): mpc.update_ptr(hp[j % NP].data_ptr()); mpc.solve_async()
state = {"prev": None}
def full(j):
    mpc.update_ptr(hp[j % NP].data_ptr()); mpc.solve_async()
    t = mpc.get_output_async(ho[j & 1].data_ptr(), hs[j & 1].data_ptr())
    if state["prev"] is not None: mpc.wait_output(state["prev"])
    state["prev"] = t
def dev_out(j):
    mpc.update_device_ptr(dp[j % NP].data_ptr()); mpc.solve_async()
    t = mpc.get_output_async(ho[j & 1].data_ptr(), hs[j & 1].data_ptr())
    if state["prev"] is not None: mpc.wait_output(state["prev"])
    state["prev"] = t
def host_only(j):
    t0 = time.perf_counter()
run("device packs, no read-back (K1 + K2 + fallback launch)", dev_only)
run("host packs (H2D staged), no read-back", h2d_only)
run("device packs + asynchronous read-back", dev_out)
state["prev"] = None
run("host packs + asynchronous read-back (the e2e leg)", full)
# host time of the calls alone: enqueue without waiting
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(50):
    mpc.update_ptr(hp[j % NP].data_ptr()); mpc.solve_async(); mpc.get_output_async(ho[j & 1].data_ptr(), hs[j & 1].data_ptr())
t_enq = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
print(f"host time to enqueue one step (3 C-ABI calls): {t_enq*1e6:.1f} us")

if smi: smi.terminate()
