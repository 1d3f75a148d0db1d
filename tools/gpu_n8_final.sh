#!/bin/bash
# N = 8 bench line of the final tree, launched the way the driver launches it.  usage (under gpurun --gpus 8): bash tools/gpu_n8_final.sh <tag>
TAG=${1:-r02n8}; O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
echo "rc=$?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_n8.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.2f M e2e %.2f M" % (d["value"]/1e6, d["e2e"]["value"]/1e6), "clocks", d["clocks"]["samples"], d["clocks"]["sm_mhz"])
for k in ("e2e_kinematics","monte_carlo","param_sweep","e2e_cpp_host"):
    v=d.get(k) or {}
    print(k, "%.2f M" % (v.get("value",0)/1e6))
print(d["gather"])
PY
