#!/bin/bash
# the default bench line (all legs) as the driver runs it, with its wall time.  usage (under gpurun): bash tools/gpu_bench_full.sh <tag>
TAG=${1:-r02i}; O=gpurun_out; mkdir -p $O
T0=$(date +%s)
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "rc=$? wall=$(python -c "import time,sys;print(round(time.time()-float(sys.argv[1]),1))" $T0) s" | tee $O/${TAG}_bench_wall.txt
tail -5 $O/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.3f M/s e2e %.3f M/s blocking %.3f M/s k2 %.1f us k1 %.1f us frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["blocking_value"]/1e6, r["kernel_ms_per_launch"]*1e3, r["linearise_ms_per_launch"]*1e3, r["frac"]))
for k in ("e2e_kinematics","gather","long_horizon","monte_carlo","closed_loop","single_solve_latency","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
PY
