#!/bin/bash
# ncu launch list + one full capture of the QP kernel on the final tree (after the same command exited 0 without ncu).
# usage (under gpurun): bash tools/gpu_ncu_final.sh <tag>
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 5 --warmup 3 --cpu-sample 8 --no-cpu --no-latency --rollout-ticks 0 --no-extras"
timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:qp_condensed -s 4 -c 1 -o $O/${TAG}_qp -f $CMD > $O/${TAG}_ncu_qp.log 2>&1
ls -la $O | tail -8
