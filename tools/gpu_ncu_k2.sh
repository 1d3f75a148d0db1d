#!/bin/bash
# full ncu capture of the K2 kernel named by $1 (regex) for bench args $2.. ; output gpurun_out/$TAG.ncu-rep
KREGEX=${1:-qp_condensed}; TAG=${2:-k2}; shift 2
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-latency --rollout-ticks 0 --no-extras $@"
timeout 600 $CMD > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 4 -c 1 -o $O/$TAG -f $CMD > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log | cut -c1-300
