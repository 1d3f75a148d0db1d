#!/bin/bash
# One GPU checkpoint: tests, smoke, both bench arms, ncu launch list, one full ncu capture of K2 and K1.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --batch 16384 --steps 20 > $O/${TAG}_bench_b16384.json 2> $O/${TAG}_bench_b16384.err
CMD="python bench.py --steps 5 --warmup 3 --cpu-sample 8 --no-cpu --no-latency --rollout-ticks 0 --no-extras"
timeout 600 $CMD > $O/${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qp_condensed -s 4 -c 1 -o $O/${TAG}_qp -f $CMD > $O/${TAG}_ncu_qp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linearise -s 4 -c 1 -o $O/${TAG}_lin -f $CMD > $O/${TAG}_ncu_lin.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log; tail -2 $O/${TAG}_smoke.log; cut -c1-400 $O/${TAG}_bench.json
