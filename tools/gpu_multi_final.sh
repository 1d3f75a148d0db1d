#!/bin/bash
# N = 2 check on the final tree: multi-GPU parity tests and the bench line launched the way the driver launches it.
# usage (under gpurun --gpus 2): bash tools/gpu_multi_final.sh <tag>
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_multi.py -q > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
tail -3 $O/${TAG}_pytest_multi.log; cut -c1-300 $O/${TAG}_bench_n2.json
