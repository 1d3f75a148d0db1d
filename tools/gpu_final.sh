#!/bin/bash
# End-of-round check on the final code: smoke, the bench line, GPU parity tests (incl. the reference-made golden files),
# the pinned-oracle tests against the travelling oracle/_ref library, the reference bench arm.
# Ordered shortest-first so a clamped GPU budget still returns the early logs.   usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 200 python -m pytest tests/test_reference_pinned.py -q > $O/${TAG}_pytest_reference_pinned.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_reference_pinned.log
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
tail -3 $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_reference_pinned.log; tail -2 $O/${TAG}_smoke.log; cut -c1-300 $O/${TAG}_bench.json
