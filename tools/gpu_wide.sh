#!/bin/bash
# development: parity of the long-horizon path + device time per tick of the horizon variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_horizons.py -x -q > gpurun_out/wide_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/wide_pytest.log
for h in 34,14,24 51,14,36 68,14,48; do
  for s in 0 1; do
    timeout 300 python tools/horizon_speed.py 1024 $h $s >> gpurun_out/wide_speed.log 2>&1
  done
done
timeout 300 python tools/horizon_speed.py 1024 17,7,12 0 >> gpurun_out/wide_speed.log 2>&1
tail -5 gpurun_out/wide_pytest.log; cat gpurun_out/wide_speed.log
