#!/bin/bash
# quick GPU check: parity tests + bench for a few solver / batch combinations
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/quick_pytest.log
cat $O/quick_pytest.log
for args in "--solver 0" "--solver 2" "--solver 0 --batch 16384 --steps 20" "--solver 0 --batch 4096 --steps 20"; do
  timeout 600 python bench.py $args --no-cpu --no-latency --rollout-ticks 0 --no-extras > $O/quick_bench.json 2> $O/quick_bench.err || tail -5 $O/quick_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/quick_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$args", "value=%.3gM/s e2e=%.3gM/s k2=%.1fus k1=%.1fus frac=%.3f solved=%.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r["kernel_ms_per_launch"]*1e3, r["linearise_ms_per_launch"]*1e3, r["frac"], d["solved_fraction"]))
except Exception as e:
    print("$args", "failed", e)
PY
done
