#!/bin/bash
# long-horizon checkpoint: parity, bench lines per horizon variant, ncu launch list + full capture of the wide kernel
tag=${1:-r01f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_horizons.py -x -q > gpurun_out/${tag}_pytest_horizons.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_horizons.log
for h in 34,14,24 51,14,36 68,14,48; do
  n=${h%%,*}
  timeout 600 python bench.py --steps 10 --warmup 3 --horizon $h > gpurun_out/${tag}_bench_h${n}.json 2> gpurun_out/${tag}_bench_h${n}.err
  timeout 600 python bench.py --steps 5 --warmup 3 --horizon $h --solver 1 > gpurun_out/${tag}_bench_h${n}_generic.json 2>> gpurun_out/${tag}_bench_h${n}.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --horizon 34,14,24 > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_wide_launches.csv \
   python bench.py --steps 5 --warmup 3 --horizon 34,14,24 > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qp_condensed_wide -s 4 -c 1 -o gpurun_out/${tag}_wide \
   python bench.py --steps 5 --warmup 3 --horizon 34,14,24 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_pytest_horizons.log
for f in gpurun_out/${tag}_bench_h*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.4g e2e %.4g k2 %.3f ms frac %.3f solved %.3f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['solved_fraction']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
