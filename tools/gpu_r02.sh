#!/bin/bash
# Round-2 GPU checkpoint.  usage (under gpurun): bash tools/gpu_r02.sh <tag> [steps...]
#   steps: tests bench bench16k adjudicate ncu_k1 ncu_k2 ncu_list reference horizons   (default: tests bench)
TAG=${1:-r02a}; shift
STEPS=${@:-tests bench}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt
CMD="python bench.py --steps 5 --warmup 3 --cpu-sample 8 --no-cpu --no-latency --rollout-ticks 0 --no-extras"
for s in $STEPS; do
  case $s in
    smoke) timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log; tail -2 $O/${TAG}_smoke.log;;
    tests) timeout 1200 python -m pytest tests -q -m gpu --maxfail 12 -x --deselect tests/test_gpu_multi.py > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -40 $O/${TAG}_pytest_gpu.log;;
    tests_all) timeout 1500 python -m pytest tests -q -m gpu --maxfail 30 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -60 $O/${TAG}_pytest_gpu.log;;
    pinned) timeout 300 python -m pytest tests/test_reference_pinned.py -q > $O/${TAG}_pytest_reference_pinned.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_reference_pinned.log; tail -3 $O/${TAG}_pytest_reference_pinned.log;;
    bench) timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cut -c1-600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err;;
    benchq) timeout 300 python bench.py --no-cpu --no-latency --rollout-ticks 0 --no-extras > $O/${TAG}_benchq.json 2> $O/${TAG}_benchq.err; python tools/bench_brief.py $O/${TAG}_benchq.json; tail -3 $O/${TAG}_benchq.err;;
    bench16k) timeout 400 python bench.py --batch 16384 --steps 20 --no-cpu --no-latency --rollout-ticks 0 --no-extras > $O/${TAG}_bench_b16384.json 2> $O/${TAG}_bench_b16384.err; python tools/bench_brief.py $O/${TAG}_bench_b16384.json;;
    reference) timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; cut -c1-300 $O/${TAG}_bench_reference.json;;
    horizons) for h in 34,14,24 51,14,36 68,14,48; do timeout 300 python bench.py --horizon $h --steps 10 > $O/${TAG}_bench_h${h%%,*}.json 2> $O/${TAG}_bench_h${h%%,*}.err; python tools/bench_brief.py $O/${TAG}_bench_h${h%%,*}.json; done;;
    adjudicate) timeout 900 python tests/adjudicate_nonsolved.py 16384 199 > $O/${TAG}_adjudicate.log 2>&1; tail -30 $O/${TAG}_adjudicate.log;;
    ncu_list) timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1;;
    ncu_k1) timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:linearise -s 4 -c 1 -o $O/${TAG}_lin -f $CMD > $O/${TAG}_ncu_lin.log 2>&1;;
    ncu_k2) timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:qp_condensed -s 4 -c 1 -o $O/${TAG}_qp -f $CMD > $O/${TAG}_ncu_qp.log 2>&1;;
    fb_sanitize) timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fallback.py -x -q -k "every_instance and (0 or 3)" > $O/${TAG}_fb_sanitize.log 2>&1; echo "rc=$?" >> $O/${TAG}_fb_sanitize.log; tail -25 $O/${TAG}_fb_sanitize.log;;
    fb_tests) timeout 900 python -m pytest tests/test_gpu_fallback.py -q --maxfail 20 > $O/${TAG}_fb_tests.log 2>&1; echo "rc=$?" >> $O/${TAG}_fb_tests.log; tail -40 $O/${TAG}_fb_tests.log;;
    *) echo "unknown step $s";;
  esac
done
ls -la $O | tail -12
