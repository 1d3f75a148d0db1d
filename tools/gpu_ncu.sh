set -x
CMD="python bench.py --steps 3 --warmup 3 --solver 0 --cpu-sample 4"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_structured -s 3 -c 1 -o gpurun_out/prof_qp -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
