set -x
CMD="python bench.py --steps 3 --warmup 3 --solver 0 --cpu-sample 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_structured -s 3 -c 1 -o gpurun_out/prof_qp -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
