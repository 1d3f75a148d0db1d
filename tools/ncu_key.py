#!/usr/bin/env python
"""Key counters of one kernel launch from an .ncu-rep (ncu -i rep --page raw --csv).  usage: ncu_key.py rep [more]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','sm__sass_inst_executed_op_shared_ld.sum']
for k in keys:
    if k in d: print(f"{k:75s} {u[k]:12s} {d[k]}")
for k in hdr:
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k and float(d[k] or 0) > 0.05:
        print(f"{k:75s} {d[k]}")
