#!/bin/bash
# Round-1 closing GPU call (one box, ~8 minutes): parity of the new AO->MO path, the default bench line, and the ncu evidence for
# profiles/ — most important first, every step under its own timeout, everything written to gpurun_out/ as it completes.
R=${1:-r01f}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }

step "pytest mo_transform"
timeout 150 python -m pytest tests/test_mo_transform.py -m gpu -x -q > $O/${R}_pytest_mo.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_mo.log)"

step "bench default"
timeout 330 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; step "rc=$? $(wc -c < $O/${R}_bench_n1.json) bytes"

BENCH="python bench.py --steps 1 --warmup 3 --no-stored"
step "ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench.csv $BENCH > $O/${R}_ncu_l.log 2>&1; step "rc=$?"

IDX=$(python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/${R}_launches_bench.csv")) if len(r) > 5 and r[0].isdigit() and "k_shell_jk_one" in r[4]]
n = len(rows)
per = n // 6                       # 3 warm-up + 1 timed + 1 e2e warm-up + 1 e2e build
last = rows[3 * per:4 * per]
best = max(range(len(last)), key=lambda i: float(last[i][-1].replace(",", ""))) if last else 0
print(3 * per + best)
PY
)
step "ncu full capture of k_shell_jk_one launch $IDX"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shell_jk_one -s ${IDX:-0} -c 1 -f -o $O/${R}_prof_shell $BENCH > $O/${R}_ncu_s.log 2>&1; step "rc=$?"

M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
step "ncu FP64 op counters, all class-job launches of one ET800 build"
TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 240 ncu --metrics $M --clock-control none --csv --log-file $O/${R}_class_metrics.csv -k regex:k_shell_jk -c 245 python tools/direct_timing.py child 800 > $O/${R}_ncu_c.log 2>&1; step "rc=$?"

step "ncu full capture: stored J/K (Ne2 UHF/cc-pVQZ) and AO->MO GEMM (N2/cc-pVTZ)"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_jk_stored_sym|k_axis_gemm" -c 6 -f -o $O/${R}_prof_stored_mo python tools/stored_check.py profile > $O/${R}_ncu_t.log 2>&1; step "rc=$?"

step "pytest -m gpu (rest of the suite)"
timeout 420 python -m pytest tests -m gpu -x -q --deselect tests/test_mo_transform.py > $O/${R}_pytest_gpu.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_gpu.log)"
du -sh $O | tee -a $O/${R}_steps.log
