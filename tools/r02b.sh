#!/bin/bash
# generation-4 engine, first GPU contact: direct-mode parity, A/B timing against generation 2, per-class counters.
R=${1:-r02b}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
step "parity direct (gen4)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_scf_closed_loop.py -m gpu -x -q > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
for n in 100 200 400 800; do
  for e in 2 4; do
    TUNA_B200_ENGINE=$e timeout 200 python tools/direct_timing.py child $n > $O/${R}_ab_${n}_e$e.json 2> $O/${R}_ab_${n}_e$e.err; step "nbf $n engine $e: $(cat $O/${R}_ab_${n}_e$e.json | cut -c1-120)"
  done
done
step "fullsize parity"
timeout 400 python -m pytest tests/test_zz_fullsize.py -m gpu -q > $O/${R}_zz.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_zz.log)"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
step "ncu per-class counters (gen4, ET800)"
TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 400 ncu --metrics $M --clock-control none --csv --log-file $O/${R}_class_metrics.csv -k regex:k_shell4 -c 245 python tools/direct_timing.py child 800 > $O/${R}_ncu_c.log 2>&1; step "rc=$?"
du -sh $O | tee -a $O/${R}_steps.log
