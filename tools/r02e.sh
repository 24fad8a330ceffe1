#!/bin/bash
R=${1:-r02e}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
step "parity (gen4)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize.py -m gpu -x -q > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-40 $O/${R}_k_${n}_$name.json) $(tail -c 200 $O/${R}_k_${n}_$name.err)"
}
for n in 100 200 400 800; do
  run base $n X=1
  run hot0 $n TUNA_B200_HOT_MAX=0
  run hot48 $n TUNA_B200_HOT_MAX=49152
  run hot8 $n TUNA_B200_HOT_MAX=8192
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
step "ncu per-class counters (gen4, ET800)"
TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 400 ncu --metrics $M --clock-control none --csv --log-file $O/${R}_class_metrics.csv -k regex:k_shell4 -c 245 python tools/direct_timing.py child 800 > $O/${R}_ncu_c.log 2>&1; step "rc=$?"
