#!/bin/bash
# generation-4 engine, second GPU contact: parity, knob sweep (term mode threshold, group size, batch size), per-class counters
R=${1:-r02d}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
step "parity (gen4)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize.py -m gpu -x -q > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
run() { # name nbf env...
  local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-40 $O/${R}_k_${n}_$name.json)"
}
for n in 100 200 400 800; do
  run gen2 $n TUNA_B200_ENGINE=2
  run base $n X=1
  run term0 $n TUNA_B200_TERM_MAX=0
  run term6k $n TUNA_B200_TERM_MAX=6000
done
for n in 400 800; do
  run spl512 $n TUNA_B200_SMEM_PER_LANE=512
  run spl128 $n TUNA_B200_SMEM_PER_LANE=128
  run gdiv16 $n TUNA_B200_G_DIV=16
  run gdiv4 $n TUNA_B200_G_DIV=4
  run nb1 $n TUNA_B200_NB=1
  run nb4 $n TUNA_B200_NB=4
  run nbbytes140 $n TUNA_B200_NB_BYTES=143360
  run tier0 $n TUNA_B200_TIER2=0
  run itb4096 $n TUNA_B200_IT_BUDGET=4096
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
step "ncu per-class counters (gen4, ET800)"
TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 400 ncu --metrics $M --clock-control none --csv --log-file $O/${R}_class_metrics.csv -k regex:k_shell4 -c 245 python tools/direct_timing.py child 800 > $O/${R}_ncu_c.log 2>&1; step "rc=$?"
du -sh $O | tee -a $O/${R}_steps.log
