#!/bin/bash
# One GPU call that regenerates the round's evidence: full -m gpu suite, smoke, the default bench line, the ncu launch list of the bench
# command, per-class-job ncu counters of one ET800 build (-> profiles/executed_fp64.json via tools/executed_from_ncu.py) and one
# `--set full` capture with source of the heaviest class job.  Usage: bash tools/gpu_round.sh <tag>
R=${1:-r02}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
step "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${R}_pytest_gpu.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_gpu.log)"
step "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_smoke.log)"
step "bench default"
timeout 600 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; step "rc=$? $(wc -c < $O/${R}_bench_n1.json) bytes"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum
step "ncu per-class counters (one ET800 build)"
TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 500 ncu --metrics $M --clock-control none --csv --log-file $O/${R}_class_metrics.csv -k regex:k_shell4 -c 231 python tools/direct_timing.py child 800 > $O/${R}_ncu_c.log 2>&1; step "rc=$?"
step "ncu DRAM traffic of the stored J/K and AO->MO kernels (N2/cc-pVTZ)"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${R}_stored_mo_traffic.csv -k regex:"k_jk_stored|k_sym_reduce|k_axis_gemm" python tools/stored_check.py profile n2_ccpvtz > $O/${R}_ncu_t.log 2>&1; step "rc=$?"
step "dense fill: engine vs per-AO-quartet kernel; ncu time and DRAM bytes of the scatter pass"
timeout 200 python tools/fill_timing.py > $O/${R}_fill_timing.log 2>&1; step "rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_fill_scatter -c 12 --csv --log-file $O/${R}_fill_scatter_ncu.csv python tools/fill_timing.py n2_ccpvtz ne2_uhf_ccpvqz > $O/${R}_ncu_f.log 2>&1; step "rc=$?"
step "ncu launch list of the bench command"
timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/${R}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-stored > $O/${R}_ncu_l.log 2>&1; step "rc=$?"
step "ncu full capture of the heaviest class job"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shell4_one -s 0 -c 1 -f -o $O/${R}_prof_shell python tools/direct_timing.py child 800 > $O/${R}_ncu_s.log 2>&1; step "rc=$?"
ncu -i $O/${R}_prof_shell.ncu-rep --page raw --csv > $O/${R}_prof_shell_raw.csv 2>/dev/null
ncu -i $O/${R}_prof_shell.ncu-rep --page source --csv > $O/${R}_prof_shell_sass.csv 2>/dev/null
du -sh $O | tee -a $O/${R}_steps.log
