#!/bin/bash
# Per-class-job counters of one ET800 direct build (development aid): job table in launch order + a few ncu metrics for every launch
# of the first build (CSV, small), and one full capture with source of a typical mid-weight class job.
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__inst_executed.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,l1tex__t_sector_hit_rate.pct
TUNA_B200_DUMP_JOBS=gpurun_out/jobs800.csv python tools/direct_timing.py child 800 > gpurun_out/prof_classes_plain.log 2>&1
timeout 420 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01c_class_metrics.csv -k regex:k_shell_jk -c 260 python tools/direct_timing.py child 800 > gpurun_out/ncu_classes.log 2>&1
tail -n 2 gpurun_out/ncu_classes.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_shell_jk_one -s 20 -c 1 -f -o gpurun_out/r01c_shell_src python tools/direct_timing.py child 800 > gpurun_out/ncu_src.log 2>&1
tail -n 2 gpurun_out/ncu_src.log
du -sh gpurun_out
