#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command + full captures of the two dominant kernels.
set -e
R=${1:-r01}
BENCH="python bench.py --steps 2 --warmup 3"
$BENCH > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv $BENCH > gpurun_out/ncu_l.log 2>&1 || true
# heaviest shell-engine launch of the last timed build
IDX=$(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${R}_launches_bench.csv")) if len(r) > 5 and r[0].isdigit() and "k_shell_jk" in r[4]]
n = len(rows)
per = n // 5                       # 3 warm-up + 2 timed builds
last = rows[-per:]
best = max(range(len(last)), key=lambda i: float(last[i][-1].replace(",", "")))
print(n - per + best)
PY
)
echo "shell kernel launch index (among k_shell_jk launches): $IDX" > gpurun_out/${R}_prof_idx.log
ncu --set full --clock-control none --import-source on -k regex:k_shell_jk -s $IDX -c 1 -f -o gpurun_out/${R}_prof_shell $BENCH > gpurun_out/ncu_s.log 2>&1 || true
STORED="python bench.py --workload stored:ne2_uhf_ccpvqz --steps 3 --warmup 3"
$STORED > gpurun_out/${R}_stored_plain.json 2> gpurun_out/${R}_stored_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_jk_stored_tma -s 4 -c 1 -f -o gpurun_out/${R}_prof_stored_tma $STORED > gpurun_out/ncu_t.log 2>&1 || true
tail -2 gpurun_out/ncu_s.log gpurun_out/ncu_t.log
