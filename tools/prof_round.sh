#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full capture of the heaviest direct-mode launch.
set -e
R=${1:-r01}
BENCH="python bench.py --steps 2 --warmup 3 --no-stored"
$BENCH > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv $BENCH > gpurun_out/ncu_l.log 2>&1 || true
IDX=$(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${R}_launches_bench.csv")) if len(r) > 5 and r[0].isdigit() and "k_shell_jk_one" in r[4]]
n = len(rows)
per = n // 8                       # 3 warm-up + 2 timed + 1 e2e warm-up + 2 e2e builds
last = rows[4 * per:5 * per]
best = max(range(len(last)), key=lambda i: float(last[i][-1].replace(",", "")))
print(4 * per + best)
PY
)
echo "k_shell_jk_one launch index: $IDX" > gpurun_out/${R}_prof_idx.log
ncu --set full --clock-control none --import-source on -k regex:k_shell_jk_one -s $IDX -c 1 -f -o gpurun_out/${R}_prof_shell $BENCH > gpurun_out/ncu_s.log 2>&1 || true
tail -n 2 gpurun_out/ncu_s.log
