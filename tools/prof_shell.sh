#!/bin/bash
# two-stage ncu capture of the heaviest k_shell_jk launch (development aid)
set -e
CMD="python tools/gsweep.py child ${1:-200}"
$CMD > gpurun_out/prof_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_shell_jk --csv --log-file gpurun_out/shell_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
IDX=$(python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/shell_launches.csv")) if len(r) > 5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
half = len(rows) // 2
best = max(range(half, len(rows)), key=lambda i: float(rows[i][-1].replace(",", "")))
print(best)
PY
)
echo "top launch index $IDX" > gpurun_out/prof_idx.log
ncu --set full --clock-control none --import-source on -k regex:k_shell_jk -s $IDX -c 1 -o gpurun_out/prof_shell_v1 -f $CMD > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
