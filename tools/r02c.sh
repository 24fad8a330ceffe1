#!/bin/bash
# source-level ncu captures of selected class jobs (launch indices in job order), CUDA-C and SASS views exported on the box
R=${1:-r02c}; shift
O=gpurun_out
mkdir -p $O
CMD="python tools/direct_timing.py child 800"
for IDX in "$@"; do
  TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_shell4_one -s $IDX -c 1 -f -o $O/${R}_prof_$IDX $CMD > $O/${R}_ncu_$IDX.log 2>&1; echo "idx $IDX rc=$?"
  ncu -i $O/${R}_prof_$IDX.ncu-rep --page source --print-source cuda --csv > $O/${R}_prof_${IDX}_cuda.csv 2>/dev/null
  ncu -i $O/${R}_prof_$IDX.ncu-rep --page source --csv > $O/${R}_prof_${IDX}_sass.csv 2>/dev/null
  ncu -i $O/${R}_prof_$IDX.ncu-rep --page raw --csv > $O/${R}_prof_${IDX}_raw.csv 2>/dev/null
  rm -f $O/${R}_prof_$IDX.ncu-rep
done
du -sh $O
