#!/bin/bash
# Round-2 opening GPU call: the checks round 1 left unrun (nbf 800 full-size parity, prepared variants, DMMA AO->MO, graph mode)
# and source-level ncu captures of three MID angular-momentum class jobs (the regime where per-quartet overhead dominates).
R=${1:-r02a}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }

step "variant sweep"
timeout 500 python tools/direct_timing.py 100,200,400,800 > $O/${R}_variants.log 2>&1; step "rc=$?"

step "fullsize parity (nbf 400 / 800)"
timeout 400 python -m pytest tests/test_zz_fullsize.py -m gpu -q > $O/${R}_zz.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_zz.log)"

step "dmma AO->MO"
TUNA_B200_LIB=$PWD/build/lib_dmma.so timeout 200 python -m pytest tests/test_mo_transform.py -m gpu -q > $O/${R}_dmma.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_dmma.log)"
TUNA_B200_LIB=$PWD/build/lib_dmma.so timeout 120 python tools/mo_quick.py > $O/${R}_dmma_mo.json 2>&1; step "rc=$?"
timeout 120 python tools/mo_quick.py > $O/${R}_simt_mo.json 2>&1; step "rc=$?"

step "graph mode parity"
TUNA_B200_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k direct > $O/${R}_graph.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_graph.log)"

CMD="python tools/direct_timing.py child 800"
for IDX in 116 30 190; do
  step "ncu full capture of class-job launch $IDX"
  TUNA_B200_DUMP_JOBS=$O/${R}_jobs800.csv timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_shell_jk_one -s $IDX -c 1 -f -o $O/${R}_prof_$IDX $CMD > $O/${R}_ncu_$IDX.log 2>&1; step "rc=$?"
  ncu -i $O/${R}_prof_$IDX.ncu-rep --page source --csv > $O/${R}_prof_${IDX}_source.csv 2>/dev/null
  ncu -i $O/${R}_prof_$IDX.ncu-rep --page raw --csv > $O/${R}_prof_${IDX}_raw.csv 2>/dev/null
  rm -f $O/${R}_prof_$IDX.ncu-rep
done
du -sh $O | tee -a $O/${R}_steps.log
