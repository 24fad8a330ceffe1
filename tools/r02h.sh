#!/bin/bash
R=${1:-r02h}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
step "parity (gen4)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize.py -m gpu -x -q > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for n in 100 200 400 800; do
  run base $n X=1
  run term6k $n TUNA_B200_TERM_MAX=6000
  run term12k $n TUNA_B200_TERM_MAX=12000
  run nb1g32 $n TUNA_B200_NB=1 TUNA_B200_SMEM_PER_LANE=2048
  run nb2g32 $n TUNA_B200_SMEM_PER_LANE=2048
  run nb2spl1k $n TUNA_B200_SMEM_PER_LANE=1024
done
for m in 1 2 4 32 64 128 255; do run skip$m 800 TUNA_B200_DBG_SKIP=$m; done
