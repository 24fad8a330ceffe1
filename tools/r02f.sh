#!/bin/bash
R=${1:-r02f}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for n in 400 800; do
  for m in 0 1 2 4 8 16 32 64 128 255 254 126 96 160; do run g4skip$m $n TUNA_B200_DBG_SKIP=$m; done
  for m in 0 64 1 2 4 8 16 32 128 255; do run g2skip$m $n TUNA_B200_ENGINE=2 TUNA_B200_DBG_SKIP=$m; done
done
step "new hardware tests: reference driver on the provider, AO->MO incl. device spin blocking"
timeout 900 python -m pytest tests/test_mo_transform.py tests/test_reference_on_gpu.py -m gpu -q > $O/${R}_pytest_new.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_new.log)"
