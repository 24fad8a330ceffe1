#!/bin/bash
R=${1:-r02j}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for n in 100 200 400 800; do
  run det1 $n X=1
  run det0 $n TUNA_B200_DETERMINISTIC=0
done
step "reproducibility + direct parity"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize.py tests/test_one_electron.py -m gpu -q -x > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
step "one-electron timing"
timeout 300 python bench.py --extra-one-electron --workload direct:et800 > $O/${R}_oneel.json 2> $O/${R}_oneel.err; step "rc=$? $(python -c "
import json
for x in json.load(open('$O/${R}_oneel.json')): print(x['ncart'], x['kernel_ms'], x['call_ms'])")"
