#!/bin/bash
R=${1:-r02i}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for n in 400 800; do
  run base $n X=1
  run tier0 $n TUNA_B200_REG_TIER=0
done
step "full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py > $O/${R}_pytest_gpu.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_gpu.log)"
step "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_smoke.log)"
step "one-electron timing"
timeout 300 python bench.py --extra-one-electron --workload direct:et800 > $O/${R}_oneel.json 2> $O/${R}_oneel.err; step "rc=$? $(cut -c1-600 $O/${R}_oneel.json)"
