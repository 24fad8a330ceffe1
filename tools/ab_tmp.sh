#!/bin/bash
R=r02l; O=gpurun_out; mkdir -p $O; date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for n in 100 200 400 800; do
  run base $n X=1
  run hb0 $n TUNA_B200_HDR_BOYS=0
  run st6 $n TUNA_B200_STREAMS=6
  run st16 $n TUNA_B200_STREAMS=16
done
step "parity"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize.py -m gpu -q -x > $O/${R}_pytest.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest.log)"
