#!/bin/bash
R=r02z; O=gpurun_out; mkdir -p $O; date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
run base 800 X=1
run nbbytes140 800 TUNA_B200_NB_BYTES=143360
run nbbytes200 800 TUNA_B200_NB_BYTES=204800
run term2k 800 TUNA_B200_TERM_MAX=2048
run term4k 800 TUNA_B200_TERM_MAX=4096
run itb4096 800 TUNA_B200_IT_BUDGET=4096
run sb6144 800 TUNA_B200_S_BUDGET=6144
run spl384 800 TUNA_B200_SMEM_PER_LANE=384
run spl192 800 TUNA_B200_SMEM_PER_LANE=192
run tier0 800 TUNA_B200_TIER2=0
