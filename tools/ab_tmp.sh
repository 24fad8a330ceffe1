#!/bin/bash
R=r02r; O=gpurun_out; mkdir -p $O; date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
run() { local name=$1 n=$2; shift 2
  env "$@" timeout 200 python tools/direct_timing.py child $n > $O/${R}_k_${n}_$name.json 2> $O/${R}_k_${n}_$name.err; step "nbf $n $name: $(cut -c1-32 $O/${R}_k_${n}_$name.json) $(tail -c 150 $O/${R}_k_${n}_$name.err)"
}
for u in 2 0.5 8; do
  for n in 100 200 400; do run units$u $n TUNA_B200_OWN_LAUNCH_UNITS=$u; done
  for wl in ne2_uhf_ccpvqz n2_ccpvtz; do
    TUNA_B200_OWN_LAUNCH_UNITS=$u timeout 300 python bench.py --workload direct:$wl --no-stored --steps 20 --warmup 3 > $O/${R}_${wl}_$u.json 2> $O/${R}_${wl}_$u.err; step "$wl units $u: $(python -c "import json;d=json.loads(open('$O/${R}_${wl}_$u.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['gpu_launches'], d['config']['parity']['within_tolerance'])")"
  done
done
run base 800 X=1
