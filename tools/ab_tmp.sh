#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/direct_timing.py child 800 > $O/r02k_direct800.log 2>&1; echo "direct800 rc=$? $(tail -n 1 $O/r02k_direct800.log | cut -c1-120)"
for ks in 1 0; do
for w in n2_ccpvtz ne2_uhf_ccpvqz; do
TUNA_B200_KSPLIT=$ks timeout 200 python bench.py --workload direct:$w --no-stored --steps 20 --warmup 5 > $O/r02k_bench_${w}_ks$ks.json 2> $O/r02k_bench_${w}_ks$ks.err; echo "ks=$ks $w rc=$? $(python -c "import json,sys; d=json.loads(open('$O/r02k_bench_${w}_ks$ks.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['value'], d['config']['parity']['within_tolerance'])")"
done
TUNA_B200_KSPLIT=$ks timeout 200 python tools/fill_timing.py n2_ccpvtz ne2_uhf_ccpvqz > $O/r02k_fill_ks$ks.log 2>&1; echo "fill ks=$ks rc=$?"; cut -c1-200 $O/r02k_fill_ks$ks.log
done
