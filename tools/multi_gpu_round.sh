#!/bin/bash
# One call on an N-GPU box (gpurun --gpus N): config-4 bench (Ne2 UHF/cc-pVQZ direct) and the even-tempered workloads at N GPUs, plus (with
# a third argument) the hardware tests of the sharded Fock build.  Usage: bash tools/multi_gpu_round.sh <tag> <N> [tests]
R=${1:-r02n}; N=${2:-8}
O=gpurun_out
mkdir -p $O
date +%s > $O/${R}_t0
step() { echo "[$(( $(date +%s) - $(cat $O/${R}_t0) )) s] $*" | tee -a $O/${R}_steps.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701"
if [ -n "$3" ]; then
  step "pytest multi-GPU ($N GPUs)"
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/${R}_pytest_multigpu.log 2>&1; step "rc=$? $(tail -n 1 $O/${R}_pytest_multigpu.log)"
fi
step "bench et800 N=$N"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${R}_bench_et800_n$N.json 2> $O/${R}_bench_et800_n$N.err; step "rc=$?"
step "bench et400 N=$N"
timeout 300 $TR bench.py --gpus $N --workload direct:et400 --steps 10 --warmup 3 > $O/${R}_bench_et400_n$N.json 2> $O/${R}_bench_et400_n$N.err; step "rc=$?"
step "bench ne2 direct N=$N"
timeout 300 $TR bench.py --gpus $N --workload direct:ne2_uhf_ccpvqz --no-stored --steps 20 --warmup 3 > $O/${R}_bench_ne2_n$N.json 2> $O/${R}_bench_ne2_n$N.err; step "rc=$?"
step "bench ne2 direct N=1"
timeout 300 python bench.py --workload direct:ne2_uhf_ccpvqz --no-stored --steps 20 --warmup 3 > $O/${R}_bench_ne2_n1.json 2> $O/${R}_bench_ne2_n1.err; step "rc=$?"
