#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/probe2.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557"
echo "== sweep torchrun nccl sims42 w0=0" >> gpurun_out/probe2.log
$T scripts/sweep_variants.py --workload c4 --cdf-mode exact --variants 0 --steps 10 --init-nccl --sims-per-root 42 2>&1 | grep variant >> gpurun_out/probe2.log
echo "== sweep torchrun NO nccl sims21" >> gpurun_out/probe2.log
$T scripts/sweep_variants.py --workload c4 --cdf-mode exact --variants 0 --steps 10 2>&1 | grep variant >> gpurun_out/probe2.log
echo "== bench torchrun kernel-only" >> gpurun_out/probe2.log
CYP_NO_CLOCKS=1 $T bench.py --gpus 2 --steps 10 --warmup 3 --kernel-only 2>&1 | grep -E "^rank" >> gpurun_out/probe2.log
echo "== bench torchrun kernel-only, walkers 625000 per GPU" >> gpurun_out/probe2.log
CYP_NO_CLOCKS=1 $T bench.py --gpus 2 --steps 10 --warmup 3 --kernel-only --walkers 625000 2>&1 | grep -E "^rank" >> gpurun_out/probe2.log
cat gpurun_out/probe2.log
