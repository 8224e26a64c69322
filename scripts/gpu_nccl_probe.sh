#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/nccl_probe.log
run() { echo "== $1" >> gpurun_out/nccl_probe.log; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/sweep_variants.py --workload c4 --cdf-mode exact --variants 0 --steps 10 --init-nccl 2>&1 | grep "variant" >> gpurun_out/nccl_probe.log; }
run "X=1"; run "NCCL_P2P_DISABLE=1"; run "NCCL_NVLS_ENABLE=0"; run "NCCL_CUMEM_ENABLE=0"; run "NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1"
cat gpurun_out/nccl_probe.log
