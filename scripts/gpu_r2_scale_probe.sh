#!/bin/bash
# N GPUs: K2-only bench under different group set-ups (why are the non-root ranks slower at N=8?)
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
N=${N_GPUS:-8}
run() { tag=$1; shift; (env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 10 --warmup 3 --kernel-only > gpurun_out/probe_$tag.json 2> gpurun_out/probe_$tag.err); echo "== $tag"; grep "ms per step" gpurun_out/probe_$tag.err | tr '\n' ' ' | sed 's/(device time of its own K2 launches)//g'; echo; cut -c1-200 gpurun_out/probe_$tag.json; }
run nocal CYP_NO_CALIBRATE=1
run cal CYP_TRACE=1
