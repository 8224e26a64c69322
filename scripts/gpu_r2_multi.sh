#!/bin/bash
# round 2, N GPUs of one box: the multi-GPU tests, then bench.py under torchrun (one rank per GPU)
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
N=${N_GPUS:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
if [ -z "$SKIP_TESTS" ]; then
(timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log); tail -15 gpurun_out/pytest_multi.log
fi
if [ -n "$SCALE_PROBE" ]; then bash scripts/gpu_r2_scale_probe.sh; grep "placement" gpurun_out/probe_cal.err | head -12; fi
for wl in ${BENCH_WORKLOADS:-c4}; do
  (timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload $wl --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err; echo "bench rc=$?" >> gpurun_out/bench_${wl}_n$N.err); tail -12 gpurun_out/bench_${wl}_n$N.err; cut -c1-6000 gpurun_out/bench_${wl}_n$N.json
done
