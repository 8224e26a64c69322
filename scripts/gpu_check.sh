#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, GPU parity tests, benches.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -3 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -${PYTEST_TAIL:-30} gpurun_out/pytest_gpu.log
for wl in ${BENCH_WORKLOADS:-c3 c4}; do
  (timeout 900 python bench.py --workload $wl --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$wl.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$wl.log); tail -3 gpurun_out/bench_$wl.log
done
