#!/bin/bash
# round 2: whole GPU suite (optionally filtered), a small bench line, then the K2 A/B of scripts/gpu_r2_k2.sh
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
(timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} -k "${K_EXPR:-not full_size}" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -${PYTEST_TAIL:-12} gpurun_out/pytest_gpu.log
for wl in ${BENCH_WORKLOADS:-tiny}; do
  (timeout 900 python bench.py --workload $wl --steps ${BENCH_STEPS:-4} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?" >> gpurun_out/bench_$wl.err); tail -5 gpurun_out/bench_$wl.err; cut -c1-${BENCH_CUT:-3000} gpurun_out/bench_$wl.json
done
if [ -n "$AB_SPECS" ]; then K2_TESTS="nothing_selected" bash scripts/gpu_r2_k2.sh | tail -40; fi
