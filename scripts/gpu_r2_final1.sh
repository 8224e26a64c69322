#!/bin/bash
# round 2, one GPU: whole GPU suite, ncu capture of K2 at a workload (plain run first), then the default bench lines (both arms)
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
if [ -z "$SKIP_TESTS" ]; then
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -6 gpurun_out/pytest_gpu.log
fi
for wl in ${NCU_WORKLOADS-c4}; do NCU_WORKLOAD=$wl NCU_TAG=r02_k2_$wl bash scripts/gpu_ncu_k2.sh; done
for wl in ${LAUNCH_LIST-}; do
python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_plain_$wl.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$wl.log 2>&1; tail -1 gpurun_out/ncu_launch_$wl.log | cut -c1-300
done
for wl in ${BENCH_WORKLOADS:-c4}; do
  (timeout 1200 python bench.py --workload $wl --steps ${BENCH_STEPS:-20} --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_${wl}_n1.json 2> gpurun_out/bench_${wl}_n1.err; echo "bench rc=$?" >> gpurun_out/bench_${wl}_n1.err); tail -4 gpurun_out/bench_${wl}_n1.err; cut -c1-7000 gpurun_out/bench_${wl}_n1.json
done
if [ -n "$REF_ARM" ]; then (timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err); cut -c1-1500 gpurun_out/bench_c4_reference.json; fi
