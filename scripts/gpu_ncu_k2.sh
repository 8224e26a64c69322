#!/bin/bash
# One ncu --set full capture of the K2 launch of bench.py --kernel-only (plain run first), report into gpurun_out/.
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
WL=${NCU_WORKLOAD:-c4}; TAG=${NCU_TAG:-k2}
# the placement calibration launches the walk kernel up to four times before the warm-up: skip well into the warm-up
CMD="python bench.py --workload $WL --steps 4 --warmup 6 --kernel-only --no-cpu ${NCU_BENCH_ARGS:-}"
$CMD > gpurun_out/ncu_${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_ --launch-skip 8 --launch-count 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}_plain.log; tail -3 gpurun_out/ncu_${TAG}.log | cut -c1-300
