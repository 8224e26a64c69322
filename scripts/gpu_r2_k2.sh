#!/bin/bash
# round 2, K2 rewrite: parity of the fused-row kernel, then same-box A/B of two builds (scripts/bin/lib_A.so = round-1 kernel,
# lib_B.so = candidate) on C4 (both CDF modes), C3 and the no-locality probe graphs.
mkdir -p gpurun_out; export BENCH_GRAPH_CACHE=/tmp/cyp_graphs
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "${K2_TESTS:-k1b or k2 or sampling_cuda or partition or k3}" > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k2.log); tail -5 gpurun_out/pytest_k2.log
: > gpurun_out/ab.log
IFS=';' read -ra SPECS <<< "${AB_SPECS:-c4 exact;c4 reference;c3 exact}"
for rep in 1 2; do for spec in "${SPECS[@]}"; do set -- $spec; for L in ${AB_LIBS:-A B}; do
  echo "== lib_$L $1 $2" >> gpurun_out/ab.log
  CYP_LIB_PATH=$PWD/scripts/bin/lib_$L.so timeout 300 python scripts/sweep_variants.py --workload $1 --cdf-mode $2 --variants ${AB_VARIANTS:-0} --steps 5 2>&1 | grep -E "variant|rror" >> gpurun_out/ab.log
done; done; done
for skew in 3 0; do for L in ${AB_LIBS:-A B}; do
  echo "== lib_$L uniform graph skew $skew" >> gpurun_out/ab.log
  CYP_LIB_PATH=$PWD/scripts/bin/lib_$L.so timeout 300 python scripts/uniform_graph_probe.py --skew $skew --variants 0 2>&1 | grep -E "variant|rror" >> gpurun_out/ab.log
done; done
cat gpurun_out/ab.log
