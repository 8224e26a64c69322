#!/bin/bash
# GPU box: parity of the paired kernel, then variant sweeps on C4 / C3 / C5.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "k1_float32 or float32_matrix or k2_ or sampling_cuda or k3_" > gpurun_out/pytest_pp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pp.log); tail -5 gpurun_out/pytest_pp.log
for spec in "c4 exact 0,9432,9428,9427,9424,9416,9332" "c4 reference 0,9432,9332" "c3 exact 0,9432,9424,9332" "c5 exact 0,9413,9408,9216"; do
  set -- $spec
  (timeout 600 python scripts/sweep_variants.py --workload $1 --cdf-mode $2 --variants $3 > gpurun_out/pp_sweep_$1_$2.log 2>&1; echo "rc=$?" >> gpurun_out/pp_sweep_$1_$2.log); tail -9 gpurun_out/pp_sweep_$1_$2.log
done
