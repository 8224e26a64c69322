#!/bin/bash
# GPU box: parity of the K2 flavours, then variant sweeps.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_pp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pp.log); tail -25 gpurun_out/pytest_pp.log
IFS=';' read -ra SPECS <<< "${SWEEPS:-c4 exact 0,9532,9524,9332;c4 reference 0,9332;c3 exact 0,9332;c5 exact 0,9216}"
for spec in "${SPECS[@]}"; do
  set -- $spec
  (timeout 600 python scripts/sweep_variants.py --workload $1 --cdf-mode $2 --variants $3 > gpurun_out/v2_sweep_$1_$2.log 2>&1; echo "rc=$?" >> gpurun_out/v2_sweep_$1_$2.log); tail -9 gpurun_out/v2_sweep_$1_$2.log
done
