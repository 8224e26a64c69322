#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_ --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_v2_c4 \
   python scripts/sweep_variants.py --workload c4 --cdf-mode exact --variants 0 --steps 1 > gpurun_out/ncu_v2_c4.log 2>&1
tail -2 gpurun_out/ncu_v2_c4.log
for extra in 0 1 2; do
  CYP_V2_EXTRA=$extra timeout 600 python scripts/sweep_variants.py --workload c4 --cdf-mode exact --variants 0,9532,9528 > gpurun_out/v2_extra${extra}_c4.log 2>&1; tail -4 gpurun_out/v2_extra${extra}_c4.log
done
CYP_V2_EXTRA=1 timeout 600 python scripts/sweep_variants.py --workload c3 --cdf-mode exact --variants 0,9532 > gpurun_out/v2_extra1_c3.log 2>&1; tail -3 gpurun_out/v2_extra1_c3.log
