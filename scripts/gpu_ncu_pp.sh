#!/bin/bash
mkdir -p gpurun_out
for v in 0 9332; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_ --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_fp250k_v$v \
     python scripts/footprint_sweep.py --cells 250000 --variants $v --steps 1 > gpurun_out/ncu_fp_v$v.log 2>&1
  tail -2 gpurun_out/ncu_fp_v$v.log
done
