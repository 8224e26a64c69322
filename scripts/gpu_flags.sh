#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/flags.log
run() { echo "== $1" >> gpurun_out/flags.log; env $1 timeout 300 python scripts/sweep_variants.py --workload ${WL:-c4} --cdf-mode exact --variants 0 --steps 5 2>&1 | grep -E "variant|cyp" >> gpurun_out/flags.log; }
for f in "CYP_V2_FLAGS=0" "CYP_V2_FLAGS=1" "CYP_V2_FLAGS=2" "CYP_V2_FLAGS=4" "CYP_V2_FLAGS=3" "CYP_V2_FLAGS=7" "CYP_V2_PERSIST=0.3" "CYP_V2_PERSIST=0.5" "CYP_V2_PERSIST=0.7" "CYP_V2_PERSIST=1.0" "CYP_V2_FLAGS=0"; do run "$f"; done
cat gpurun_out/flags.log
