#!/bin/bash
# same-box A/B of two library builds (scripts/bin/lib_A.so, lib_B.so)
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do for L in A B; do for wl in ${AB_WORKLOADS:-c4 c3}; do
  echo "== lib_$L $wl" >> gpurun_out/ab.log
  CYP_LIB_PATH=$PWD/scripts/bin/lib_$L.so timeout 300 python scripts/sweep_variants.py --workload $wl --cdf-mode ${AB_MODE:-exact} --variants ${AB_VARIANTS:-0,9532} --steps 5 2>&1 | grep variant >> gpurun_out/ab.log
done; done; done
cat gpurun_out/ab.log
