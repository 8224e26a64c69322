#!/bin/bash
# GPU box: smoke, whole GPU suite, bench (both arms), ncu launch list + full capture of K2 under bench.py
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -3 gpurun_out/smoke.log
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -8 gpurun_out/pytest_gpu.log
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?" >> gpurun_out/bench_c4.err); tail -2 gpurun_out/bench_c4.err; cut -c1-1500 gpurun_out/bench_c4.json
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err; echo "ref rc=$?" >> gpurun_out/bench_c4_reference.err); cut -c1-600 gpurun_out/bench_c4_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_ --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_v2_bench_c4 python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
