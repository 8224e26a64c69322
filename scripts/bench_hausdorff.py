#!/usr/bin/env python
"""f2 measurement: Hausdorff distance matrix on the GPU vs the numpy restatement on the host."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytopath_b200 import _lib
from oracle import hausdorff_port as hp

ap = argparse.ArgumentParser()
ap.add_argument("--S", type=int, default=714)      # config 1: 2,856 output rows / 4 terminal clusters
ap.add_argument("--L", type=int, default=112)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--cpu-pairs", type=int, default=300)
args = ap.parse_args()
rng = np.random.default_rng(0)
coords = (rng.normal(size=(args.S, 1, args.d)) + np.cumsum(rng.normal(scale=0.05, size=(args.S, args.L, args.d)), axis=1))
_lib.hausdorff_matrix(coords[:8])
t0 = time.perf_counter(); D, ms = _lib.hausdorff_matrix(coords, return_ms=True); wall = time.perf_counter() - t0
pairs = args.S * (args.S - 1) / 2
flops = pairs * 2 * args.L * args.L * args.d * 3.0
t0 = time.perf_counter()
for i in range(args.cpu_pairs):
    hp.hausdorff_distance(coords[i % args.S], coords[(i * 7 + 1) % args.S])
cpu_pair = (time.perf_counter() - t0) / args.cpu_pairs
print("S=%d L=%d d=%d: GPU kernel %.1f ms (%.2f TFLOP/s fp64), call %.3f s; numpy restatement %.2f ms/pair -> %.1f s for the matrix on one core (x%.0f)" % (
    args.S, args.L, args.d, ms, flops / ms / 1e9, wall, cpu_pair * 1e3, cpu_pair * pairs, cpu_pair * pairs / wall))
