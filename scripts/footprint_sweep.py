#!/usr/bin/env python
"""K2 throughput against graph footprint: C4-shaped graphs (k=50, 16 layers) of 125k .. 1M cells, same walkers and steps.
Shows how much of the C4 rate is lost to DRAM random gathers (the L2 holds 126 MB).  GPU box only."""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cytopath_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--cells", default="125000,250000,500000,1000000")
ap.add_argument("--variants", default="0,9332")
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--base", default="c4")
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
L = _lib.load()
for n in [int(x) for x in args.cells.split(",")]:
    w = dict(bench.WORKLOADS[args.base], n=n, k=args.k)
    if args.walkers: w["walkers"] = args.walkers
    g, C, roots = bench.make_graph(w, 1)
    graph = _lib.Graph(C, _lib.CDF_EXACT, 0)
    W, nt = w["walkers"], w["max_steps"] - 1
    sims = -(-W // len(roots))
    d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
    d_seq = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    info = graph.info()
    for v in [int(x) for x in args.variants.split(",")]:
        def step(i):
            _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, i, None,
                                         d_seq.data_ptr(), None, v, sptr))
        step(0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps): step(1 + i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print("cells %8d nnz %9d roots %6d device_bytes %6.0f MB  variant %5d %-70s %7.2f ms %7.2f G walker-steps/s" % (
            n, C.nnz, len(roots), info["device_bytes"] / 1e6, v, graph.variant_name(v), ms, W * nt / ms / 1e6), flush=True)
    graph.close(); del d_seq
