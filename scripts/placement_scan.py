#!/usr/bin/env python
"""K2 time against the position of the fused rows in memory (CYP_ROWS2_OFFSET_KB): which L2 slices serve the hot rows
depends on it.  One process, one graph generation; the device graph is rebuilt per offset.  GPU box only."""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cytopath_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--cdf-mode", default="exact")
ap.add_argument("--offsets-kb", default="0,4,16,64,128,192,256,512,1024,2048,4096,65536")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
g, C, roots = bench.make_graph(w, 1)
mode = _lib.CDF_EXACT if args.cdf_mode == "exact" else _lib.CDF_REFERENCE
L = _lib.load()
W, nt = w["walkers"], w["max_steps"] - 1
sims = -(-W // len(roots))
d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
d_seq = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
times = []
for off in [int(x) for x in args.offsets_kb.split(",")]:
    os.environ["CYP_ROWS2_OFFSET_KB"] = str(off)
    graph = _lib.Graph(C, mode, 0)
    def step(i):
        _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, i, None, d_seq.data_ptr(), None, 0, sptr))
    step(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps): step(1 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    times.append(ms)
    print("%s %s offset %6d KB: %6.2f ms  %6.2f G walker-steps/s" % (args.workload, args.cdf_mode, off, ms, W * nt / ms / 1e6), flush=True)
    graph.close()
print("min %.2f ms max %.2f ms (x%.2f)" % (min(times), max(times), max(times) / min(times)))
