#!/usr/bin/env python
"""Times every K2 variant (lanes x unroll) on one workload; prints a table.  GPU box only."""
import argparse, ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cytopath_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--cdf-mode", default="exact")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--variants", default="404,804,1604,3204")
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--sims-per-root", type=int, default=0)
ap.add_argument("--walker-begin", type=int, default=0)
ap.add_argument("--pad-mb", type=int, default=0, help="allocate this much device memory first (shifts every later allocation)")
ap.add_argument("--barrier", action="store_true", help="dist.barrier() before every timed block (with --init-nccl)")
ap.add_argument("--init-nccl", action="store_true", help="initialise a 1-rank NCCL group first (does the communicator change K2's speed?)")
args = ap.parse_args()
if args.init_nccl:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    rk, lrk, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(lrk)
    dist.init_process_group("nccl", rank=rk, world_size=ws, device_id=torch.device("cuda", lrk))
    x = torch.ones(4, device="cuda"); dist.all_reduce(x); torch.cuda.synchronize()
if args.pad_mb:
    _pad = torch.empty(args.pad_mb << 20, dtype=torch.uint8, device="cuda")
    import ctypes as _ct                                     # ... including the library's (stream-ordered pool)
    _pp = _ct.c_void_p(); _lib.load(); _rt = _ct.CDLL("libcudart.so.12"); _rt.cudaMallocAsync(_ct.byref(_pp), _ct.c_size_t(args.pad_mb << 20), None)
w = bench.WORKLOADS[args.workload]
g, C, roots = bench.make_graph(w, 1)
mode = _lib.CDF_EXACT if args.cdf_mode == "exact" else _lib.CDF_REFERENCE
dev_index = torch.cuda.current_device()
graph = _lib.Graph(C, mode, dev_index)
L = _lib.load()
W = args.walkers or w["walkers"]; nt = w["max_steps"] - 1
sims = args.sims_per_root or -(-W // len(roots))
w0 = args.walker_begin
d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
d_seq = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
elem = 8 if mode == _lib.CDF_EXACT else 4
bps = elem * (C.nnz / C.shape[0]) + 16
print("workload %s mode %s W=%d nt=%d build_ms=%.2f" % (args.workload, args.cdf_mode, W, nt, graph.build_ms()))
print("d_seq at 0x%x roots at 0x%x" % (d_seq.data_ptr(), d_roots.data_ptr()))
for v in [int(x) for x in args.variants.split(",")]:
    def step(i):
        _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), len(roots), sims, w0, W, nt, 1234, i, None,
                                     d_seq.data_ptr(), None, v, sptr))
    step(0); torch.cuda.synchronize()
    if args.barrier and args.init_nccl:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps): step(1 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    rate = W * nt / ms / 1e6
    print("variant %5d  %8.2f ms  %7.3f G walker-steps/s  %6.1f GB/s algorithmic (%.1f%% of 6546.6)" % (v, ms, rate, rate * bps, rate * bps / 65.466))
