#!/usr/bin/env python
"""K2 on a graph WITHOUT locality or hot rows (every row: k random targets, skewed or flat probabilities), at C4's
footprint: separates what the access skew of the velocity graph costs from what the kernel itself can do.  GPU box only."""
import argparse, ctypes, os, sys
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cytopath_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=51)
ap.add_argument("--walkers", type=int, default=1_250_000)
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--skew", type=float, default=3.0, help="probabilities = exp(skew * N(0,1)); 0 = flat")
ap.add_argument("--variants", default="0,9332")
args = ap.parse_args()
rng = np.random.default_rng(0)
n, k = args.cells, args.k
cols = np.sort(rng.integers(0, n, size=(n, k), dtype=np.int64), axis=1)
cols += np.arange(k)[None, :] * 0                                    # duplicates are summed by the CSR constructor
vals = np.exp(args.skew * rng.normal(size=(n, k)))
vals /= vals.sum(axis=1, keepdims=True)
C = sparse.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), np.arange(0, n * k + 1, k, dtype=np.int64)), shape=(n, n))
C.sum_duplicates(); C.sort_indices()
rs = np.asarray(C.sum(axis=1)).ravel()
C = sparse.diags(1.0 / rs) @ C
C = C.tocsr(); C.sort_indices()
graph = _lib.Graph(C, _lib.CDF_REFERENCE, 0)            # reference mode: rounding closes the 1e-16 gaps of the renormalisation
L = _lib.load()
roots = rng.choice(n, size=60_000, replace=False).astype(np.int32)
W, nt = args.walkers, args.steps - 1
sims = -(-W // len(roots))
d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
d_seq = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("cells %d nnz %d skew %.1f layout %s" % (n, C.nnz, args.skew, graph.layout()))
for v in [int(x) for x in args.variants.split(",")]:
    def step(i):
        _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, i, None, d_seq.data_ptr(), None, v, sptr))
    step(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3): step(1 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("variant %5d %-70s %7.2f ms %7.2f G walker-steps/s" % (v, graph.variant_name(v), ms, W * nt / ms / 1e6), flush=True)
