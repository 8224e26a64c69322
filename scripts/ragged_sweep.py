#!/usr/bin/env python
"""K2 on a ragged graph (most rows 8..40 entries, 2 % hubs of 150..300): fused rows with a slot for the typical row
(hubs searched in HBM) against a slot for the longest row and against the 16-bit-row kernel.  GPU box only."""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cytopath_b200 import _lib, synth
from cytopath_b200.sampler import _canonical_csr

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100_000)
ap.add_argument("--walkers", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--variants", default="0,9332,9016")
args = ap.parse_args()
C = _canonical_csr(synth.hub_graph(n=args.cells, seed=1))
graph = _lib.Graph(C, _lib.CDF_EXACT, 0)
L = _lib.load()
roots = np.arange(0, args.cells, 97, dtype=np.int32)
W, nt = args.walkers, args.steps - 1
sims = -(-W // len(roots))
d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
d_seq = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("cells %d nnz %d max row %d layout %s" % (args.cells, C.nnz, int(np.diff(C.indptr).max()), graph.layout()))
ref = None
for v in [int(x) for x in args.variants.split(",")]:
    def step(i):
        _lib.check(L.cyp_walk_device(graph.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, i, None, d_seq.data_ptr(), None, v, sptr))
    step(0); torch.cuda.synchronize()
    first = d_seq[:2000].cpu().numpy()
    if ref is None: ref = first
    assert np.array_equal(ref, first), "variants disagree"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): step(1 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("variant %5d %-75s %7.2f ms %7.2f G walker-steps/s" % (v, graph.variant_name(v), ms, W * nt / ms / 1e6), flush=True)
