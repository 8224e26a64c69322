#!/usr/bin/env python
"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
K1, K2 (register, bulk, cp.async flavours; both CDF modes; filter modes), K3 (dedup with forced
hash collisions), scores, K4.  Sizes are tiny: the sanitizer slows kernels by 10-100x."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytopath_b200 import _lib, synth
from oracle import cwalk, sampler_port as sp

L = _lib.load()
gph = synth.branching_graph(n=400, k=20, n_branches=3, seed=1, workers=1, scale=4.0, noise=0.03)
C = sp.canonical_csr(gph.T)
roots = np.where(gph.clusters == "stem0")[0][:8]
for mode in (0, 1):
    g = _lib.Graph(C, mode)
    cdf = cwalk.build_cdf(C, mode)
    assert np.array_equal(g.cdf(), cdf)
    want = cwalk.walk(C, cdf, np.repeat(roots, 9), 30, seed=5)
    d_roots = np.ascontiguousarray(roots, dtype=np.int32)
    assert np.array_equal(g.walk(roots, 9, 30, seed=5), want)
    import torch
    t_roots = torch.tensor(d_roots, device="cuda")
    for variant in (404, 804, 9004, 9204, 9304, 9308):
        out = torch.zeros((72, 31), dtype=torch.int32, device="cuda")
        _lib.check(L.cyp_walk_device(g.handle, t_roots.data_ptr(), len(roots), 9, 0, 72, 30, 5, 0, None, out.data_ptr(), None,
                                     variant, None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want), variant
    code = np.unique(np.array([c[:8] for c in gph.clusters]), return_inverse=True)[1].astype(np.int32)
    ends = np.where(np.isin(gph.clusters, gph.end_clusters))[0]
    slot = np.full(C.shape[0], -1, np.int32); slot[ends] = np.arange(len(ends))
    s = _lib.Session(g, roots, code, int(code.max()) + 1, 2, slot, len(ends), 40, True, 7)
    s.set_hash_bits(4)
    st = s.round(0, 40, 0, len(roots) * 40)
    rows = s.fetch_rows(np.arange(s.rows()))
    sc = g.transition_scores(rows) if len(rows) else None
    s.close(); g.close()
    print("mode", mode, "kept", st["n_kept"], "passes", st["dedup_passes"])
eps = _lib.power_iteration(gph.T, gph.root_prob / gph.root_prob.sum(), gph.end_prob / gph.end_prob.sum(), 10)
print("sanitize smoke ok", eps[-1])
