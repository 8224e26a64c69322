#!/usr/bin/env python
"""f4 measurement: directionality scores on the GPU vs the numpy restatement of the reference's loop on the host."""
import argparse, os, sys, time
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cytopath_b200
from cytopath_b200 import synth
from oracle import directionality_port as dp

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100_000)
ap.add_argument("--k", type=int, default=30)
ap.add_argument("--dims", type=int, default=50)
ap.add_argument("--trajectories", type=int, default=4)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--neighborhood", type=int, default=2000)
ap.add_argument("--cpu-steps", type=int, default=2)
args = ap.parse_args()
rng = np.random.default_rng(0)
g = synth.branching_graph(n=args.cells, k=args.k, n_branches=4, seed=0)
T = g.T.tocsr()
vg = sparse.csr_matrix((rng.random(T.nnz).astype(np.float32), T.indices, T.indptr), shape=T.shape)
vgn = sparse.csr_matrix((-(rng.random(T.nnz) * (rng.random(T.nnz) < 0.3)).astype(np.float32), T.indices, T.indptr), shape=T.shape)
vgn.eliminate_zeros()
pos = rng.normal(size=(args.cells, args.dims))
coords = {"trajectory_%d_coordinates" % i: np.cumsum(rng.normal(size=(args.steps, args.dims)), axis=0) for i in range(args.trajectories)}
nbhd = [[rng.choice(args.cells, size=args.neighborhood, replace=False) for _ in range(args.steps)] for _ in range(args.trajectories)]

class Shim:
    pass
ad = Shim()
ad.uns = {"T_forward": T, "velocity_graph": vg, "velocity_graph_neg": vgn, "trajectories": {"trajectories_coordinates": {"end": coords}}}
small = Shim(); small.uns = dict(ad.uns, trajectories={"trajectories_coordinates": {"end": {"trajectory_0_coordinates": coords["trajectory_0_coordinates"][:3]}}})
cytopath_b200.directionality_score(small, [nbhd[0][:3]], "end", pos)                       # warm-up
t0 = time.perf_counter(); got = cytopath_b200.directionality_score(ad, nbhd, "end", pos); wall = time.perf_counter() - t0
n_inst = args.trajectories * args.steps * args.neighborhood
sub = Shim(); sub.uns = dict(ad.uns, trajectories={"trajectories_coordinates": {"end": {"trajectory_0_coordinates": coords["trajectory_0_coordinates"][:args.cpu_steps + 1]}}})
t0 = time.perf_counter(); want = dp.directionality_score(sub, [nbhd[0][:args.cpu_steps + 1]], "end", pos); cpu = time.perf_counter() - t0
np.testing.assert_allclose(got[0][0], want[0][0], rtol=1e-11, atol=1e-14)
cpu_inst = cpu / ((args.cpu_steps + 1) * args.neighborhood)
print("cells %d k %d dims %d: %d instances: GPU call %.3f s (host preparation of the masked graph included); numpy restatement %.1f us/instance -> %.1f s on one core (x%.0f)" % (
    args.cells, args.k, args.dims, n_inst, wall, cpu_inst * 1e6, cpu_inst * n_inst, cpu_inst * n_inst / wall))
