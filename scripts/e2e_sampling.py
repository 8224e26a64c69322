#!/usr/bin/env python
"""End-to-end `cytopath_b200.sampling()` at atlas scale (C4: 1M cells, k=50) with a wall-clock breakdown."""
import argparse, cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, cytopath_b200

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--traj-number", type=int, default=500)
ap.add_argument("--sim-number", type=int, default=1_250_000)
ap.add_argument("--max-steps", type=int, default=0)
ap.add_argument("--cdf-mode", default="reference")
ap.add_argument("--auto", action="store_true")
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
t0 = time.time()
g, C, roots = bench.make_graph(w, 1)
ad = g.to_adata(matrix_dtype=np.float32)
print("graph + adata: %.1f s; cells %d nnz %d roots %d" % (time.time() - t0, C.shape[0], C.nnz, len(roots)))
kw = dict(auto_adjust=args.auto, traj_number=args.traj_number, sim_number=args.sim_number,
          max_steps=args.max_steps or w["max_steps"], cdf_mode=args.cdf_mode, seed=1)
if not args.auto:
    kw.update(root_clusters=g.root_clusters, end_clusters=g.end_clusters)
np.random.seed(0)
pr = cProfile.Profile() if args.profile else None
t0 = time.time()
if pr: pr.enable()
cytopath_b200.sampling(ad, **kw)
if pr: pr.disable()
dt = time.time() - t0
s = ad.uns["samples"]; b = ad.uns["run_info"]["b200"]
steps = sum(r["n_walkers"] for r in b["rounds"]) * (ad.uns["run_info"]["simulation_steps"] - 1)
print("sampling(): %.2f s wall, %d rounds, %d walker-steps -> %.2f G walker-steps/s end to end; output %s; kernel %s" % (
    dt, len(b["rounds"]), steps, steps / dt / 1e9, s["cell_sequences"].shape, b["kernel"]))
for r in b["rounds"]:
    print("  round: walkers %d terminal %d kept %d walk %.1f ms filter %.1f ms" % (r["n_walkers"], r["n_terminal"], r["n_kept"], r["ms_walk"], r["ms_filter"]))
if pr:
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18); print(st.getvalue()[:3500])
