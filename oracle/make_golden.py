"""Generate tests/golden/*.npz by running the REAL reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

It loads /root/reference/cytopath/markov_chain_sampler.py by path (the package itself
cannot be imported: fastdtw/anndata/scvelo are absent), replaces
`np.random.rand` with the frozen Philox stream (oracle/philox.py) for the duration of
each call, seeds the global numpy RNG for the final with-replacement draw (:395) and
stores inputs + outputs.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import os
os.environ.setdefault("TQDM_DISABLE", "1")
import contextlib
import importlib.util
import io
import json
import os
import sys
import warnings

import numpy as np

from cytopath_b200 import synth
from oracle import philox

REF = "/root/reference/cytopath/markov_chain_sampler.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("cytopath_ref_markov_chain_sampler", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class PhiloxRand:
    """Stand-in for np.random.rand(a, b).  With num_cores=1 the reference calls it once
    per root cell, in root order, every round (:302-308); call number c therefore maps
    to round c // n_roots and walker rows [(c % n_roots) * a, ... + a)."""

    def __init__(self, seed, n_roots):
        self.seed, self.n_roots, self.calls = seed, n_roots, 0

    def __call__(self, a, b):
        rnd, j = divmod(self.calls, self.n_roots)
        self.calls += 1
        return philox.walker_uniforms(self.seed, rnd, j * a, a, b)


@contextlib.contextmanager
def patched_rand(fn):
    keep = np.random.rand
    np.random.rand = fn
    try:
        yield
    finally:
        np.random.rand = keep


def csr_fields(prefix, M):
    return {prefix + "data": M.data, prefix + "indices": M.indices, prefix + "indptr": M.indptr,
            prefix + "shape": np.asarray(M.shape)}


def run_sampling_case(ref, name, graph, kwargs, seed, np_seed, matrix_dtype=None, cluster_key="louvain",
                      row_scale_seed=None):
    ad = graph.to_adata(cluster_key=cluster_key, matrix_dtype=matrix_dtype)
    if row_scale_seed is not None:          # un-normalise the rows so normalize=True has work to do
        from scipy import sparse
        f = np.random.default_rng(row_scale_seed).uniform(0.5, 2.0, size=ad.shape[0])
        ad.uns["T_forward"] = sparse.csr_matrix(sparse.diags(f) @ ad.uns["T_forward"])
    # first pass only to learn the number of root cells the reference resolves
    probe = ad.copy()
    n_roots = None
    kw = dict(kwargs, cluster_key=cluster_key)

    class Stop(Exception):
        pass

    def probe_rand(a, b):
        raise Stop()

    with patched_rand(probe_rand), contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref.sampling(probe, **kw)
        except Stop:
            n_roots = len(probe.uns["run_info"]["root_cells"])
    assert n_roots, "could not resolve root cells"
    rand = PhiloxRand(seed, n_roots)
    np.random.seed(np_seed)
    log = io.StringIO()
    with patched_rand(rand), contextlib.redirect_stdout(log), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref.sampling(ad, **kw)
    s = ad.uns["samples"]
    ri = ad.uns["run_info"]
    rounds = rand.calls // n_roots
    out = dict(
        csr_fields("T_", ad.uns["T_forward"]),
        clusters=np.asarray(graph.clusters, dtype=str),
        root_prob=graph.root_prob, end_prob=graph.end_prob,
        kwargs=np.asarray(json.dumps(kw)), seed=np.asarray(seed), np_seed=np.asarray(np_seed),
        cell_sequences=s["cell_sequences"], transition_score=s["transition_score"],
        cluster_sequences=s["cluster_sequences"],
        run_root_cells=np.asarray(ri["root_cells"]), run_end_points=np.asarray(ri["end_points"]),
        simulation_steps=np.asarray(ri["simulation_steps"]), rounds=np.asarray(rounds),
    )
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s rows=%d steps=%d rounds=%d roots=%d" % (
        name, s["cell_sequences"].shape[0], s["cell_sequences"].shape[1], rounds, n_roots))


def load_reference_utils():
    """cytopath/utils.py needs `anndata` (absent) only for a module-level import and the package's
    __init__ pulls in fastdtw/hausdorff; load it as cytopath.utils under a stub package instead."""
    import importlib
    import types
    if "anndata" not in sys.modules:
        stub = types.ModuleType("anndata")
        stub.AnnData = object                      # only used in type annotations (utils.py:148 ff.)
        sys.modules["anndata"] = stub
    pkg = types.ModuleType("cytopath")
    pkg.__path__ = [os.path.dirname(REF)]
    sys.modules["cytopath"] = pkg
    try:
        return importlib.import_module("cytopath.utils")
    finally:
        sys.modules.pop("cytopath", None)


def run_undirected_case(name, graph, kwargs, seed, np_seed):
    utils = load_reference_utils()
    ad = graph.to_adata()
    n = ad.shape[0]
    sim_number = kwargs.get("sim_number") or n
    rand = PhiloxRand(seed, min(sim_number, n))
    np.random.seed(np_seed)
    with patched_rand(rand), contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        utils.undirected_simulations(ad, **kwargs)
    s = ad.uns["undirected_samples"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **csr_fields("T_", ad.uns["T_forward"]),
                        clusters=np.asarray(graph.clusters, dtype=str), root_prob=graph.root_prob, end_prob=graph.end_prob,
                        kwargs=np.asarray(json.dumps(kwargs)), seed=np.asarray(seed), np_seed=np.asarray(np_seed),
                        log_terminal_state_freq=np.asarray(ad.obs["log_terminal_state_freq"], dtype=np.float64),
                        cell_sequences=s["cell_sequences"], transition_score=s["transition_score"],
                        cluster_sequences=s["cluster_sequences"])
    print("%-28s rows=%d steps=%d" % (name, s["cell_sequences"].shape[0], s["cell_sequences"].shape[1]))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    und = synth.branching_graph(n=300, k=10, n_branches=2, seed=6, workers=1, scale=3.0, noise=0.03)
    run_undirected_case("undirected_all_cells", und, dict(max_steps=6), seed=21, np_seed=12)
    run_undirected_case("undirected_subset", und, dict(max_steps=8, sim_number=120, unique=False), seed=22, np_seed=13)
    if "--undirected-only" in sys.argv:
        return

    # ---- matrix_prep on a messy CSR (unsorted, duplicates, explicit zeros), float64 and float32
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        M = synth.random_sparse_rows(n=200, max_nnz=12, seed=3, messy=True, dtype=dt)
        dense = np.float64(np.asarray(M.toarray()))                     # what sampling hands over (:168)
        idx, cdf = ref.matrix_prep(dense)
        np.savez_compressed(os.path.join(OUT, "matrix_prep_messy_%s.npz" % tag),
                            idx=idx, cdf=cdf, **csr_fields("T_", M))
    print("matrix_prep_messy_*          done")

    # ---- markov_sim on a small graph, two roots, Philox uniforms
    g = synth.branching_graph(n=400, k=10, n_branches=3, seed=1, workers=1)
    dense = np.float64(g.T.toarray())
    idx, cdf = ref.matrix_prep(dense)
    roots = np.array([5, 17], dtype=np.int64)
    sim, steps, seed = 64, 40, 99
    seqs, scores, labels = [], [], []
    for j in range(len(roots)):
        with patched_rand(lambda a, b, j=j: philox.walker_uniforms(seed, 0, j * a, a, b)):
            st, sc, cl = ref.markov_sim(j, sim, steps, roots, g.clusters.astype(object), dense, idx, cdf)
        seqs.append(st); scores.append(sc); labels.append(cl)
    np.savez_compressed(os.path.join(OUT, "markov_sim_small.npz"), roots=roots, sim=np.asarray(sim),
                        steps=np.asarray(steps), seed=np.asarray(seed), clusters=g.clusters,
                        seq=np.vstack(seqs), score=np.concatenate(scores), labels=np.vstack(labels),
                        idx=idx, cdf=cdf, **csr_fields("T_", g.T))
    print("markov_sim_small             done")

    # ---- check_convergence_criteria probes (:54-60)
    probes = [np.zeros(10), np.arange(10.0), np.array([5, 3, 2, 1.5, 1.5, 1.5, 1.5]),
              np.array([1.0, 0.5, 0.4999, 0.4998, 0.2]), np.array([0.3])]
    res = [ref.check_convergence_criteria(p, tol=1e-3) for p in probes]
    np.savez_compressed(os.path.join(OUT, "convergence_probes.npz"),
                        **{"p%d" % i: p for i, p in enumerate(probes)},
                        result=np.asarray([-1 if r is None else r for r in res]))
    print("convergence_probes           ", res)

    # ---- full sampling() cases
    small = synth.branching_graph(n=600, k=16, n_branches=3, seed=2, workers=1, scale=4.0, noise=0.03)
    run_sampling_case(ref, "sampling_small_clusters", small,
                      dict(auto_adjust=False, max_steps=70, traj_number=200, sim_number=100,
                           root_clusters=small.root_clusters, end_clusters=small.end_clusters, num_cores=1),
                      seed=11, np_seed=5)
    run_sampling_case(ref, "sampling_small_probs_f32", small,
                      dict(auto_adjust=False, max_steps=70, traj_number=40, sim_number=200, min_clusters=3,
                           num_cores=1),
                      seed=12, np_seed=6, matrix_dtype=np.float32)
    run_sampling_case(ref, "sampling_small_notunique", small,
                      dict(auto_adjust=False, max_steps=60, traj_number=50, sim_number=500, unique=False,
                           root_clusters=small.root_clusters, end_clusters=small.end_clusters[:2], num_cores=1),
                      seed=13, np_seed=7, cluster_key="leiden")
    root_cell = int(np.where(small.clusters == "stem0")[0][0])
    ends = np.where(np.isin(small.clusters, small.end_clusters))[0][::-1][:40]      # unsorted explicit end points
    run_sampling_case(ref, "sampling_small_single_root", small,
                      dict(auto_adjust=False, max_steps=80, traj_number=30, sim_number=100,
                           root_cells=[root_cell], end_points=ends.tolist(), num_cores=1),
                      seed=14, np_seed=8)
    dup_ends = np.where(np.isin(small.clusters, small.end_clusters))[0]
    dup_ends = np.concatenate([dup_ends[::-1][:25], dup_ends[:10], dup_ends[::-1][:5]])         # unsorted, with repeats
    run_sampling_case(ref, "sampling_small_dup_endpoints", small,
                      dict(auto_adjust=False, max_steps=70, traj_number=40, sim_number=300,
                           root_clusters=small.root_clusters, end_points=dup_ends.tolist(), num_cores=1),
                      seed=18, np_seed=14)
    if "--dup-only" in sys.argv:
        return
    run_sampling_case(ref, "sampling_small_auto", small, dict(auto_adjust=True, traj_number=20, num_cores=1),
                      seed=15, np_seed=9)
    two = synth.branching_graph(n=900, k=16, n_branches=3, seed=4, n_root_stems=2, workers=1, scale=4.0, noise=0.03)
    run_sampling_case(ref, "sampling_two_stems_normalize", two,
                      dict(auto_adjust=False, max_steps=75, traj_number=40, sim_number=400, normalize=True,
                           root_clusters=two.root_clusters, end_clusters=two.end_clusters, num_cores=1),
                      seed=16, np_seed=10, row_scale_seed=21)
    # tiny graph, short walks, every cell terminal (the undirected_simulations call shape, utils.py:86-90):
    # many duplicate sequences, so the first-occurrence de-duplication (:329-333) matters
    tiny = synth.branching_graph(n=60, k=4, n_branches=2, seed=5, workers=1, scale=2.0, noise=0.05)
    tiny.clusters[:] = "1"
    run_sampling_case(ref, "sampling_tiny_dups", tiny,
                      dict(auto_adjust=False, max_steps=5, traj_number=300, sim_number=300, min_clusters=1,
                           root_cells=[3, 7, 11, 20], end_points=np.arange(60).tolist(), num_cores=1),
                      seed=17, np_seed=11, cluster_key="undirected_sim_included")
    if "--small-only" not in sys.argv:
        c1 = synth.branching_graph(n=3696, k=30, n_branches=4, seed=0, workers=1)
        run_sampling_case(ref, "sampling_config1_auto", c1, dict(auto_adjust=True, sim_number=2000, num_cores=1),
                          seed=1234, np_seed=0, matrix_dtype=np.float32)


if __name__ == "__main__":
    main()
