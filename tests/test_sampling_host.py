"""Host logic of cytopath_b200.sampling (a7, a3 host side, a4, a5) on CPU, with the oracle
standing in for the CUDA kernels (tests/oracle_backend.py).  The GPU suite repeats the
golden cases through the real library (tests/test_gpu_parity.py)."""
import contextlib
import io
import os
import subprocess
import sys
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

import cytopath_b200
from cytopath_b200 import synth
from cytopath_b200.sampler import _canonical_csr, check_convergence_criteria
from tests import oracle_backend
from tests.helpers import SAMPLING_CASES, adata_from_case, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def check_against_golden(ad, z, score_rtol=2e-6):
    s = ad.uns["samples"]
    ri = ad.uns["run_info"]
    assert ri["simulation_steps"] == int(z["simulation_steps"])
    assert ri["didrun"] is True
    np.testing.assert_array_equal(np.asarray(ri["root_cells"]), z["run_root_cells"])
    np.testing.assert_array_equal(np.asarray(ri["end_points"]), z["run_end_points"])
    assert s["cell_sequences"].dtype == np.int32
    np.testing.assert_array_equal(s["cell_sequences"], z["cell_sequences"])
    assert s["cluster_sequences"].dtype == z["cluster_sequences"].dtype
    np.testing.assert_array_equal(s["cluster_sequences"], z["cluster_sequences"])
    assert s["transition_score"].dtype == np.float32
    np.testing.assert_allclose(s["transition_score"], z["transition_score"], rtol=score_rtol)
    assert len(ri["b200"]["rounds"]) == int(z["rounds"])


@pytest.mark.parametrize("name", SAMPLING_CASES)
def test_sampling_host_logic_matches_reference(name):
    z = load(name)
    ad, kw = adata_from_case(z)
    np.random.seed(int(z["np_seed"]))
    run_quiet(cytopath_b200.sampling, ad, **kw, seed=int(z["seed"]), _backend=oracle_backend, dist=False)
    check_against_golden(ad, z)


def test_copy_returns_new_object_and_leaves_input_alone():
    z = load("sampling_tiny_dups")
    ad, kw = adata_from_case(z)
    np.random.seed(int(z["np_seed"]))
    out = run_quiet(cytopath_b200.sampling, ad, **kw, seed=int(z["seed"]), copy=True, _backend=oracle_backend, dist=False)
    assert out is not ad and "samples" not in ad.uns
    check_against_golden(out, z)


def test_signature_is_the_reference_signature():
    import inspect
    sig = inspect.signature(cytopath_b200.sampling)
    names = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert names == ["data", "auto_adjust", "matrix_key", "cluster_key", "max_steps", "min_sim_ratio", "rounds_limit",
                     "traj_number", "sim_number", "end_point_probability", "root_cell_probability", "end_points",
                     "root_cells", "end_clusters", "root_clusters", "min_clusters", "max_iter", "tol", "normalize",
                     "unique", "num_cores", "copy"]
    d = {p.name: p.default for p in sig.parameters.values()}
    assert (d["auto_adjust"], d["matrix_key"], d["cluster_key"], d["max_steps"], d["min_sim_ratio"]) == (True, "T_forward", "louvain", 10, 0.6)
    assert (d["rounds_limit"], d["traj_number"], d["sim_number"], d["min_clusters"], d["max_iter"], d["tol"]) == (10, 200, 500, 2, 200, 1e-3)
    assert (d["end_point_probability"], d["root_cell_probability"], d["normalize"], d["unique"], d["num_cores"], d["copy"]) == (0.99, 0.99, False, True, 1, False)


# ---- error behaviour (reference :156-262, :367) --------------------------------------------
def small_adata():
    g = synth.branching_graph(n=300, k=10, n_branches=2, seed=3, workers=1, scale=4.0, noise=0.03)
    return g, g.to_adata()


def test_errors_match_reference_messages():
    g, ad = small_adata()
    kw = dict(_backend=oracle_backend, dist=False, auto_adjust=False)
    bad = ad.copy(); bad.uns["T_forward"] = bad.uns["T_forward"][:100]
    with pytest.raises(ValueError, match="num_cells x num_cells"):
        run_quiet(cytopath_b200.sampling, bad, **kw)
    bad = ad.copy(); bad.uns["T_forward"] = bad.uns["T_forward"] * 1.5
    with pytest.raises(ValueError, match="has not been normalized"):
        run_quiet(cytopath_b200.sampling, bad, **kw)
    bad = ad.copy(); bad.obs["louvain"] = bad.obs["louvain"].astype(str)
    with pytest.raises(ValueError, match="Categorical cluster labels"):
        run_quiet(cytopath_b200.sampling, bad, **kw)
    with pytest.raises(ValueError, match="Categorical cluster labels"):
        run_quiet(cytopath_b200.sampling, ad.copy(), cluster_key="nope", **kw)
    bad = ad.copy(); del bad.obs["root_cells"]
    with pytest.raises(ValueError, match="Root cells must be provided"):
        run_quiet(cytopath_b200.sampling, bad, **kw)
    bad = ad.copy(); del bad.obs["end_points"]
    with pytest.raises(ValueError, match="End points must be provided"):
        run_quiet(cytopath_b200.sampling, bad, root_clusters=g.root_clusters, **kw)
    with pytest.raises(ValueError, match="Terminal cluster set is empty"):
        run_quiet(cytopath_b200.sampling, ad.copy(), end_point_probability=2.0, **kw)
    # unreachable quota -> the reference's failure rule fires at round rounds_limit + 1 (:360-367)
    with pytest.raises(ValueError, match="Sampling failed"):
        run_quiet(cytopath_b200.sampling, ad.copy(), max_steps=3, rounds_limit=1, traj_number=50, sim_number=40,
                  root_clusters=g.root_clusters, end_clusters=g.end_clusters, **kw)


def test_non_louvain_key_warns():
    g, ad = small_adata()
    ad.obs["leiden"] = ad.obs["louvain"]
    with pytest.warns(UserWarning, match="finely resolved clustering"):
        with contextlib.redirect_stdout(io.StringIO()):
            cytopath_b200.sampling(ad, auto_adjust=False, cluster_key="leiden", max_steps=60, traj_number=5, sim_number=200,
                                   _backend=oracle_backend, dist=False)


def test_numpy_array_arguments_like_undirected_simulations():
    """utils.undirected_simulations passes numpy arrays for root_cells / end_points (utils.py:86-90)."""
    g, ad = small_adata()
    ad.obs["undirected_sim_included"] = pd.Categorical(["1"] * ad.shape[0])
    run_quiet(cytopath_b200.sampling, ad, root_cells=np.arange(0, 300, 7), end_points=np.arange(300),
              cluster_key="undirected_sim_included", auto_adjust=False, min_clusters=1, sim_number=300, traj_number=300,
              max_steps=6, _backend=oracle_backend, dist=False)
    assert ad.uns["samples"]["cell_sequences"].shape == (300, 6)


UNDIRECTED_CASES = ["undirected_all_cells", "undirected_subset"]


def check_undirected(name, **extra):
    """cytopath_b200.undirected_simulations against the reference's utils.undirected_simulations (utils.py:42-101)."""
    import json
    z = load(name)
    ad, _ = adata_from_case(dict(z, kwargs=np.asarray(json.dumps({"cluster_key": "louvain"}))))
    kw = json.loads(str(z["kwargs"]))
    np.random.seed(int(z["np_seed"]))
    run_quiet(cytopath_b200.undirected_simulations, ad, **kw, seed=int(z["seed"]), **extra)
    s = ad.uns["undirected_samples"]
    np.testing.assert_array_equal(s["cell_sequences"], z["cell_sequences"])
    np.testing.assert_array_equal(s["cluster_sequences"], z["cluster_sequences"])
    np.testing.assert_allclose(s["transition_score"], z["transition_score"], rtol=2e-6)
    np.testing.assert_array_equal(np.asarray(ad.obs["log_terminal_state_freq"], dtype=np.float64), z["log_terminal_state_freq"])
    assert "samples" not in ad.uns and "undirected_sim_included" not in ad.obs      # the reference works on a copy (utils.py:73)


@pytest.mark.parametrize("name", UNDIRECTED_CASES)
def test_undirected_simulations_host_logic_matches_reference(name):
    check_undirected(name, _backend=oracle_backend, dist=False)


def test_duplicate_and_unsorted_end_points_follow_the_reference_scan_order():
    """The reference scans `end_points` in list order and counts a cell listed twice twice (:343-354, :386-394)."""
    from oracle import sampler_port as sp
    g, ad = small_adata()
    ends = np.where(np.isin(g.clusters, g.end_clusters))[0]
    ends = np.concatenate([ends[::-1][:25], ends[:10], ends[::-1][:5]]).tolist()        # unsorted, with repeats
    kw = dict(auto_adjust=False, max_steps=60, traj_number=40, sim_number=300, root_clusters=g.root_clusters, end_points=ends)
    a, b = ad.copy(), ad.copy()
    np.random.seed(3)
    run_quiet(cytopath_b200.sampling, a, **kw, seed=5, _backend=oracle_backend, dist=False)
    np.random.seed(3)
    run_quiet(sp.sampling_port, b, **kw, seed=5)
    np.testing.assert_array_equal(a.uns["samples"]["cell_sequences"], b.uns["samples"]["cell_sequences"])
    np.testing.assert_array_equal(a.uns["samples"]["cluster_sequences"], b.uns["samples"]["cluster_sequences"])


def test_canonical_csr_equals_densify_resparsify():
    M = synth.random_sparse_rows(n=150, max_nnz=9, seed=11, messy=True, dtype=np.float32)
    want = sparse.csr_matrix(np.float64(np.asarray(M.toarray())))
    got = _canonical_csr(M)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    dense = _canonical_csr(np.asarray(M.toarray()))
    np.testing.assert_array_equal(dense.data, want.data)


def test_messy_csr_is_sampled_in_its_canonical_form():
    """A float CSR goes to the graph build as it is; when K1 reports unsorted / duplicate columns or explicit zeros the
    host canonicalises it like the reference's densify + re-sparsify (:168-170, :14) and rebuilds: same result as
    handing over the canonical matrix in the first place."""
    import pandas as pd
    from cytopath_b200.synth import AnnDataShim
    M = synth.random_sparse_rows(n=150, max_nnz=9, seed=11, messy=True)
    assert not (M.has_canonical_format and np.count_nonzero(M.data) == M.nnz)
    clusters = np.array(["c%d" % (i % 3) for i in range(150)])
    out = []
    for T in (M, _canonical_csr(M)):
        obs = pd.DataFrame({"louvain": pd.Categorical(clusters)})
        ad = AnnDataShim(150, obs, {"T_forward": T})
        np.random.seed(2)
        run_quiet(cytopath_b200.sampling, ad, auto_adjust=False, max_steps=8, traj_number=20, sim_number=200, root_cells=[0, 5, 9],
                  end_points=list(range(100, 150)), min_clusters=1, seed=3, _backend=oracle_backend, dist=False)
        out.append(ad)
    np.testing.assert_array_equal(out[0].uns["samples"]["cell_sequences"], out[1].uns["samples"]["cell_sequences"])
    assert out[0].uns["run_info"]["b200"]["timing"]["graph"].get("rebuilt_from_canonical_form") is True
    assert "rebuilt_from_canonical_form" not in out[1].uns["run_info"]["b200"]["timing"]["graph"]


def test_label_helpers_equal_the_per_cell_formulas():
    """_Labels works per category and expands through the codes; the results are those of the reference's per-cell
    expressions (`obs[key].astype(str)`, `isin`, `np.unique`, the U8 store of :39) -- with unused categories, labels
    longer than 8 characters that collide once cut, categories that share a str(), and missing values (fallback)."""
    from cytopath_b200.sampler import _Labels
    rng = np.random.default_rng(0)
    cats = pd.Index(["stem0", "a_very_long_label_one", "a_very_long_label_two", "b", 1, "1", "zz_unused"], dtype=object)
    codes = rng.integers(0, len(cats) - 1, size=5000)
    for with_nan in (False, True):
        c = codes.copy()
        if with_nan:
            c[::97] = -1
        col = pd.Series(pd.Categorical.from_codes(c, categories=cats))
        lab = _Labels(col)
        assert (lab.cats is None) == with_nan
        strs = np.array([str(x) for x in col.astype(str).values], dtype=str)
        for wanted in (["b", 1], ["1"], ["nope"], ["stem0", "a_very_long_label_two", "zz_unused"]):
            np.testing.assert_array_equal(lab.cells_in(col, wanted), np.where(np.asarray(col.isin(wanted)))[0])
        cells = rng.choice(5000, size=700, replace=True)
        names, local = lab.groups_of(cells)
        want_names = np.unique(strs[cells])
        np.testing.assert_array_equal(names, want_names)
        assert names.dtype == want_names.dtype
        np.testing.assert_array_equal(names[local], strs[cells])
        only_b = np.flatnonzero(strs == "b")[:5]
        np.testing.assert_array_equal(lab.groups_of(only_b)[0], np.array(["b"]))
        assert lab.groups_of(np.empty(0, dtype=np.int64))[0].size == 0
        lookup, code, n_codes = lab.truncated()
        cut = np.array([x[:8] for x in strs], dtype="U8")
        want_u, want_code = np.unique(cut, return_inverse=True)
        np.testing.assert_array_equal(lookup(cells), cut[cells])
        np.testing.assert_array_equal(lookup(cells.reshape(70, 10)), cut[cells.reshape(70, 10)])
        np.testing.assert_array_equal(code, want_code)
        assert code.dtype == np.int32 and n_codes == len(want_u)
        for width in ("U8", "U32"):
            got = lookup(cells.reshape(70, 10), width)
            assert got.dtype == np.dtype(width)
            np.testing.assert_array_equal(got, cut[cells.reshape(70, 10)].astype(width))
        big = rng.integers(0, 5000, size=(1500, 200))                 # 38 MB of '<U32': the threaded gather
        got = lookup(big, "U32")
        assert got.dtype == np.dtype("U32") and got.shape == big.shape
        np.testing.assert_array_equal(got, cut[big].astype("U32"))


def test_convergence_helpers():
    z = load("convergence_probes")
    for i, want in enumerate(z["result"]):
        got = check_convergence_criteria(z["p%d" % i], tol=1e-3)
        assert (-1 if got is None else got) == want


# ---- multi-rank sharding over gloo (world_size 2), CPU -----------------------------------
WORKER = r"""
import contextlib, io, os, sys, warnings
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
import cytopath_b200
from tests import oracle_backend
from tests.helpers import adata_from_case, load
dist.init_process_group("gloo")
z = load({case!r})
ad, kw = adata_from_case(z)
# only rank 0 has the golden run's numpy seed: the final draw (:395) must come from rank 0 on every rank
np.random.seed(int(z["np_seed"]) + 7919 * dist.get_rank())
with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cytopath_b200.sampling(ad, **kw, seed=int(z["seed"]), _backend=oracle_backend)
s = ad.uns["samples"]
np.savez({out!r} + ".%d.npz" % dist.get_rank(), cell_sequences=s["cell_sequences"], transition_score=s["transition_score"],
         cluster_sequences=s["cluster_sequences"], world=ad.uns["run_info"]["b200"]["world_size"])
dist.destroy_process_group()
"""


@pytest.mark.parametrize("case", ["sampling_tiny_dups", "sampling_small_clusters"])
def test_two_rank_gloo_run_equals_reference(case, tmp_path):
    script = tmp_path / "worker.py"
    out = str(tmp_path / "out")
    script.write_text(WORKER.format(root=ROOT, case=case, out=out))
    env = dict(os.environ, TQDM_DISABLE="1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    z = load(case)
    for rank in (0, 1):
        got = np.load(out + ".%d.npz" % rank)
        assert int(got["world"]) == 2
        np.testing.assert_array_equal(got["cell_sequences"], z["cell_sequences"])
        np.testing.assert_array_equal(got["cluster_sequences"], z["cluster_sequences"])
        np.testing.assert_allclose(got["transition_score"], z["transition_score"], rtol=2e-6)
