"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors of the real reference.  Integer outputs must be bit-exact; transition scores are
float32 and compared at rtol 2e-6 (numpy's SIMD log differs from log() in the last bit)."""
import contextlib
import io
import warnings

import numpy as np
import pytest

import cytopath_b200
from cytopath_b200 import _lib, synth
from oracle import cwalk, philox, sampler_port as sp
from tests.helpers import SAMPLING_CASES, adata_from_case, csr_from, load
from tests.test_sampling_host import check_against_golden, run_quiet

pytestmark = pytest.mark.gpu

MODES = [(_lib.CDF_REFERENCE, 0), (_lib.CDF_EXACT, 1)]


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    _lib.require_device(0)          # raises (fails the test) if the CUDA path cannot run


# ---- K1 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_k1_cdf_matches_reference_matrix_prep(tag):
    z = load("matrix_prep_messy_" + tag)
    C = sp.canonical_csr(csr_from(z))
    g = _lib.Graph(C, _lib.CDF_REFERENCE)
    idx, ell = sp.to_ell(C, g.cdf())
    np.testing.assert_array_equal(idx, z["idx"])
    np.testing.assert_array_equal(ell, z["cdf"])
    g.close()


@pytest.mark.parametrize("n,k", [(2000, 30), (500, 200), (300, 1)])
@pytest.mark.parametrize("mode,omode", MODES)
def test_k1_cdf_bit_exact_vs_oracle(n, k, mode, omode):
    if k == 1:
        C = sp.canonical_csr(synth.random_sparse_rows(n=n, max_nnz=1, seed=1, messy=False))
    else:
        C = sp.canonical_csr(synth.branching_graph(n=n, k=k, n_branches=3, seed=n + k, workers=1).T)
    g = _lib.Graph(C, mode)
    got = g.cdf()
    want = cwalk.build_cdf(C, omode)
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)
    info = g.info()
    assert info["n_cells"] == n and info["nnz"] == C.nnz and info["max_row_nnz"] == int(np.diff(C.indptr).max())
    g.close()


def test_k1_ragged_rows_with_staging_overflow():
    rng = np.random.default_rng(5)
    n = 700
    lens = rng.integers(1, 400, size=n)
    lens[10] = 1; lens[11] = 399
    rows = np.repeat(np.arange(n), lens)
    cols = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens])
    vals = rng.random(len(cols)) + 1e-3
    from scipy import sparse
    M = sparse.csr_matrix((vals, (rows, cols)), shape=(n, n))
    M = sparse.diags(1.0 / np.asarray(M.sum(axis=1)).ravel()) @ M
    C = sp.canonical_csr(M)
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        np.testing.assert_array_equal(g.cdf(), cwalk.build_cdf(C, omode))
        g.close()


def test_k1_reports_whether_the_csr_was_canonical():
    """cyp_graph_canonical: K1 sees every entry, so it tells for free whether every row's columns are strictly increasing
    and whether an explicit zero is stored -- in the staged path and in the direct path of long rows."""
    from scipy import sparse

    def variants(C):
        r = int(np.argmax(np.diff(C.indptr)))                 # a row with at least two entries
        a = int(C.indptr[r])
        assert C.indptr[r + 1] - a >= 2
        yield "canonical", C, True
        U = C.copy(); U.indices = U.indices.copy(); U.data = U.data.copy()
        U.indices[[a, a + 1]] = U.indices[[a + 1, a]]; U.data[[a, a + 1]] = U.data[[a + 1, a]]
        yield "unsorted", U, False
        D = C.copy(); D.indices = D.indices.copy()
        D.indices[a + 1] = D.indices[a]
        yield "duplicate", D, False
        Z = C.copy(); Z.data = Z.data.copy()
        Z.data[int(C.indptr[-1]) - 1] = 0.0
        yield "explicit zero", Z, False

    small = sp.canonical_csr(synth.random_sparse_rows(n=300, max_nnz=12, seed=4, messy=False))
    rng = np.random.default_rng(5)                            # rows of up to 400 entries: 128 of them overflow K1's staging buffer
    lens = rng.integers(1, 400, size=700)
    rows = np.repeat(np.arange(700), lens)
    cols = np.concatenate([np.sort(rng.choice(700, size=l, replace=False)) for l in lens])
    big = sp.canonical_csr(sparse.csr_matrix((rng.random(len(cols)) + 1e-3, (rows, cols)), shape=(700, 700)))
    for C in (small, big, small.astype(np.float32)):
        for name, M, want in variants(C):
            g = _lib.Graph(M, _lib.CDF_EXACT)
            assert g.canonical() == want, name
            g.close()


def test_sampling_canonicalises_a_messy_csr_only_when_k1_says_so():
    """Unsorted columns, duplicates and explicit zeros (the reference densifies and re-sparsifies, :168-170, :14): sampling()
    hands the CSR to the graph build as it is, K1 reports it, the host canonicalises and rebuilds -- same result as for the
    canonical matrix, which is built once."""
    import pandas as pd
    from cytopath_b200.sampler import _canonical_csr
    M = synth.random_sparse_rows(n=150, max_nnz=9, seed=11, messy=True)
    clusters = np.array(["c%d" % (i % 3) for i in range(150)])
    out = []
    for T in (M, _canonical_csr(M)):
        ad = synth.AnnDataShim(150, pd.DataFrame({"louvain": pd.Categorical(clusters)}), {"T_forward": T})
        np.random.seed(2)
        run_quiet(cytopath_b200.sampling, ad, auto_adjust=False, max_steps=8, traj_number=20, sim_number=200, root_cells=[0, 5, 9],
                  end_points=list(range(100, 150)), min_clusters=1, seed=3)
        out.append(ad)
    np.testing.assert_array_equal(out[0].uns["samples"]["cell_sequences"], out[1].uns["samples"]["cell_sequences"])
    assert out[0].uns["run_info"]["b200"]["timing"]["graph"].get("rebuilt_from_canonical_form") is True
    assert "rebuilt_from_canonical_form" not in out[1].uns["run_info"]["b200"]["timing"]["graph"]


@pytest.mark.parametrize("mode,omode", MODES)
def test_k1_float32_values_equal_the_widened_matrix(mode, omode):
    """scvelo stores T_forward as float32; the reference widens it with np.float64(toarray()) (:168).
    cyp_graph_create_f32 uploads the float32 values and widens on the device: same CDF bits, and the
    row-sum range the kernel reports equals the left-to-right float64 sums."""
    g32 = synth.branching_graph(n=1500, k=20, n_branches=3, seed=9, workers=1).T.astype(np.float32)
    C32 = sp.canonical_csr(g32).astype(np.float32)
    assert C32.dtype == np.float32
    C64 = C32.astype(np.float64)
    a, b = _lib.Graph(C32, mode), _lib.Graph(C64, mode)
    np.testing.assert_array_equal(a.cdf(), b.cdf())
    np.testing.assert_array_equal(a.cdf(), cwalk.build_cdf(C64, omode))
    sums = np.array([np.cumsum(C64.data[s:e])[-1] for s, e in zip(C64.indptr[:-1], C64.indptr[1:])])
    assert a.row_sum_range() == b.row_sum_range() == (sums.min(), sums.max())
    roots = np.arange(40)
    np.testing.assert_array_equal(a.walk(roots, 20, 30, seed=3), b.walk(roots, 20, 30, seed=3))
    a.close(); b.close()


def test_sampling_accepts_float32_matrix_in_place_and_checks_normalisation_on_device():
    z = load("sampling_small_clusters")
    ad, kw = adata_from_case(z)
    T32 = sp.canonical_csr(ad.uns[kw.get("matrix_key", "T_forward")]).astype(np.float32)
    T32.sort_indices()
    before = T32.data.copy()
    ad.uns[kw.get("matrix_key", "T_forward")] = T32
    ad2, _ = adata_from_case(z)
    ad2.uns[kw.get("matrix_key", "T_forward")] = T32.astype(np.float64)
    for a in (ad, ad2):
        np.random.seed(1)
        run_quiet(cytopath_b200.sampling, a, **kw, seed=5)
    np.testing.assert_array_equal(ad.uns["samples"]["cell_sequences"], ad2.uns["samples"]["cell_sequences"])
    np.testing.assert_array_equal(T32.data, before)                       # the caller's matrix is never modified
    bad = T32.copy()
    bad.data[bad.indptr[7]:bad.indptr[8]] *= 1.01
    ad.uns[kw.get("matrix_key", "T_forward")] = bad
    with pytest.raises(ValueError, match="has not been normalized"):
        run_quiet(cytopath_b200.sampling, ad, **kw, seed=5)
    kw2 = dict(kw, normalize=True)
    run_quiet(cytopath_b200.sampling, ad, **kw2, seed=5)


@pytest.mark.parametrize("gname", ["k4", "k12", "k30", "k50", "k200", "ragged", "hubs"])
@pytest.mark.parametrize("mode,omode", MODES)
def test_k1b_fused_rows_byte_exact_vs_oracle(gname, mode, omode, monkeypatch):
    """K1b (csrc/rows_v2.cu): header, 12-bit code planes, inlined targets of the most probable entries, row words
    and target words, byte for byte against the numpy restatement built from the oracle's CDF -- with the automatic
    choice of extra inline sectors and with 0 / 2 forced."""
    from oracle import fused_rows_port as frp
    C = sp.canonical_csr(GRAPHS[gname]())
    cdf = cwalk.build_cdf(C, omode)
    for extra in (None, 0, 2):
        if extra is None:
            monkeypatch.delenv("CYP_V2_EXTRA", raising=False)
        else:
            monkeypatch.setenv("CYP_V2_EXTRA", str(extra))
        g = _lib.Graph(C, mode)
        lay = g.layout()
        assert lay["fused"]
        rows, words, targets = g.fused()
        want_rows, want_words, want_targets = frp.build(C, cdf, exact=(mode == _lib.CDF_EXACT), extra=extra)
        np.testing.assert_array_equal(words, want_words)
        np.testing.assert_array_equal(targets, want_targets)
        assert rows.shape == want_rows.shape and lay["max_block_bytes"] == int((want_words >> 27).max()) * 32
        np.testing.assert_array_equal(rows, want_rows)
        g.close()


def test_k1b_not_applicable_falls_back_to_edge_records():
    """A negative probability (CDF not monotone): no fused rows, no 16-bit codes; the literal-scan kernels take over
    and still match the oracle."""
    from scipy import sparse
    P = np.array([[0.5, -0.25, 0.75, 0.0], [0.0, 1.0, 0.0, 0.0], [0.25, 0.25, 0.25, 0.25], [1.0, 0, 0, 0]])
    C = sp.canonical_csr(sparse.csr_matrix(P))
    g = _lib.Graph(C, _lib.CDF_EXACT)
    assert not g.layout()["fused"] and "walk_v2" not in g.variant_name(0)
    roots = np.array([0, 1, 2, 3])
    want = cwalk.walk(C, cwalk.build_cdf(C, 1), np.repeat(roots, 30), 40, seed=2, strict=False)
    np.testing.assert_array_equal(g.walk(roots, 30, 40, seed=2, strict=False), want)
    g.close()


def test_k1b_very_long_row_is_searched_in_hbm(monkeypatch):
    """One row of 900 entries (its block needs 43 sectors, more than a row word can say): the word's sector field
    saturates, the step kernel takes the length from the header and searches the block in global memory."""
    from scipy import sparse
    from oracle import fused_rows_port as frp
    rng = np.random.default_rng(11)
    n = 900
    M = sparse.lil_matrix((n, n))
    M[0, :] = np.exp(3.0 * rng.normal(size=n))
    for i in range(1, n):
        c = rng.choice(n, size=5, replace=False)
        M[i, c] = rng.random(5) + 0.05
        M[i, 0] = 2.0                                                  # everybody keeps returning to the long row
    M = M.tocsr()
    M = sparse.diags(1.0 / np.asarray(M.sum(axis=1)).ravel()) @ M
    C = sp.canonical_csr(M)
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        lay = g.layout()
        assert lay["fused"] and lay["max_block_bytes"] > 31 * 32 and "walk_v2" in g.variant_name(0)
        cdf = cwalk.build_cdf(C, omode)
        rows, words, targets = g.fused()
        want_rows, want_words, want_targets = frp.build(C, cdf, exact=(mode == _lib.CDF_EXACT))
        np.testing.assert_array_equal(words, want_words)
        np.testing.assert_array_equal(rows, want_rows)
        np.testing.assert_array_equal(targets, want_targets)
        roots = np.array([0, 1, 2, 3])
        want = cwalk.walk(C, cdf, np.repeat(roots, 200), 60, seed=2)
        assert (want == 0).mean() > 0.2                               # the long row is visited all the time
        np.testing.assert_array_equal(g.walk(roots, 200, 60, seed=2), want)
        g.close()


# ---- K4 ---------------------------------------------------------------------------------
def test_k4_power_iteration_matches_scipy():
    """eps history of iterate_state_probability (:62-87): the state vector is bit-equal to scipy's
    `p @ T` (same accumulation order), eps agrees to 1e-13 (the cosine's dot products are reduced in a
    different association than BLAS), and the derived max_steps is the same integer."""
    from scipy.spatial.distance import cosine
    from cytopath_b200.sampler import check_convergence_criteria
    for n, k, dtype in ((3696, 30, np.float32), (900, 16, np.float64)):
        gph = synth.branching_graph(n=n, k=k, n_branches=4, seed=0, workers=1)
        T = gph.T.astype(dtype)
        init = gph.root_prob / gph.root_prob.sum()
        stat = gph.end_prob / gph.end_prob.sum()
        want = np.empty(120)
        p = init
        for i in range(120):
            p = p @ T
            want[i] = cosine(p, stat)
        got, state = _lib.power_iteration(T, init, stat, 120, return_state=True)
        np.testing.assert_array_equal(state, p)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
        assert check_convergence_criteria(got, tol=1e-3) == check_convergence_criteria(want, tol=1e-3)


def test_k4_hub_columns_and_empty_columns():
    """SpMV-transpose with columns of thousands of entries next to empty ones (which produce 0): the state stays bit-equal
    to scipy's `p @ T`."""
    T = synth.hub_graph(n=3000, k_typ=(4, 30), hub_share=0.01, k_hub=(2000, 2600), seed=3).T.tocsr()
    T = sp.canonical_csr(T)
    keep = np.ones(3000, dtype=bool); keep[[5, 6, 1500, 2999]] = False              # four empty columns
    T = sp.canonical_csr(T @ __import__("scipy.sparse", fromlist=["diags"]).diags(keep.astype(np.float64)))
    assert int(np.diff(T.tocsc().indptr).max()) > 1500 and (np.diff(T.tocsc().indptr) == 0).sum() >= 4
    rng = np.random.default_rng(0)
    init = rng.random(3000); init /= init.sum()
    stat = rng.random(3000); stat /= stat.sum()
    p = init
    for _ in range(7):
        p = p @ T
    got, state = _lib.power_iteration(T, init, stat, 7, return_state=True)
    np.testing.assert_array_equal(state, p)
    assert state[5] == 0.0 and state[2999] == 0.0


def test_k4_iterate_state_probability_returns_the_reference_history():
    """The drop-in of iterate_state_probability (:62-87): state history bit-equal to the scipy loop, same `converged`;
    and K4 on the matrix a graph already holds (no second upload) gives the same eps history."""
    from scipy.spatial.distance import cosine
    from cytopath_b200.sampler import iterate_state_probability, check_convergence_criteria
    g = synth.branching_graph(n=300, k=10, n_branches=2, seed=5, workers=1)
    ad = g.to_adata()
    init = (ad.obs["root_cells"] / ad.obs["root_cells"].sum()).values
    stat = (ad.obs["end_points"] / ad.obs["end_points"].sum()).values
    h, full, c = run_quiet(iterate_state_probability, ad, init=init, stationary=stat, max_iter=50, tol=1e-3)
    assert full.shape == (50, 300) and h.shape[0] == c
    p, eps = init, np.empty(50)
    for i in range(50):
        np.testing.assert_array_equal(full[i], p)
        p = p @ ad.uns["T_forward"]
        eps[i] = cosine(p, stat)
    want = check_convergence_criteria(eps, tol=1e-3)
    assert c == (want if isinstance(want, int) and want < 50 else 50)
    C = sp.canonical_csr(g.T)
    graph = _lib.Graph(C, _lib.CDF_REFERENCE)
    np.testing.assert_allclose(graph.power_iteration(init, stat, 50), eps, rtol=0, atol=1e-13)
    np.testing.assert_array_equal(graph.power_iteration(init, stat, 50), _lib.power_iteration(C, init, stat, 50))
    graph.close()


# ---- K2 ---------------------------------------------------------------------------------
def test_k2_matches_reference_markov_sim_golden():
    z = load("markov_sim_small")
    C = sp.canonical_csr(csr_from(z))
    g = _lib.Graph(C, _lib.CDF_REFERENCE)
    sim, steps, seed = int(z["sim"]), int(z["steps"]), int(z["seed"])
    got = g.walk(z["roots"], sim, steps - 1, seed=seed)
    np.testing.assert_array_equal(got, z["seq"])                      # native Philox == the stream fed to the reference
    u = philox.walker_uniforms(seed, 0, 0, 2 * sim, steps - 1)
    np.testing.assert_array_equal(g.walk(z["roots"], sim, steps - 1, uniforms=u), z["seq"])    # injected uniforms
    np.testing.assert_allclose(g.transition_scores(z["seq"]), z["score"], rtol=2e-6)
    g.close()


GRAPHS = {
    "k4": lambda: synth.branching_graph(n=500, k=4, n_branches=2, seed=1, workers=1, scale=3.0, noise=0.03).T,
    "k12": lambda: synth.branching_graph(n=1500, k=12, n_branches=3, seed=2, workers=1, scale=4.0).T,
    "k30": lambda: synth.branching_graph(n=3696, k=30, n_branches=4, seed=0, workers=1).T,
    "k50": lambda: synth.branching_graph(n=4000, k=50, n_branches=4, seed=3, workers=1).T,
    "k200": lambda: synth.cycling_graph(n=3000, k=200, seed=4, workers=1).T,
    "ragged": lambda: synth.random_sparse_rows(n=400, max_nnz=70, seed=9, messy=True),
    "hubs": lambda: synth.hub_graph(n=3000, seed=6),        # 2 % of the rows are longer than the fused-row kernel's slot
}


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("mode,omode", MODES)
def test_k2_bit_exact_vs_oracle_all_variants(gname, mode, omode, monkeypatch):
    C = sp.canonical_csr(GRAPHS[gname]())
    g = _lib.Graph(C, mode)
    cdf = cwalk.build_cdf(C, omode)
    rng = np.random.default_rng(7)
    roots = rng.choice(C.shape[0], size=13, replace=False)
    sims, nt, seed, rnd = 37, 61, 0xDEADBEEFCAFE, 2
    want = cwalk.walk(C, cdf, np.repeat(roots, sims), nt, seed=seed, round_idx=rnd, strict=False)
    L = _lib.load()
    import ctypes
    import torch
    d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
    for lanes in (1, 2, 4, 8, 16, 32):
        for unroll in (4,):
            out = torch.full((len(roots) * sims, nt + 1), -7, dtype=torch.int32, device="cuda")
            miss = torch.zeros(1, dtype=torch.int64, device="cuda")
            _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, len(roots) * sims, nt, seed, rnd,
                                         None, out.data_ptr(), miss.data_ptr(), lanes * 100 + unroll,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg="lanes=%d unroll=%d" % (lanes, unroll))
    # TMA flavour: 128 / 256 / 512 / 1024 threads per CTA (1024 forces rows to stream through small slots)
    for variant in (9004, 9008, 9016, 9032, 9204, 9216, 9232, 9304, 9316, 9332):     # 90xx: CDF rows as stored; 92xx: 16-bit CDF codes
        for n_t in (nt, 62):                                  # 63 columns: scalar stores; 64 columns: int4 stores
            wnt = want if n_t == nt else cwalk.walk(C, cdf, np.repeat(roots, sims), n_t, seed=seed, round_idx=rnd, strict=False)
            out = torch.full((len(roots) * sims, n_t + 1), -7, dtype=torch.int32, device="cuda")
            miss = torch.zeros(1, dtype=torch.int64, device="cuda")
            _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, len(roots) * sims, n_t, seed, rnd,
                                         None, out.data_ptr(), miss.data_ptr(), variant,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            np.testing.assert_array_equal(out.cpu().numpy(), wnt, err_msg="%s nt=%d" % (g.variant_name(variant), n_t))
    # fused-row flavour (12-bit codes, inlined targets): 32 / 128 / 512 / 1024 threads per CTA, one (95xx) or two (97xx) CTAs
    # per SM; L1-allocating and L1-bypassing row copies; 62, 63 and 64 columns
    for variant, l1 in ((9501, "1"), (9504, "0"), (9516, "1"), (9532, "0"), (9701, "0"), (9708, "0"), (9720, "0")):
        monkeypatch.setenv("CYP_V2_L1", l1)
        if gname == "k200":
            assert g.variant_name(variant) == "not applicable"       # blocks above 256 bytes: walk_tma.cu
            continue
        assert "walk_v2_kernel" in g.variant_name(variant), g.variant_name(variant)
        if variant > 9700:
            name = g.variant_name(variant)                            # two CTAs per SM: uniform 128-byte blocks only
            assert ("x 2 CTAs/SM" in name) == ("slot=128B" in name and "uniform rows" in name), name
        for n_t in (nt, 62, 63):
            wnt = want if n_t == nt else cwalk.walk(C, cdf, np.repeat(roots, sims), n_t, seed=seed, round_idx=rnd, strict=False)
            out = torch.full((len(roots) * sims, n_t + 1), -7, dtype=torch.int32, device="cuda")
            miss = torch.zeros(1, dtype=torch.int64, device="cuda")
            _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, len(roots) * sims, n_t, seed, rnd,
                                         None, out.data_ptr(), miss.data_ptr(), variant,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            np.testing.assert_array_equal(out.cpu().numpy(), wnt, err_msg="%s nt=%d" % (g.variant_name(variant), n_t))
    monkeypatch.delenv("CYP_V2_L1")
    assert ("walk_v2_kernel" in g.variant_name(0)) == (gname != "k200")     # the auto variant picks it, except for long rows
    if gname == "hubs":                                       # slot for the typical row; the hubs are searched in HBM
        assert "slot=128B" in g.variant_name(0) and g.layout()["max_block_bytes"] > 256
    # auto variant through the host entry point, a shard in the middle of the round
    got = g.walk(roots, sims, nt, seed=seed, round_idx=rnd, walker_begin=100, n_walkers=211, strict=False)
    np.testing.assert_array_equal(got, want[100:311])
    g.close()


def test_k2_edge_cases_ties_zero_and_last_bin():
    """u == cdf[j] picks bin j (`<=` at :49); u = 0 picks bin 0; u just below 1 picks the last bin."""
    from scipy import sparse
    P = np.array([[0.25, 0.25, 0.5, 0.0], [0.0, 1.0, 0.0, 0.0], [0.125, 0.125, 0.25, 0.5], [1.0, 0, 0, 0]])
    C = sp.canonical_csr(sparse.csr_matrix(P))
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        cdf = cwalk.build_cdf(C, omode)
        u = np.array([[0.0, 0.25, 0.5], [0.25, 0.0, 0.999999], [0.5, 0.125, 0.75], [np.nextafter(0.5, 1), 0.5, 1.0 - 2**-53]])
        roots = np.array([0, 0, 2, 2])
        want = sp.walk(C, cdf, roots, u)
        got = g.walk(roots, 1, 3, uniforms=u)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(cwalk.walk(C, cdf, roots, 3, uniforms=u), want)
        g.close()


def test_k2_negative_probabilities_use_the_literal_scan():
    """A row whose CDF is not monotone (negative entry): the reference takes the FIRST index with
    u <= cdf (np.where(...)[0][0], :49); a binary search would be wrong, so the library must scan."""
    from scipy import sparse
    P = np.array([[0.7, -0.2, 0.5, 0.0], [0.0, 1.0, 0.0, 0.0], [0.3, 0.3, -0.1, 0.5], [0.25, 0.25, 0.25, 0.25]])
    C = sp.canonical_csr(sparse.csr_matrix(P))
    L = _lib.load()
    import ctypes
    import torch
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        cdf = cwalk.build_cdf(C, omode)
        roots = np.array([0, 2, 3])
        want = cwalk.walk(C, cdf, np.repeat(roots, 500), 20, seed=3)
        np.testing.assert_array_equal(g.walk(roots, 500, 20, seed=3), want)
        d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
        out = torch.zeros((1500, 21), dtype=torch.int32, device="cuda")
        _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), 3, 500, 0, 1500, 20, 3, 0, None, out.data_ptr(), None, 9008,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert "scan" in g.variant_name(9008)
        np.testing.assert_array_equal(out.cpu().numpy(), want)
        g.close()


def test_k2_under_normalised_row_raises_like_the_reference():
    from scipy import sparse
    P = np.array([[0.4995, 0.4995], [0.5, 0.5]])          # row 0 sums to 0.999: passes the :184-185 check
    C = sp.canonical_csr(sparse.csr_matrix(P))
    g = _lib.Graph(C, _lib.CDF_REFERENCE)
    u = np.array([[0.9995, 0.1]])
    with pytest.raises(IndexError):
        g.walk(np.array([0]), 1, 2, uniforms=u)
    with pytest.raises(IndexError):
        cwalk.walk(C, cwalk.build_cdf(C, 0), np.array([0]), 2, uniforms=u)
    g.close()


def test_k2_q16_ambiguous_codes_are_resolved_in_float64():
    """Rows whose CDF values crowd into single 1/65536 (and 1/4096) buckets, and uniforms aimed exactly at and
    next to those values: the 16-bit and 12-bit code flavours must still return what the float64 comparison returns."""
    from scipy import sparse
    rng = np.random.default_rng(3)
    n, k = 64, 40
    rows, cols, vals = [], [], []
    for i in range(n):
        c = np.sort(rng.choice(n, size=k, replace=False))
        v = np.full(k, 2.0 ** -20)                       # 16 consecutive entries share one 16-bit code
        v[rng.choice(k, size=3, replace=False)] = rng.random(3) + 0.1
        v = v / v.sum()
        rows += [i] * k; cols += list(c); vals += list(v)
    C = sp.canonical_csr(sparse.csr_matrix((vals, (rows, cols)), shape=(n, n)))
    import ctypes
    import torch
    L = _lib.load()
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        assert "q12" in g.variant_name(0) and "q16" in g.variant_name(9204)
        cdf = cwalk.build_cdf(C, omode)
        cdf64 = cdf.astype(np.float64)
        roots = np.arange(n)
        nt = 24
        # uniforms: exact CDF values of row 0..n-1 (ties), their neighbours, and plain random numbers
        u = rng.random((n * 8, nt))
        pool = np.concatenate([cdf64, np.nextafter(cdf64, 0), np.nextafter(cdf64, 2)])
        pool = pool[(pool >= 0) & (pool < 1)]
        mask = rng.random(u.shape) < 0.7
        u[mask] = rng.choice(pool, size=int(mask.sum()))
        want = cwalk.walk(C, cdf, np.repeat(roots, 8), nt, uniforms=u, strict=False)
        d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
        d_u = torch.tensor(u, dtype=torch.float64, device="cuda")
        for variant in (0, 9504, 9204, 9304, 9008, 404):
            out = torch.zeros((n * 8, nt + 1), dtype=torch.int32, device="cuda")
            miss = torch.zeros(1, dtype=torch.int64, device="cuda")
            _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), n, 8, 0, n * 8, nt, 0, 0, d_u.data_ptr(), out.data_ptr(),
                                         miss.data_ptr(), variant, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg=g.variant_name(variant))
        g.close()


def test_k2_single_step_and_zero_transitions():
    C = sp.canonical_csr(GRAPHS["k12"]())
    g = _lib.Graph(C, _lib.CDF_EXACT)
    roots = np.array([3, 9, 11])
    np.testing.assert_array_equal(g.walk(roots, 2, 0), np.repeat(roots, 2)[:, None])
    want = cwalk.walk(C, cwalk.build_cdf(C, 1), np.repeat(roots, 2), 1, seed=5)
    np.testing.assert_array_equal(g.walk(roots, 2, 1, seed=5), want)
    g.close()


# ---- K3 + sampling() end to end --------------------------------------------------------
@pytest.mark.parametrize("name", SAMPLING_CASES)
def test_sampling_cuda_matches_reference_golden(name):
    z = load(name)
    ad, kw = adata_from_case(z)
    np.random.seed(int(z["np_seed"]))
    run_quiet(cytopath_b200.sampling, ad, **kw, seed=int(z["seed"]))
    check_against_golden(ad, z)
    assert "walk_" in ad.uns["run_info"]["b200"]["kernel"]


@pytest.mark.parametrize("name", ["undirected_all_cells", "undirected_subset"])
def test_undirected_simulations_cuda_matches_reference(name):
    from tests.test_sampling_host import check_undirected
    check_undirected(name)


def test_f1_fused_mode_needs_no_sequences():
    """undirected_simulations(store_samples=False): log_terminal_state_freq equals the reference's (utils.py:92-95)
    without a single sequence being produced; also at 300k cells with every cell a root (K2 runs without a
    trajectory buffer, the sessions keep 16-byte records)."""
    import json
    for name in ("undirected_all_cells", "undirected_subset"):
        z = load(name)
        ad, _ = adata_from_case(dict(z, kwargs=np.asarray(json.dumps({"cluster_key": "louvain"}))))
        kw = json.loads(str(z["kwargs"]))
        np.random.seed(int(z["np_seed"]))
        run_quiet(cytopath_b200.undirected_simulations, ad, **kw, seed=int(z["seed"]), store_samples=False)
        np.testing.assert_array_equal(np.asarray(ad.obs["log_terminal_state_freq"], dtype=np.float64), z["log_terminal_state_freq"])
        assert "undirected_samples" not in ad.uns
    gph = synth.branching_graph(n=300_000, k=20, n_branches=6, seed=2)
    ad = gph.to_adata()
    np.random.seed(1)
    run_quiet(cytopath_b200.undirected_simulations, ad, max_steps=60, seed=3, store_samples=False)
    f = np.asarray(ad.obs["log_terminal_state_freq"], dtype=np.float64)
    visited = np.isfinite(f)
    assert 1000 < visited.sum() < 300_000 and abs(np.expm1(f[visited]).sum() - 1.0) < 1e-9      # frequencies of 300k selected sequences


def test_k2_randomised_shapes_vs_oracle():
    """40 random (graph, mode, roots, walkers, steps, seed, round, shard) combinations through the default kernel."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(20, 600))
        max_nnz = int(rng.integers(1, 90))
        C = sp.canonical_csr(synth.random_sparse_rows(n=n, max_nnz=min(max_nnz, n - 2), seed=trial, messy=bool(trial % 2)))
        mode = int(rng.integers(2))
        g = _lib.Graph(C, mode)
        cdf = cwalk.build_cdf(C, mode)
        roots = rng.choice(n, size=int(rng.integers(1, 9)), replace=True)
        sims, nt = int(rng.integers(1, 70)), int(rng.integers(0, 45))
        seed, rnd = int(rng.integers(2**62)), int(rng.integers(100))
        total = len(roots) * sims
        a = int(rng.integers(0, total)); b = int(rng.integers(a, total + 1))
        want = cwalk.walk(C, cdf, np.repeat(roots, sims)[a:b], nt, seed=seed, round_idx=rnd, walker_begin=a, strict=False)
        got = g.walk(roots, sims, nt, seed=seed, round_idx=rnd, walker_begin=a, n_walkers=b - a, strict=False)
        np.testing.assert_array_equal(got, want, err_msg="trial %d n=%d nnz<=%d mode=%d" % (trial, n, max_nnz, mode))
        g.close()


def make_session(C, g, roots, code, min_clusters, end_cells, max_steps, unique, seed):
    n = C.shape[0]
    slot = np.full(n, -1, dtype=np.int32)
    slot[end_cells] = np.arange(len(end_cells), dtype=np.int32)
    return _lib.Session(g, roots, code, int(code.max()) + 1, min_clusters, slot, len(end_cells), max_steps, unique, seed), slot


def oracle_round(C, cdf, roots, sims, max_steps, seed, rnd, code, min_clusters, slot, unique, w0=0, w1=None):
    starts = np.repeat(roots, sims)
    w1 = len(starts) if w1 is None else w1
    seq = cwalk.walk(C, cdf, starts[w0:w1], max_steps - 1, seed=seed, round_idx=rnd, walker_begin=w0)
    cs = np.sort(code[seq], axis=1)
    ok = (1 + (np.diff(cs, axis=1) != 0).sum(axis=1)) >= min_clusters
    cand = np.flatnonzero(ok & (slot[seq[:, -1]] >= 0))
    if unique and cand.size:
        cand = cand[np.sort(np.unique(seq[cand], axis=0, return_index=True)[1])]
    return seq, ok, cand


@pytest.mark.parametrize("chunk", [0, 97])
@pytest.mark.parametrize("hash_bits", [128, 6, 0])
@pytest.mark.parametrize("n_codes", [1, 40, 200, 700])
def test_k3_round_filters_and_dedup_vs_oracle(hash_bits, n_codes, chunk):
    """Short walks on a tiny graph produce many duplicate sequences; truncating the hash to a
    few bits forces collisions so the verify / re-hash path is exercised too."""
    C = sp.canonical_csr(synth.branching_graph(n=80, k=4, n_branches=2, seed=5, workers=1, scale=2.0, noise=0.05).T)
    n = C.shape[0]
    rng = np.random.default_rng(n_codes)
    code = rng.integers(n_codes, size=n).astype(np.int32)
    code[:n_codes] = np.arange(min(n_codes, n))[:n]
    code = np.unique(code, return_inverse=True)[1].astype(np.int32)
    roots = np.array([3, 7, 11, 20, 33])
    end_cells = np.sort(rng.choice(n, size=50, replace=False))
    min_clusters = 1 if n_codes == 1 else 3
    g = _lib.Graph(C, _lib.CDF_REFERENCE)
    cdf = cwalk.build_cdf(C, 0)
    s, slot = make_session(C, g, roots, code, min_clusters, end_cells, 6, True, 77)
    s.set_hash_bits(hash_bits)
    s.set_chunk_walkers(chunk)              # 97 walkers per chunk: duplicates must also be found across chunks
    total = 0
    for rnd, sims in enumerate((50, 400)):
        st = s.round(rnd, sims, 0, len(roots) * sims)
        seq, ok, cand = oracle_round(C, cdf, roots, sims, 6, 77, rnd, code, min_clusters, slot, True)
        assert st["n_walkers"] == len(seq) and st["n_pass_clusters"] == int(ok.sum())
        assert st["n_terminal"] == int((ok & (slot[seq[:, -1]] >= 0)).sum())
        assert st["n_kept"] == len(cand), (st, len(cand))
        if hash_bits < 128 and st["n_terminal"] > st["n_kept"] + 5:
            assert st["dedup_passes"] >= 2
        assert st["n_chunks"] == (1 if chunk == 0 else -(-len(seq) // chunk))
        r, w, sl = s.fetch_index(total)
        np.testing.assert_array_equal(w, cand)
        np.testing.assert_array_equal(sl, slot[seq[cand, -1]])
        assert (r == rnd).all()
        np.testing.assert_array_equal(s.fetch_rows(np.arange(total, total + len(cand))), seq[cand])
        total += len(cand)
    assert s.rows() == total
    pick = np.random.default_rng(1).integers(total, size=33)
    assert s.fetch_rows(pick).shape == (33, 6)
    s.close(); g.close()


class _DevPtr:
    """Zero-copy view of raw device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr, n_u64):
        self.__cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}


@pytest.mark.parametrize("hash_bits", [128, 5])
def test_k3_not_unique_keeps_duplicates_and_sharded_prune(hash_bits):
    """... and with the hash cut to 5 bits the cross-shard prune must tell colliding sequences apart by replaying them."""
    C = sp.canonical_csr(synth.branching_graph(n=80, k=4, n_branches=2, seed=5, workers=1, scale=2.0, noise=0.05).T)
    n = C.shape[0]
    code = np.zeros(n, dtype=np.int32)
    roots = np.array([3, 7, 11, 20])
    end_cells = np.arange(n)
    g = _lib.Graph(C, _lib.CDF_REFERENCE)
    cdf = cwalk.build_cdf(C, 0)
    s, slot = make_session(C, g, roots, code, 1, end_cells, 5, False, 9)
    st = s.round(0, 300, 0, 1200)
    assert st["n_kept"] == 1200 and st["dedup_passes"] == 0
    s.close()
    # two shards of one round, then global first-occurrence pruning == single-shard result
    seq, ok, cand = oracle_round(C, cdf, roots, 300, 5, 9, 0, code, 1, slot, True)
    import torch
    sa, _ = make_session(C, g, roots, code, 1, end_cells, 5, True, 9)
    sb, _ = make_session(C, g, roots, code, 1, end_cells, 5, True, 9)
    sa.set_hash_bits(hash_bits); sb.set_hash_bits(hash_bits)
    sa.round(0, 300, 0, 500)
    sb.round(0, 300, 500, 1200)
    recs = []
    for sess in (sa, sb):
        p, m = sess.round_records()
        recs.append(torch.as_tensor(_DevPtr(p, m * 3), device="cuda").clone())
    allr = torch.cat(recs)
    torch.cuda.synchronize()
    da = sa.round_prune(allr.data_ptr(), allr.numel() // 3)
    db = sb.round_prune(allr.data_ptr(), allr.numel() // 3)
    assert da == 0 and db > 0
    wa = sa.fetch_index()[1]; wb = sb.fetch_index()[1]
    np.testing.assert_array_equal(np.concatenate([wa, wb]), cand)
    np.testing.assert_array_equal(np.vstack([sa.fetch_rows(np.arange(sa.rows())), sb.fetch_rows(np.arange(sb.rows()))]), seq[cand])
    sa.close(); sb.close(); g.close()


# ---- size-independent properties at larger sizes ------------------------------------------
def test_partition_independence_and_edge_validity_large():
    """100k cells, 200k walkers: any sharding of the walker range gives the same rows, and
    every step follows an edge of T (checked against the CSR)."""
    gph = synth.branching_graph(n=100_000, k=30, n_branches=8, seed=0)
    C = sp.canonical_csr(gph.T)
    g = _lib.Graph(C, _lib.CDF_EXACT)
    roots = np.where(gph.clusters == "stem0")[0][:200]
    sims, nt = 1000, 50
    full = g.walk(roots, sims, nt, seed=1234)
    parts = [g.walk(roots, sims, nt, seed=1234, walker_begin=a, n_walkers=b - a) for a, b in ((0, 70_001), (70_001, 200_000))]
    np.testing.assert_array_equal(np.vstack(parts), full)
    src, dst = full[:, :-1].ravel(), full[:, 1:].ravel()
    assert (np.asarray(C[src, dst]).ravel() > 0).all()
    np.testing.assert_array_equal(full[:, 0], np.repeat(roots, sims))
    # the C oracle agrees on a 4k-walker sample of the same round
    pick = np.arange(0, 200_000, 50)
    cdf = cwalk.build_cdf(C, 1)
    for w in pick[:200]:
        np.testing.assert_array_equal(cwalk.walk(C, cdf, np.repeat(roots, sims)[w:w + 1], nt, seed=1234, walker_begin=int(w))[0], full[w])
    g.close()


def test_full_size_c4_properties():
    """BASELINE.json configs[3] at full size (1M cells, k=50, 1.25M walkers x 500 steps on one GPU): the default
    kernel and the rows-as-stored kernel produce identical trajectories (checksum over all 625M entries), a
    sub-range of the walkers reproduces its rows, every sampled step is an edge of T, and the C oracle agrees
    on a sample of walkers."""
    import ctypes
    import torch
    import bench
    w = bench.WORKLOADS["c4"]
    gph, C, roots = bench.make_graph(w, 1)
    W, nt = w["walkers"], w["max_steps"] - 1
    sims = -(-W // len(roots))
    L = _lib.load()
    g = _lib.Graph(C, _lib.CDF_EXACT)
    d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for variant in (0, 9016):
        out = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
        miss = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, 0, None, out.data_ptr(),
                                     miss.data_ptr(), variant, sptr))
        torch.cuda.synchronize()
        assert int(miss.item()) == 0
        outs.append(out)
    assert torch.equal(outs[0], outs[1])                                  # 16-bit-code kernel == float64-row kernel, all entries
    weights = torch.arange(1, nt + 2, device="cuda", dtype=torch.int64)
    checksum = (outs[0].to(torch.int64) * weights).sum(dim=1)             # per-walker checksum
    part = torch.empty((1000, nt + 1), dtype=torch.int32, device="cuda")
    _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 777_000, 1000, nt, 1234, 0, None, part.data_ptr(),
                                 None, 0, sptr))
    torch.cuda.synchronize()
    assert torch.equal((part.to(torch.int64) * weights).sum(dim=1), checksum[777_000:778_000])
    sample = outs[0][::2500].cpu().numpy()                                # 500 walkers
    src, dst = sample[:, :-1].ravel(), sample[:, 1:].ravel()
    assert (np.asarray(C[src, dst]).ravel() > 0).all()
    np.testing.assert_array_equal(sample[:, 0], roots[(np.arange(0, W, 2500)) // sims])
    cdf = cwalk.build_cdf(C, 1)
    for a, b in ((0, 60_000), (1_210_000, 1_250_000)):                    # 100k walkers x 499 transitions against the C oracle
        starts = roots[np.arange(a, b) // sims]
        want = cwalk.walk(C, cdf, starts, nt, seed=1234, walker_begin=a)
        np.testing.assert_array_equal(outs[0][a:b].cpu().numpy(), want)
    # the second large walk with the default kernel calibrates the placement of the fused rows (csrc/placement.cu: probes
    # on fresh allocations, possibly moving the arrays): same trajectories before and after
    assert g.placement()["probe_ms_before"] == 0.0
    again = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
    _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 1234, 0, None, again.data_ptr(), None, 0, sptr))
    torch.cuda.synchronize()
    assert g.placement()["probe_ms_before"] > 0.0
    assert torch.equal(again, outs[0])
    g.close()


@pytest.mark.parametrize("wl,mode,omode", [("c2", _lib.CDF_REFERENCE, 0), ("c3", _lib.CDF_EXACT, 1), ("c5", _lib.CDF_EXACT, 1)])
def test_bench_sizes_c2_c3_c5_sampled_vs_oracle(wl, mode, omode):
    """The other BASELINE.json configs at bench.py's sizes (5,780 cells x 710k walkers x 200 steps; 100k cells x 1M
    walkers x 200 steps; 250k cells k=200 x 1.25M walkers x 200 steps): the default kernel for each graph shape
    (fused rows with L1 copies / fused rows / 16-bit rows with bulk copies) against the C oracle on two blocks of
    10,000 walkers, plus the per-walker checksum of a re-walked sub-range."""
    import ctypes
    import torch
    import bench
    w = bench.WORKLOADS[wl]
    gph, C, roots = bench.make_graph(w, 1)
    W, nt = w["walkers"], w["max_steps"] - 1
    sims = -(-W // len(roots))
    L = _lib.load()
    g = _lib.Graph(C, mode)
    d_roots = torch.tensor(roots, dtype=torch.int32, device="cuda")
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.empty((W, nt + 1), dtype=torch.int32, device="cuda")
    miss = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, 0, W, nt, 4321, 3, None, out.data_ptr(), miss.data_ptr(), 0, sptr))
    torch.cuda.synchronize()
    assert int(miss.item()) == 0
    cdf = cwalk.build_cdf(C, omode)
    for a in (0, W - 10_000):
        starts = roots[np.arange(a, a + 10_000) // sims]
        want = cwalk.walk(C, cdf, starts, nt, seed=4321, round_idx=3, walker_begin=a)
        np.testing.assert_array_equal(out[a:a + 10_000].cpu().numpy(), want, err_msg=g.variant_name(0))
    weights = torch.arange(1, nt + 2, device="cuda", dtype=torch.int64)
    part = torch.empty((5000, nt + 1), dtype=torch.int32, device="cuda")
    _lib.check(L.cyp_walk_device(g.handle, d_roots.data_ptr(), len(roots), sims, W // 2, 5000, nt, 4321, 3, None, part.data_ptr(), None, 0, sptr))
    torch.cuda.synchronize()
    assert torch.equal((part.to(torch.int64) * weights).sum(dim=1), (out[W // 2:W // 2 + 5000].to(torch.int64) * weights).sum(dim=1))
    g.close()


def test_native_rng_one_step_frequencies_chi_square():
    """With the native Philox stream the empirical one-step transition frequencies match the
    effective probabilities (bin widths of the CDF actually sampled) at p > 1e-4 per row
    after Bonferroni correction over the tested rows."""
    from scipy import stats
    C = sp.canonical_csr(GRAPHS["k30"]())
    for mode, omode in MODES:
        g = _lib.Graph(C, mode)
        cdf = cwalk.build_cdf(C, omode).astype(np.float64)
        rows = np.array([0, 17, 500, 1234, 3000, 3695])
        sims = 200_000
        seq = g.walk(rows, sims, 1, seed=99)
        for i, r in enumerate(rows):
            a, b = C.indptr[r], C.indptr[r + 1]
            width = np.diff(np.concatenate([[0.0], cdf[a:b]]))
            width = np.maximum(width, 0.0)
            counts = np.array([(seq[i * sims:(i + 1) * sims, 1] == c).sum() for c in C.indices[a:b]])
            assert counts.sum() == sims
            keep = width * sims >= 5                                  # chi-square validity
            assert counts[~keep].sum() <= max(50, 10 * (width[~keep].sum() * sims))
            exp = width[keep] / width[keep].sum() * counts[keep].sum()
            chi2, p = stats.chisquare(counts[keep], exp)
            assert p > 1e-4 / len(rows), (mode, r, chi2, p)
        g.close()


# ---- f2: Hausdorff distance matrix ---------------------------------------------------------
@pytest.mark.parametrize("S,L,d", [(40, 37, 5), (12, 200, 50), (25, 129, 2), (3, 1, 7)])
@pytest.mark.parametrize("metric", ["euclidean", "manhattan", "chebyshev", "cosine"])
def test_f2_hausdorff_matrix_vs_oracle(S, L, d, metric):
    """distance_matrix (sample_clustering.py:57-71) against the numpy restatement; float64, rtol 1e-12
    (sums of d terms in a different association).  The `hausdorff` package itself is absent: parity unpinned."""
    from oracle import hausdorff_port as hp
    rng = np.random.default_rng(S * 1000 + L + d)
    base = rng.normal(size=(S, 1, d)) * 0.3
    coords = (base + np.cumsum(rng.normal(scale=0.05, size=(S, L, d)), axis=1) + 1.0).astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        got = cytopath_b200.distance_matrix(coords, distance=metric)
    want = hp.distance_matrix(coords, distance=metric)
    assert got.shape == (S, S) and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(got, got.T)
    assert (np.diag(got) == 0).all()


def test_f2_runs_on_sampler_output():
    """sampling() -> coordinate_assigner -> distance_matrix, the first three calls of cytopath.trajectories
    (trajectory_inference.py:57 -> sample_clustering.py:21-71), on a golden case."""
    from oracle import hausdorff_port as hp
    z = load("sampling_small_clusters")
    ad, kw = adata_from_case(z)
    rng = np.random.default_rng(0)
    ad.obsm["X_pca"] = rng.normal(size=(ad.shape[0], 10)).astype(np.float32)
    np.random.seed(int(z["np_seed"]))
    run_quiet(cytopath_b200.sampling, ad, **kw, seed=int(z["seed"]))
    seqs = ad.uns["samples"]["cell_sequences"][:60]
    coords = cytopath_b200.coordinate_assigner(ad, seqs, basis="pca")
    np.testing.assert_array_equal(coords, hp.coordinate_assigner(ad.obsm["X_pca"], seqs))
    D = run_quiet(cytopath_b200.distance_matrix, coords)
    np.testing.assert_allclose(D, hp.distance_matrix(coords), rtol=1e-12, atol=1e-14)


# ---- f4: directionality score ------------------------------------------------------------
@pytest.mark.parametrize("name", ["directionality_d2_f32", "directionality_d10_f64"])
def test_f4_directionality_score_matches_reference_golden(name):
    """cytopath_b200.directionality_score (csrc/directionality.cu) against outputs of the reference's own
    directionality_score (cyto_path.py:155-243): same nested list shape, nan in the same places, values to 1e-12
    (float64 sums in a different order; the reference's numba cosine is fastmath)."""
    from tests.test_oracle_golden import _directionality_case
    ad, nbhd, pos, want = _directionality_case(name)
    got = cytopath_b200.directionality_score(ad, nbhd, "end", pos, num_cores=1)
    assert [len(t) for t in got] == [len(t) for t in nbhd]
    assert all(len(a) == len(b) for t, u in zip(got, nbhd) for a, b in zip(t, u))
    flat = np.concatenate([np.concatenate(t) for t in got])
    assert np.array_equal(np.isnan(flat), np.isnan(want))
    np.testing.assert_allclose(flat, want, rtol=1e-12, atol=1e-15, equal_nan=True)


def test_f4_directionality_score_vs_oracle_large():
    from scipy import sparse
    from oracle import directionality_port as dp
    g = synth.branching_graph(n=6000, k=30, n_branches=3, seed=5, workers=1)
    rng = np.random.default_rng(0)
    T = g.T.tocsr()
    vg = sparse.csr_matrix((rng.random(T.nnz).astype(np.float32), T.indices, T.indptr), shape=T.shape)
    vgn = sparse.csr_matrix((-(rng.random(T.nnz) * (rng.random(T.nnz) < 0.3)).astype(np.float32), T.indices, T.indptr), shape=T.shape)
    vgn.eliminate_zeros()
    pos = rng.normal(size=(6000, 30))
    coords = {"trajectory_%d_coordinates" % i: np.cumsum(rng.normal(size=(40, 30)), axis=0) for i in range(4)}
    nbhd = [[rng.choice(6000, size=int(rng.integers(50, 400)), replace=False) for _ in range(40)] for _ in range(4)]

    class Shim:
        pass
    ad = Shim()
    ad.uns = {"T_forward": T, "velocity_graph": vg, "velocity_graph_neg": vgn, "trajectories": {"trajectories_coordinates": {"end": coords}}}
    got = cytopath_b200.directionality_score(ad, nbhd, "end", pos)
    want = dp.directionality_score(ad, nbhd, "end", pos)
    for t, u in zip(got, want):
        for a, b in zip(t, u):
            np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-14, equal_nan=True)


def test_release_cached_memory_then_reuse():
    """Large blocks are kept for exact-size reuse across calls (graph.cu); cyp_release_cached_memory returns them and
    the library keeps working afterwards."""
    C = sp.canonical_csr(synth.branching_graph(n=200_000, k=50, n_branches=4, seed=1, workers=1).T)     # ~80 MB per CDF array
    want = None
    for _ in range(3):
        g = _lib.Graph(C, _lib.CDF_EXACT)
        got = g.walk(np.arange(16), 8, 30, seed=4)
        want = got if want is None else want
        np.testing.assert_array_equal(got, want)
        g.close()
        _lib.release_cached_memory()
    _lib.release_cached_memory(0)
