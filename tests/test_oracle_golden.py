"""Pin the CPU oracle (oracle/) against outputs of the REAL reference (tests/golden/,
made by oracle/make_golden.py from /root/reference/cytopath/markov_chain_sampler.py)."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from oracle import cwalk, philox, sampler_port as sp
from tests.helpers import SAMPLING_CASES, adata_from_case, csr_from, load


# Random123 known-answer vectors for Philox4x32-10
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_kat_numpy(ctr, key, want):
    got = philox.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
    assert tuple(int(g) for g in got) == want


def test_uniform_stream_numpy_equals_c():
    u = philox.walker_uniforms(0x1234567890ABCDEF, 3, 2**32 - 2, 5, 9)      # crosses the 32-bit walker boundary
    for i in range(5):
        for t in range(9):
            assert u[i, t] == cwalk.uniform(0x1234567890ABCDEF, 3, 2**32 - 2 + i, t)
    assert (u >= 0).all() and (u < 1).all()


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_matrix_prep_matches_reference(tag):
    z = load("matrix_prep_messy_" + tag)
    C = sp.canonical_csr(csr_from(z))
    for cdf in (sp.build_cdf(C, sp.CDF_REFERENCE), cwalk.build_cdf(C, 0)):
        idx, ell = sp.to_ell(C, cdf)
        assert idx.dtype == z["idx"].dtype and ell.dtype == z["cdf"].dtype
        np.testing.assert_array_equal(idx, z["idx"])
        np.testing.assert_array_equal(ell, z["cdf"])
    np.testing.assert_array_equal(sp.build_cdf(C, sp.CDF_EXACT), cwalk.build_cdf(C, 1))


def test_markov_sim_matches_reference():
    z = load("markov_sim_small")
    C = sp.canonical_csr(csr_from(z))
    cdf = sp.build_cdf(C, sp.CDF_REFERENCE)
    idx, ell = sp.to_ell(C, cdf)
    np.testing.assert_array_equal(idx, z["idx"])
    np.testing.assert_array_equal(ell, z["cdf"])
    sim, steps, seed = int(z["sim"]), int(z["steps"]), int(z["seed"])
    roots = z["roots"]
    starts = np.repeat(roots, sim)
    u = philox.walker_uniforms(seed, 0, 0, len(starts), steps - 1)
    np.testing.assert_array_equal(sp.walk(C, cdf, starts, u), z["seq"])                     # numpy restatement
    np.testing.assert_array_equal(cwalk.walk(C, cdf, starts, steps - 1, seed=seed), z["seq"])   # C, own Philox
    np.testing.assert_array_equal(cwalk.walk(C, cdf, starts, steps - 1, uniforms=u), z["seq"])  # C, injected
    # faithful loop port, second root only
    st, sc, cl = sp.markov_sim_port(1, sim, steps, roots, z["clusters"].astype(object), C, idx, ell,
                                    uniforms=u[sim:])
    np.testing.assert_array_equal(st, z["seq"][sim:])
    np.testing.assert_array_equal(cl, z["labels"][sim:])
    np.testing.assert_array_equal(sc, z["score"][sim:])
    np.testing.assert_array_equal(sp.transition_scores(C, z["seq"]), z["score"])


def test_convergence_probes():
    z = load("convergence_probes")
    for i, want in enumerate(z["result"]):
        got = sp.check_convergence_criteria(z["p%d" % i], tol=1e-3)
        assert (-1 if got is None else got) == want


@pytest.mark.parametrize("name", SAMPLING_CASES)
@pytest.mark.parametrize("stepper", ["numpy", "c"])
def test_sampling_port_matches_reference(name, stepper):
    z = load(name)
    if name == "sampling_config1_auto" and stepper == "numpy":
        pytest.skip("numpy stepper is slow at config-1 size; the C stepper covers it")
    ad, kw = adata_from_case(z)
    seed = int(z["seed"])
    walk_fn = None
    if stepper == "c":
        def walk_fn(C, cdf, starts, rnd, begin, nt):
            return cwalk.walk(C, cdf, starts, nt, seed=seed, round_idx=rnd, walker_begin=begin)
    np.random.seed(int(z["np_seed"]))
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp.sampling_port(ad, **kw, seed=seed, walk_fn=walk_fn)
    s = ad.uns["samples"]
    assert ad.uns["run_info"]["simulation_steps"] == int(z["simulation_steps"])
    np.testing.assert_array_equal(np.asarray(ad.uns["run_info"]["root_cells"]), z["run_root_cells"])
    np.testing.assert_array_equal(np.asarray(ad.uns["run_info"]["end_points"]), z["run_end_points"])
    assert s["cell_sequences"].dtype == np.int32
    np.testing.assert_array_equal(s["cell_sequences"], z["cell_sequences"])
    np.testing.assert_array_equal(s["cluster_sequences"], z["cluster_sequences"])
    assert s["cluster_sequences"].dtype == z["cluster_sequences"].dtype
    assert s["transition_score"].dtype == np.float32
    np.testing.assert_allclose(s["transition_score"], z["transition_score"], rtol=2e-6)


def _directionality_case(name):
    from scipy import sparse
    z = load(name)
    n = int(z["T_shape"][0])
    mk = lambda p: sparse.csr_matrix((z[p + "_data"], z[p + "_indices"], z[p + "_indptr"]), shape=(n, n))
    coords, sizes, cells = z["coords"], z["sizes"], z["cells"]

    class Shim:
        pass
    ad = Shim()
    ad.uns = {"T_forward": mk("T"), "velocity_graph": mk("vg"), "velocity_graph_neg": mk("vgn"),
              "trajectories": {"trajectories_coordinates": {"end": {"trajectory_%d_coordinates" % i: coords[i] for i in range(len(coords))}}}}
    nbhd, o = [], 0
    for i in range(sizes.shape[0]):
        t = []
        for j in range(sizes.shape[1]):
            t.append(cells[o:o + sizes[i, j]])
            o += sizes[i, j]
        nbhd.append(t)
    return ad, nbhd, z["pos"], z["scores"]


@pytest.mark.parametrize("name", ["directionality_d2_f32", "directionality_d10_f64"])
def test_directionality_port_matches_reference_golden(name):
    """f4: the numpy restatement of cyto_path.py:155-243 against outputs of the reference's own function (numba
    fastmath cosine: agreement to 1e-12, nan for cells without graph entries in the same places)."""
    from oracle import directionality_port as dp
    ad, nbhd, pos, want = _directionality_case(name)
    got = np.concatenate([np.concatenate(t) for t in dp.directionality_score(ad, nbhd, "end", pos)])
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).any() and (~np.isnan(want)).sum() > 100
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15, equal_nan=True)


@pytest.mark.parametrize("exact", [True, False])
def test_fused_rows_port_is_self_consistent(exact):
    """The numpy restatement of the device layout (oracle/fused_rows_port.py), decoded again: the 12-bit codes bracket
    the oracle's CDF (exact mode) or equal 4 m (reference mode), the header fields are right, the inlined targets are the
    row words of the most probable entries, every target word points at the block of its column."""
    from cytopath_b200 import synth
    from oracle import fused_rows_port as frp
    C = sp.canonical_csr(synth.hub_graph(n=600, hub_share=0.03, seed=3))
    cdf = cwalk.build_cdf(C, 1 if exact else 0)
    rows, words, targets = frp.build(C, cdf, exact)
    first = (words & np.uint32((1 << 27) - 1)).astype(np.int64)
    units = np.concatenate([[0], np.cumsum((np.diff(C.indptr) + 3) // 4)[:-1]])
    for r in range(C.shape[0]):
        a, b = C.indptr[r], C.indptr[r + 1]
        n = b - a
        blk = rows[first[r] * 32:]
        assert int(np.frombuffer(blk[0:4].tobytes(), dtype=np.int32)[0]) == r
        assert int(np.frombuffer(blk[4:6].tobytes(), dtype=np.uint16)[0]) == n
        m = int(blk[7])
        mask = int(np.frombuffer(blk[8:16].tobytes(), dtype=np.uint64)[0])
        assert bin(mask).count("1") == m and int(np.frombuffer(blk[16:20].tobytes(), dtype=np.uint32)[0]) == units[r]
        hi = blk[20:20 + n].astype(np.int64)
        lo_bytes = blk[20 + n:20 + n + (n + 1) // 2].astype(np.int64)
        lo = np.stack([lo_bytes & 15, lo_bytes >> 4], axis=1).ravel()[:n]
        code = (hi << 4) | lo
        c = np.asarray(cdf[a:b], dtype=np.float64)
        if exact:
            assert np.all(code / 4096.0 <= c) and np.all((c < (code + 1) / 4096.0) | (code == 4095))
        else:
            np.testing.assert_array_equal(code, 4 * np.rint(c * 1000.0).astype(np.int64))
        tw = words[C.indices[a:b]]
        np.testing.assert_array_equal(targets[units[r] * 4:units[r] * 4 + n], tw)
        inl_j = [j for j in range(min(n, 64)) if (mask >> j) & 1]
        used = frp.used_bytes(n)
        np.testing.assert_array_equal(np.frombuffer(blk[used:used + 4 * m].tobytes(), dtype=np.uint32), tw[inl_j])
        if 0 < m < min(n, 64):                                       # the inlined entries are the most probable ones
            p = C.data[a:b][:64]
            assert p[inl_j].min() >= np.delete(p, inl_j).max()
