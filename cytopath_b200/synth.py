"""Synthetic velocity graphs and a duck-typed AnnData for tests and benchmarks.

Nothing here is on the product path: `sampling()` never imports this module.  The
generators follow the recipe in SURVEY.md section 8(d) ("branching" and "cycling +
convergent" families).  They stand in for scvelo's `T_forward`, which the reference
never computes itself (it only reads `adata.uns[matrix_key]`,
reference cytopath/markov_chain_sampler.py:152-170).
"""
from __future__ import annotations

import copy as _copy
from dataclasses import dataclass

import numpy as np
import pandas as pd
from scipy import sparse
from scipy.spatial import cKDTree


class AnnDataShim:
    """The ~10 attributes of anndata.AnnData that cytopath's sampler touches.

    anndata is not installed in this image; the reference (and our drop-in) only use
    `.X`, `.shape`, `.obs` (a DataFrame with a RangeIndex, see SURVEY.md 8c caveat 2),
    `.uns` and `.copy()`.
    """

    def __init__(self, n_obs: int, obs: pd.DataFrame, uns: dict | None = None, X=None):
        self.X = np.zeros((n_obs, 1), dtype=np.float32) if X is None else X
        self.obs = obs
        self.uns = {} if uns is None else uns
        self.obsm = {}

    @property
    def shape(self):
        return (self.obs.shape[0], self.X.shape[1])

    def copy(self):
        out = AnnDataShim(self.obs.shape[0], self.obs.copy(), None, self.X.copy())
        out.uns = {k: _copy.deepcopy(v) for k, v in self.uns.items()}
        out.obsm = {k: v.copy() for k, v in self.obsm.items()}
        return out


@dataclass
class SynthGraph:
    T: sparse.csr_matrix          # row-stochastic transition matrix (float64, canonical CSR)
    clusters: np.ndarray          # str labels per cell (<= 8 chars)
    root_prob: np.ndarray         # stand-in for obs['root_cells']
    end_prob: np.ndarray          # stand-in for obs['end_points']
    pos: np.ndarray               # 2-D embedding (downstream smoke tests use it as X_pca)
    root_clusters: list
    end_clusters: list

    def to_adata(self, cluster_key: str = "louvain", matrix_dtype=None) -> AnnDataShim:
        obs = pd.DataFrame(
            {
                cluster_key: pd.Categorical(self.clusters),
                "root_cells": self.root_prob,
                "end_points": self.end_prob,
            }
        )
        T = self.T if matrix_dtype is None else self.T.astype(matrix_dtype)
        ad = AnnDataShim(self.T.shape[0], obs, {"T_forward": T})
        ad.obsm["X_pca"] = self.pos.copy()
        return ad


def _knn_transition(pos, vel, k, scale=10.0, self_loops=True, workers=-1, layer=None):
    """kNN graph with scvelo-like exp(scale * cos(velocity, displacement)) weights.  With `layer`
    (an integer per cell) neighbours are searched within the cell's layer only: L interleaved,
    independent sub-graphs whose hop length is ~sqrt(L) times longer, so that terminal regions stay
    reachable in a few hundred steps at atlas scale while every row of the big matrix is still used."""
    n = pos.shape[0]
    if layer is None:
        tree = cKDTree(pos)
        _, nbr = tree.query(pos, k=k + 1, workers=workers)       # column 0 is the cell itself
    else:
        nbr = np.empty((n, k + 1), dtype=np.int64)
        for l in np.unique(layer):
            idx = np.flatnonzero(layer == l)
            _, loc = cKDTree(pos[idx]).query(pos[idx], k=k + 1, workers=workers)
            nbr[idx] = idx[loc]
    nbr = nbr.astype(np.int64)
    disp = pos[nbr] - pos[:, None, :]
    norm = np.linalg.norm(disp, axis=2)
    cosv = np.einsum("nkd,nd->nk", disp, vel) / np.where(norm > 0, norm, 1.0)
    cosv[nbr == np.arange(n)[:, None]] = 0.0                  # self-loop: cos := 0
    w = np.exp(scale * cosv)
    if not self_loops:
        w[nbr == np.arange(n)[:, None]] = 0.0
    w /= w.sum(axis=1, keepdims=True)
    indptr = np.arange(0, (k + 1) * n + 1, k + 1, dtype=np.int64)
    T = sparse.csr_matrix((w.ravel(), nbr.ravel().astype(np.int32), indptr), shape=(n, n))
    T.sum_duplicates()
    T.eliminate_zeros()
    T.sort_indices()
    return T


def branching_graph(n=3696, k=30, n_branches=4, seed=0, n_root_stems=1, self_loops=True,
                    shuffle=True, workers=-1, scale=10.0, noise=0.01, n_layers=1) -> SynthGraph:
    """Stem that splits into `n_branches` rays; velocity points along pseudotime.

    `n_root_stems=2` adds a second stem (offset in y) that joins the first at the
    branch point, giving two root clusters (BASELINE.json configs[1]).
    Cell order is random (as in real data) so neighbour indices carry no locality.
    """
    rng = np.random.default_rng(seed)
    t = rng.random(n)
    b = rng.integers(n_branches, size=n)
    stem_id = rng.integers(n_root_stems, size=n)
    spread = min(0.5, 2.6 / n_branches)
    ang = (b - (n_branches - 1) / 2.0) * spread
    on_stem = t < 0.3
    # stems fan in towards the branch point (0.3, 0)
    stem_ang = (stem_id - (n_root_stems - 1) / 2.0) * 0.6
    sx = 0.3 - (0.3 - t) * np.cos(stem_ang)
    sy = -(0.3 - t) * np.sin(stem_ang)
    bx = 0.3 + (t - 0.3) * np.cos(ang)
    by = (t - 0.3) * np.sin(ang)
    pos = np.where(on_stem[:, None], np.stack([sx, sy], 1), np.stack([bx, by], 1))
    pos = pos + rng.normal(0.0, noise, size=(n, 2))
    vel = np.where(on_stem[:, None],
                   np.stack([np.cos(stem_ang), np.sin(stem_ang)], 1),
                   np.stack([np.cos(ang), np.sin(ang)], 1))
    if not shuffle:                                   # optional locality-preserving order
        order = np.argsort(t, kind="stable")
        t, b, stem_id, on_stem, pos, vel = (a[order] for a in (t, b, stem_id, on_stem, pos, vel))
    layer = rng.integers(n_layers, size=n) if n_layers > 1 else None
    T = _knn_transition(pos, vel, k, scale=scale, self_loops=self_loops, workers=workers, layer=layer)

    stem_bin = np.minimum((t / 0.3 * 5).astype(int), 4)
    br_bin = np.minimum(((t - 0.3) / 0.7 * 5).astype(int), 4)
    clusters = np.empty(n, dtype=object)
    for i in range(n):
        if on_stem[i]:
            clusters[i] = ("stem%d" % stem_bin[i]) if stem_id[i] == 0 else ("s%dm%d" % (stem_id[i], stem_bin[i]))
        else:
            clusters[i] = "b%d_%d" % (b[i], br_bin[i])
    root_clusters = ["stem0"] + ["s%dm0" % s for s in range(1, n_root_stems)]
    end_clusters = ["b%d_4" % i for i in range(n_branches)]
    root_prob = np.isin(clusters, root_clusters).astype(np.float64)
    end_prob = np.isin(clusters, end_clusters).astype(np.float64)
    return SynthGraph(T, clusters.astype(str), root_prob, end_prob, pos, root_clusters, end_clusters)


def cycling_graph(n=20000, k=200, seed=0, workers=-1, n_layers=1) -> SynthGraph:
    """Annulus (cycle, tangential velocity) feeding two arms that re-converge to one
    terminal region; self-loop on every row; long rows (k=200) for the wide-row path.
    Roots / end points come from smooth probability bumps so the reference's 0.99
    thresholds (markov_chain_sampler.py:226,237) select a small set.
    """
    rng = np.random.default_rng(seed)
    part = rng.choice(4, size=n, p=[0.4, 0.2, 0.2, 0.2])   # 0 annulus, 1/2 arms, 3 terminal stem
    s = rng.random(n)
    pos = np.zeros((n, 2))
    vel = np.zeros((n, 2))
    th = 2 * np.pi * s
    r = 0.5
    m = part == 0
    pos[m] = np.stack([r * np.cos(th[m]) - r, r * np.sin(th[m])], 1)      # circle touching (0,0)
    vel[m] = np.stack([-np.sin(th[m]), np.cos(th[m])], 1)
    for arm, sign in ((1, 1.0), (2, -1.0)):
        m = part == arm
        x = s[m]
        y = sign * 0.35 * np.sin(np.pi * x)
        pos[m] = np.stack([x, y], 1)
        dy = sign * 0.35 * np.pi * np.cos(np.pi * x)
        v = np.stack([np.ones_like(x), dy], 1)
        vel[m] = v / np.linalg.norm(v, axis=1, keepdims=True)
    m = part == 3
    pos[m] = np.stack([1.0 + 0.5 * s[m], np.zeros(m.sum())], 1)
    vel[m] = np.array([1.0, 0.0])
    pos += rng.normal(0.0, 0.01, size=(n, 2))
    # n_layers interleaved, independent kNN graphs (see _knn_transition): hops n_layers times longer, so that the stem tip
    # stays reachable from the far side of the cycle in a couple of hundred steps at 250k cells
    layer = rng.integers(n_layers, size=n) if n_layers > 1 else None
    T = _knn_transition(pos, vel, k, self_loops=True, workers=workers, layer=layer)

    names = {0: "cyc", 1: "armA", 2: "armB", 3: "term"}
    bins = np.minimum((s * 4).astype(int), 3)
    clusters = np.array(["%s%d" % (names[p], q) for p, q in zip(part, bins)], dtype=str)
    # smooth bumps: root around annulus angle pi (far side of the cycle), end at the stem tip
    root_prob = np.where(part == 0, np.exp(-((th - np.pi) / 0.25) ** 2), 0.0)
    end_prob = np.where(part == 3, np.exp(-((1.0 - s) / 0.2) ** 2), 0.0)
    return SynthGraph(T, clusters, root_prob, end_prob, pos, ["cyc1", "cyc2"], ["term3"])


def random_sparse_rows(n=200, max_nnz=12, seed=0, messy=True, dtype=np.float64) -> sparse.csr_matrix:
    """Row-stochastic matrix with ragged rows; with `messy` the CSR has unsorted
    indices, duplicate entries and explicit zeros, which the reference's
    densify/re-sparsify (markov_chain_sampler.py:168,14) silently canonicalises."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for i in range(n):
        m = int(rng.integers(1, max_nnz + 1))
        c = rng.choice(n, size=m, replace=False)
        v = rng.random(m) + 1e-3
        if messy and m >= 3:
            c = np.concatenate([c, c[:1]])            # duplicate entry
            v = np.concatenate([v, rng.random(1)])
            c = np.concatenate([c, rng.choice(np.setdiff1d(np.arange(n), c), 1)])
            v = np.concatenate([v, [0.0]])            # explicit zero
        v = v / v.sum()
        p = rng.permutation(len(c))                   # unsorted
        rows.append(np.full(len(c), i)); cols.append(c[p]); vals.append(v[p])
    rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals).astype(dtype)
    order = np.argsort(rows, kind="stable")
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))])
    M = sparse.csr_matrix((vals[order], cols[order].astype(np.int32), indptr), shape=(n, n))
    return M


def hub_graph(n=3000, k_typ=(8, 40), hub_share=0.02, k_hub=(150, 300), seed=0) -> sparse.csr_matrix:
    """Row-stochastic matrix whose rows have `k_typ` entries except a few hubs with `k_hub` entries, with skewed
    probabilities: the shape of a real velocity graph (most rows ~30 entries, a few long ones)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(k_typ[0], k_typ[1] + 1, size=n)
    hubs = rng.random(n) < hub_share
    lens[hubs] = rng.integers(k_hub[0], k_hub[1] + 1, size=int(hubs.sum()))
    lens = np.minimum(lens, n)
    rows = np.repeat(np.arange(n), lens)
    cols = np.concatenate([np.sort(rng.choice(n, size=int(l), replace=False)) for l in lens])
    vals = np.exp(4.0 * rng.normal(size=len(cols)))
    M = sparse.csr_matrix((vals, (rows, cols)), shape=(n, n))
    M = sparse.diags(1.0 / np.asarray(M.sum(axis=1)).ravel()) @ M
    return M.tocsr()
