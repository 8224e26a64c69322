"""CPU restatement of cytopath's Markov-chain sampler on CSR.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows
(reference = /root/reference/cytopath/markov_chain_sampler.py).  The restatement never
densifies the n x n matrix (the reference does at :168), so it also runs at 100k - 1M
cells, where the reference itself cannot.  It is pinned against outputs of the real
reference by tests/test_oracle_golden.py (fixtures made by oracle/make_golden.py).
"""
from __future__ import annotations

import math
import warnings

import numpy as np
from scipy import sparse
from scipy.spatial.distance import cosine

from . import philox

CDF_REFERENCE = 0      # float32(cumsum_f64) rounded to 3 decimals   (reference :19,:23,:28,:31)
CDF_EXACT = 1          # float64 sequential cumsum                   (north-star mode)


# --------------------------------------------------------------------------- a1
def canonical_csr(T) -> sparse.csr_matrix:
    """What `sparse.csr_matrix(dense_float64_matrix)` (:14, after :168-170) holds:
    columns ascending, duplicates summed (in the matrix's own dtype, by `toarray`),
    explicit zeros dropped, values float64."""
    if sparse.issparse(T):
        C = sparse.csr_matrix(T, copy=True)
        C.sum_duplicates()
        C = C.astype(np.float64)
    else:
        C = sparse.csr_matrix(np.float64(np.asarray(T)))
    C.eliminate_zeros()
    C.sort_indices()
    return C


def build_cdf(C: sparse.csr_matrix, mode: int = CDF_REFERENCE) -> np.ndarray:
    """Per-row inclusive prefix sums, strictly left to right in float64 (`np.cumsum`, :18-19).
    Reference mode then narrows to float32 (:23,:28) and rounds to 3 decimals (:31)."""
    data = C.data.astype(np.float64)
    cdf = np.empty_like(data)
    indptr = C.indptr
    for r in range(C.shape[0]):
        a, b = indptr[r], indptr[r + 1]
        if b > a:
            cdf[a:b] = np.cumsum(data[a:b])
    if mode == CDF_EXACT:
        return cdf
    return np.round(cdf.astype(np.float32), decimals=3)


def to_ell(C: sparse.csr_matrix, cdf: np.ndarray):
    """The zero-padded [n, max_row_nnz] arrays matrix_prep returns (:22-28)."""
    n = C.shape[0]
    lens = np.diff(C.indptr)
    W = int(lens.max()) if n else 0
    idx = np.zeros((n, W), dtype=np.int32)
    val = np.zeros((n, W), dtype=cdf.dtype)
    col = np.arange(W)[None, :]
    mask = col < lens[:, None]
    idx[mask] = C.indices
    val[mask] = cdf
    return idx, val


# --------------------------------------------------------------------------- a2
def markov_sim_port(j, sim_number, max_steps, root_cells, clusters, C, idx_ell, cdf_ell, uniforms=None, prob_ell=None):
    """Faithful restatement of markov_sim (:35-52): same loop nest (step outer, walker
    inner), same `np.where(u <= row)[0][0]` first-crossing rule on the padded arrays.
    `C` (canonical CSR) replaces the dense matrix for the T[cur, nxt] lookup at :50; at
    sizes where even that is too slow to be fair to the reference (its dense lookup is one
    fancy index), `prob_ell` (raw probabilities in the same padded layout) gives the same
    number with the same cost.  This is the function the CPU baseline times."""
    prob_state = np.empty((sim_number, max_steps - 1), dtype=np.float32)
    state = np.empty((sim_number, max_steps), dtype=np.int32)
    clust_state = np.empty((sim_number, max_steps), dtype="U8")
    state[:, 0] = root_cells[j]
    clust_state[:, 0] = clusters[root_cells[j]]
    u = np.random.rand(sim_number, max_steps - 1) if uniforms is None else uniforms
    for i in range(1, max_steps):
        for k in range(sim_number):
            cur = state[k, i - 1]
            pos = np.where(u[k, i - 1] <= cdf_ell[cur, :])[0][0]
            nxt = idx_ell[cur, pos]
            state[k, i] = nxt
            prob_state[k, i - 1] = C[cur, nxt] if prob_ell is None else prob_ell[cur, pos]
            clust_state[k, i] = clusters[nxt]
    return state, np.sum(-np.log(prob_state), axis=1), clust_state


def walk(C: sparse.csr_matrix, cdf: np.ndarray, start_cells: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Vectorised selection rule of :48-49 on CSR: next = indices[row_start + first j
    with u <= cdf[row_start + j]] (ties go to the bin).  Raises IndexError like the
    reference when no bin satisfies the test (under-normalised row).
    start_cells [W], uniforms [W, steps-1] float64 -> int32 [W, steps]."""
    W, nt = uniforms.shape
    indptr = C.indptr.astype(np.int64)
    indices = C.indices
    lens = np.diff(indptr)
    Wd = int(lens.max())
    seq = np.empty((W, nt + 1), dtype=np.int32)
    seq[:, 0] = start_cells
    cur = np.asarray(start_cells, dtype=np.int64)
    ar = np.arange(Wd)[None, :]
    cdf64 = cdf.astype(np.float64)
    for t in range(nt):
        start = indptr[cur]
        ln = lens[cur]
        pos = np.minimum(start[:, None] + ar, len(cdf64) - 1)
        hit = (uniforms[:, t][:, None] <= cdf64[pos]) & (ar < ln[:, None])
        if not hit.any(axis=1).all():
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        j = hit.argmax(axis=1)
        cur = indices[start + j].astype(np.int64)
        seq[:, t + 1] = cur
    return seq


def transition_scores(C: sparse.csr_matrix, seq: np.ndarray) -> np.ndarray:
    """score = np.sum(-np.log(float32(T[c_t, c_t+1])), axis=1) in float32 (:50,:52)."""
    if seq.shape[0] == 0:
        return np.empty(0, dtype=np.float32)
    p = np.asarray(C[seq[:, :-1].ravel(), seq[:, 1:].ravel()]).reshape(seq.shape[0], -1).astype(np.float32)
    return np.sum(-np.log(p), axis=1)


# --------------------------------------------------------------------------- a5
def check_convergence_criteria(eps_history, tol=1e-5):
    """:54-60 — 1 + #differences - (index, from the end, of the last |diff| > tol);
    None if every difference is within tol."""
    diffs = np.abs(np.diff(np.asarray(eps_history, dtype=np.float64)))
    big = np.nonzero(diffs[::-1] > tol)[0]
    if big.size == 0:
        return None
    return int(1 + len(diffs) - big[0])


def auto_max_steps(T, init, stationary, max_iter=1000, tol=1e-5):
    """iterate_state_probability (:62-87), returning only its last element
    (what sampling uses at :275-278).  p <- p @ T with the matrix as stored."""
    eps = np.empty(max_iter)
    p = init
    for i in range(max_iter):
        p = sparse.csr_matrix.dot(p, T)
        eps[i] = cosine(p, stationary)
    c = check_convergence_criteria(eps, tol=tol)
    if isinstance(c, int) and c < max_iter:
        return c
    return max_iter


# --------------------------------------------------------------------------- a3, a4, a7
def sampling_port(adata, auto_adjust=True, matrix_key="T_forward", cluster_key="louvain", max_steps=10,
                  min_sim_ratio=0.6, rounds_limit=10, traj_number=200, sim_number=500,
                  end_point_probability=0.99, root_cell_probability=0.99, end_points=None, root_cells=None,
                  end_clusters=None, root_clusters=None, min_clusters=2, max_iter=200, tol=1e-3,
                  normalize=False, unique=True, num_cores=1, copy=False, *, seed=0, cdf_mode=CDF_REFERENCE,
                  walk_fn=None):
    """Restatement of sampling (:89-414).  Uniforms come from the frozen Philox stream
    (oracle/philox.py) instead of np.random.rand; the final with-replacement draw uses
    the global np.random like :395.  `walk_fn(C, cdf, starts, round, walker_begin, nt)`
    may replace the step loop (tests inject the C oracle there)."""
    data = adata
    adata = data.copy() if copy else data
    M = adata.uns[matrix_key]
    try:                                                                       # :159-164
        assert len(M.shape) == 2 and M.shape[0] == M.shape[1] and M.shape[0] == adata.shape[0]
    except Exception:
        raise ValueError("Transition matrix must be of shape num_cells x num_cells")
    C = canonical_csr(M)                                                       # :167-170 + :14
    n = C.shape[0]
    if normalize:                                                              # :176-180
        rs = np.asarray(np.float64(M.toarray() if sparse.issparse(M) else np.asarray(M)).sum(axis=1)).ravel() \
            if n <= 20000 else np.asarray(C.sum(axis=1)).ravel()
        C = sparse.csr_matrix((C.data / np.repeat(rs, np.diff(C.indptr)), C.indices, C.indptr), shape=C.shape)
        C.eliminate_zeros()
    else:                                                                      # :181-187
        rs = np.asarray(C.sum(axis=1)).ravel()
        if not (rs.max() <= 1.0 + 1e-3 and rs.min() >= 1.0 - 1e-3):
            raise ValueError("Transition matrix has not been normalized. Either pass normalize==True or provide a row sum normalized matrix.")
    try:                                                                       # :190-194
        assert adata.obs[cluster_key].dtype == "category"
    except Exception:
        raise ValueError("Categorical cluster labels must be provided in adata.obs[cluster_key]")
    if cluster_key != "louvain":                                               # :197-198
        warnings.warn("A finely resolved clustering is required to identify terminal regions. Consider using louvain clustering.", UserWarning)

    def as_list(x):                                                            # :202-212
        try:
            not x
        except Exception:
            x = list(x)
        return x

    root_cells, end_points = as_list(root_cells), as_list(end_points)
    root_clusters, end_clusters = as_list(root_clusters), as_list(end_clusters)

    clusters = np.asarray(adata.obs[cluster_key].astype(str).values, dtype=object)   # :217
    adata.uns["run_info"] = {"cluster_key": cluster_key}                       # :218
    cdf = build_cdf(C, cdf_mode)                                               # :220

    if not root_cells:                                                         # :223-231
        if not root_clusters:
            try:
                root_cells = np.asarray(np.where(np.asarray(adata.obs["root_cells"]) >= root_cell_probability))[0]
            except Exception:
                raise ValueError('Root cells must be provided in adata.obs["root_cells"]. Run cytopath.terminal_states')
        else:
            root_cells = np.asarray(np.where(np.asarray(adata.obs[cluster_key].isin(root_clusters))))[0]
    adata.uns["run_info"]["root_cells"] = root_cells
    if not end_points:                                                         # :234-257
        if not end_clusters:
            try:
                end_points = np.asarray(np.where(np.asarray(adata.obs["end_points"]) >= end_point_probability))[0]
            except Exception:
                raise ValueError('End points must be provided in adata.obs["end_points"]. Run cytopath.terminal_states')
        else:
            end_points = np.asarray(np.where(np.asarray(adata.obs[cluster_key].isin(end_clusters))))[0]
    adata.uns["run_info"]["end_points"] = end_points
    end_points_arr = np.asarray(end_points, dtype=np.int64)
    end_point_clusters = clusters[end_points_arr].astype(str)                  # :258
    end_clusters_ = np.unique(end_point_clusters)                              # :259
    if len(end_clusters_) == 0:                                                # :261-262
        raise ValueError("Terminal cluster set is empty.")

    if auto_adjust:                                                            # :265-279
        traj_number = math.ceil(np.log10(adata.shape[0]) * traj_number)
        sim_number = math.ceil(traj_number * len(end_clusters_) * np.log2(len(root_cells) + 1))
        max_steps = auto_max_steps(adata.uns[matrix_key],
                                   (adata.obs["root_cells"] / adata.obs["root_cells"].sum()).values,
                                   (adata.obs["end_points"] / adata.obs["end_points"].sum()).values,
                                   max_iter=max_iter, tol=tol)

    roots = np.asarray(root_cells, dtype=np.int64)
    n_roots = len(roots)
    sim_number_ = max(int(sim_number / n_roots), 1)                            # :285
    trunc = np.array([c[:8] for c in clusters], dtype="U8")                    # labels land in a U8 array (:39,:51)
    _, code = np.unique(trunc, return_inverse=True)

    glob_seq, glob_score = [], []
    traj_num = [0]
    count = 0
    while min(traj_num) < traj_number:                                         # :293
        rnd = count
        count += 1
        sim_number_ *= 2                                                       # :299
        W = n_roots * sim_number_
        starts = np.repeat(roots, sim_number_)                                 # root-major rows (:314-317)
        if walk_fn is None:
            seq = walk(C, cdf, starts, philox.walker_uniforms(seed, rnd, 0, W, max_steps - 1))
        else:
            seq = walk_fn(C, cdf, starts, rnd, 0, max_steps - 1)
        cs = np.sort(code[seq], axis=1)                                        # :322 distinct truncated labels
        distinct = 1 + (np.diff(cs, axis=1) != 0).sum(axis=1)
        seq = seq[distinct >= min_clusters]
        if unique and seq.shape[0]:                                            # :329-333 first occurrence, original order
            first = np.unique(seq, axis=0, return_index=True)[1]
            seq = seq[np.sort(first)]
        glob_seq.append(seq)
        all_seq = np.vstack(glob_seq)                                          # :340
        last = all_seq[:, -1]
        traj_num = []
        for name in end_clusters_:                                             # :342-354
            cells = end_points_arr[end_point_clusters == name]
            traj_num.append(int(sum((last == c).sum() for c in cells)))
        if count == rounds_limit + 1 and min(traj_num) < min_sim_ratio * traj_number:   # :360-367
            raise ValueError("Sampling failed. Try lowering min_sim_ratio or increase rounds_limit.")

    all_seq = np.vstack(glob_seq)                                              # :380
    last = all_seq[:, -1]
    out_seq = []
    for name in end_clusters_:                                                 # :385-403
        cells = end_points_arr[end_point_clusters == name]
        cand = np.concatenate([np.where(last == c)[0] for c in cells])
        pick = np.random.randint(cand.shape[0], size=traj_number)              # :395
        out_seq.append(all_seq[cand[pick]])
    cell_sequences = np.vstack(out_seq)
    adata.uns["run_info"]["simulation_steps"] = max_steps                      # :406-407
    adata.uns["run_info"]["didrun"] = True
    width = "U8" if n_roots == 1 else "U32"                                    # :39 vs :312
    adata.uns["samples"] = {
        "cell_sequences": cell_sequences,
        "transition_score": transition_scores(C, cell_sequences),
        "cluster_sequences": trunc[cell_sequences].astype(width),
    }
    return adata if copy else None
