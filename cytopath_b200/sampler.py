"""`sampling()` — drop-in for `cytopath.sampling` (reference cytopath/markov_chain_sampler.py:89-414).

Same positional signature, defaults, exceptions, printed progress and `adata.uns` outputs as
the reference; the work is done by the CUDA library (cytopath_b200/csrc) through its C ABI:

  a7  validation / argument resolution (:148-262)      host, this file (no densification)
  a1  matrix_prep (:11-31)                              K1  cyp_graph_create
  a5  auto max_steps (:62-87, :265-279)                 K4  cyp_power_iteration (SpMV^T + cosine on the device)
  a2  markov_sim (:35-52)                               K2  inside cyp_session_round
  a3  min_clusters / terminal / unique filters (:322-354)  K3  inside cyp_session_round
  a6  joblib fan-out over root cells (:306-319)         walkers sharded by id over ranks/GPUs
  a4  final with-replacement draw + write-back (:378-412)  host index arithmetic; only the
                                                           selected rows leave the GPU

Deviations, all documented in DESIGN.md: `adata.X` is not densified (:172-173); uniforms come
from the counter-based Philox stream (include/cytopath_b200.h) instead of the unseeded
per-process `np.random.rand` (:43); `num_cores` is accepted and ignored.
"""
from __future__ import annotations

import math
import os
import time
import warnings

import numpy as np
from scipy import sparse

from . import _lib

__all__ = ["sampling", "iterate_state_probability", "check_convergence_criteria"]

_CDF_MODES = {"reference": _lib.CDF_REFERENCE, "exact": _lib.CDF_EXACT}


# ------------------------------------------------------------------------------------ a5
def check_convergence_criteria(eps_history, tol=1e-5):
    """Same contract as the reference helper (:54-60): the iteration count after which every
    successive |difference| stays within `tol`, or None if none ever exceeded it."""
    eps = np.asarray(eps_history, dtype=np.float64)
    diffs = np.abs(eps[1:] - eps[:-1])
    over = np.flatnonzero(diffs > tol)
    if over.size == 0:
        return None
    return int(over[-1]) + 2


def iterate_state_probability(adata, matrix_key="T_forward", init=None, stationary=None, max_iter=1000, tol=1e-5, *, device=0):
    """Power iteration p <- p @ T with cosine distance to `stationary` (:62-87) on the GPU (K4).  Returns
    (history[:converged], history, converged) like the reference; `sampling` itself only needs `converged` and never
    materialises the [max_iter, n] history."""
    history = np.empty((max_iter, adata.shape[0]))
    eps_history = _lib.power_iteration(adata.uns[matrix_key], init, stationary, max_iter, device, history=history)
    converged = check_convergence_criteria(eps_history, tol=tol)
    if isinstance(converged, int) and converged < max_iter:
        print("Tolerance reached after {} iterations of {}.".format(converged, max_iter))
    else:
        print("Max number ({}) of iterations reached.".format(max_iter))
        converged = max_iter
    return history[:converged], history, converged


# ------------------------------------------------------------------------------------ a7
def _canonical_csr(M) -> sparse.csr_matrix:
    """The matrix the reference actually samples from: it densifies (:168-170) and
    re-sparsifies (:14), which sorts columns, sums duplicates (in the input dtype) and drops
    explicit zeros.  Same result here without the n x n array."""
    if sparse.issparse(M):
        if (M.format == "csr" and M.dtype in (np.float32, np.float64) and M.has_canonical_format
                and np.count_nonzero(M.data) == M.nnz):
            return M            # already canonical: used in place, never modified (float32 is widened on the device)
        C = sparse.csr_matrix(M, copy=True)
        if not C.has_canonical_format:
            C.sum_duplicates()
        C = C.astype(np.float64)
    else:
        C = sparse.csr_matrix(np.float64(np.asarray(M)))
    C.eliminate_zeros()
    C.sort_indices()
    return C


def _row_sums(C) -> np.ndarray:
    out = np.zeros(C.shape[0], dtype=np.float64)
    nz = np.flatnonzero(np.diff(C.indptr))
    if nz.size:
        out[nz] = np.add.reduceat(C.data.astype(np.float64, copy=False), C.indptr[nz])
    return out


def _listify(x):
    """The reference's `try: not x / except: list(x)` idiom (:202-212): numpy arrays (whose
    truth value is ambiguous) become lists, everything else passes through."""
    try:
        not x
    except Exception:
        x = list(x)
    return x


# ------------------------------------------------------------------------------------ sampling
def sampling(data, auto_adjust=True, matrix_key="T_forward", cluster_key="louvain", max_steps=10, min_sim_ratio=0.6,
             rounds_limit=10, traj_number=200, sim_number=500, end_point_probability=0.99, root_cell_probability=0.99,
             end_points=None, root_cells=None, end_clusters=None, root_clusters=None, min_clusters=2, max_iter=200,
             tol=1e-3, normalize=False, unique=True, num_cores=1, copy=False, *, seed=0, cdf_mode="reference",
             device=None, devices=None, dist="auto", _backend=None, _group=None, _rows=True):
    """Markov sampling of cell sequences from root cells to terminal regions on B200 GPUs.

    Positional arguments, defaults, errors and outputs are those of `cytopath.sampling`
    (reference cytopath/markov_chain_sampler.py:89-146).  Keyword-only extras:

    seed : int
        Key of the Philox4x32-10 uniform stream (walker, step and round are the counter).
    cdf_mode : 'reference' | 'exact'
        'reference' reproduces the reference's float32, 3-decimal CDF bit for bit (:23,:31);
        'exact' keeps the float64 cumulative sums.  In 'exact' mode a row whose float64 sum is below 1 leaves a gap
        above its last CDF value, and a uniform that falls into it raises like the reference's IndexError at :49
        -- at 1e9+ steps even a float32 rounding error of 1e-7 is hit, so pass normalize=True with float32
        matrices (the 3-decimal rounding of 'reference' mode closes such gaps by itself).
    device : int, optional
        CUDA device of a single-GPU run (default 0; LOCAL_RANK under torchrun).
    devices : sequence of int, optional
        Several GPUs driven by THIS process -- a plain function call, no launcher: every round's walkers are cut
        into one block per device, every device uploads its share of the matrix and an all-gather over NVLink completes it,
        the devices exchange 24-byte records of the kept sequences with NCCL all-gathers (the reference's joblib
        fan-out, :306-319).  The result is bit-identical for any number of devices.
    dist : 'auto' | bool
        One process per GPU (torchrun): every process calls `sampling` with the same arguments, every process
        receives the same outputs.  'auto' uses an initialised torch.distributed group if there is one.

    Writes adata.uns['run_info'] and adata.uns['samples'] ('cell_sequences' int32,
    'transition_score' float32, 'cluster_sequences' <U8/<U32) and returns `adata` if `copy`.
    """
    adata = data.copy() if copy else data
    if cdf_mode not in _CDF_MODES:
        raise ValueError("cdf_mode must be 'reference' or 'exact'")

    # -- checks (:152-198)
    try:
        M = adata.uns[matrix_key]
    except NameError:
        raise ValueError("Transition matrix must be provided in adata.uns[matrix_key].")
    try:
        assert len(M.shape) == 2
        assert M.shape[0] == M.shape[1]
        assert M.shape[0] == adata.shape[0]
    except Exception:
        raise ValueError("Transition matrix must be of shape num_cells x num_cells")

    backend = _backend if _backend is not None else _lib
    t_start = time.perf_counter()
    group = _group if _group is not None else backend.open_group(device=device, devices=devices, dist=dist)
    try:
        t_group = time.perf_counter()
        # A float CSR is taken as it is: K1 reads every entry anyway and reports whether the matrix was canonical
        # (cyp_graph_canonical), so the host does not walk 50M entries to find out what is almost always true.
        trusted = (not normalize and sparse.issparse(M) and M.format == "csr" and M.dtype in (np.float32, np.float64))
        C = M if trusted else _canonical_csr(M)
        pristine = C is M                 # the stored matrix IS what is sampled from: K4 can reuse the graph's copy
        if normalize:                                                          # :176-180
            row_sums = _row_sums(C)
            with np.errstate(divide="ignore", invalid="ignore"):
                C = sparse.csr_matrix((C.data.astype(np.float64, copy=False) / np.repeat(row_sums, np.diff(C.indptr)),
                                       C.indices, C.indptr), shape=C.shape)
            C.eliminate_zeros()
            pristine = False
        # a1 (K1): every rank of the group holds the matrix, so each uploads 1/world of it, an all-gather over NVLink
        # completes it and every device builds the CDF itself.  The row sums of the normalisation check (:181-187)
        # come back from the same kernel, so the host never walks the 50M-entry value array.
        t_csr = time.perf_counter()
        timing = {"graph": group.create_graph(C, _CDF_MODES[cdf_mode], sharded=True)}
        if trusted and not group.graph.canonical():       # unsorted or duplicate columns, or explicit zeros: the rare slow path
            C = _canonical_csr(M)
            pristine = C is M
            timing["graph"] = group.create_graph(C, _CDF_MODES[cdf_mode], sharded=True)
            timing["graph"]["rebuilt_from_canonical_form"] = True
        timing["host_ms"] = {"open_group": 1e3 * (t_group - t_start), "canonical_csr": 1e3 * (t_csr - t_group),
                             "create_graph": 1e3 * (time.perf_counter() - t_csr)}
        if not normalize:
            lo, hi = group.graph.row_sum_range()
            if not (hi <= 1.0 + 1e-3 and lo >= 1.0 - 1e-3):
                raise ValueError("Transition matrix has not been normalized. Either pass normalize==True or provide a row sum normalized matrix.")
        return _sampling_on_graph(adata, copy, group, C, pristine, auto_adjust, matrix_key, cluster_key, max_steps,
                                  min_sim_ratio, rounds_limit, traj_number, sim_number, end_point_probability,
                                  root_cell_probability, end_points, root_cells, end_clusters, root_clusters, min_clusters,
                                  max_iter, tol, unique, seed, timing, _rows)
    finally:
        if _group is None:
            group.close()


def _auto_max_steps(group, T, pristine, init, stationary, max_iter, tol):
    """The only thing `sampling` keeps from iterate_state_probability is its last element (:275-278): max_steps.
    The power iteration runs on the GPU (K4) without the [max_iter, n] history buffer (1.6 GB at 1M cells): on the
    matrix the graph already holds when that is the stored matrix, else on an upload of adata.uns[matrix_key]."""
    if pristine:
        eps_history = group.graph.power_iteration(init, stationary, max_iter)
    else:
        eps_history = group.power_iteration_matrix(T, init, stationary, max_iter)
    converged = check_convergence_criteria(eps_history, tol=tol)
    if isinstance(converged, int) and converged < max_iter:
        print("Tolerance reached after {} iterations of {}.".format(converged, max_iter))
        return converged
    print("Max number ({}) of iterations reached.".format(max_iter))
    return max_iter


def _sampling_on_graph(adata, copy, group, C, pristine, auto_adjust, matrix_key, cluster_key, max_steps,
                       min_sim_ratio, rounds_limit, traj_number, sim_number, end_point_probability, root_cell_probability,
                       end_points, root_cells, end_clusters, root_clusters, min_clusters, max_iter, tol, unique, seed, timing,
                       want_rows=True):
    """Everything of `sampling` after the transition matrix has been accepted (:189-412)."""
    n = C.shape[0]
    try:
        adata.obs[cluster_key]
        assert adata.obs[cluster_key].dtype == "category"
    except Exception:
        raise ValueError("Categorical cluster labels must be provided in adata.obs[cluster_key]")
    if cluster_key != "louvain":
        warnings.warn("A finely resolved clustering is required to identify terminal regions. Consider using louvain clustering.", UserWarning)

    root_cells, end_points = _listify(root_cells), _listify(end_points)
    root_clusters, end_clusters = _listify(root_clusters), _listify(end_clusters)

    # -- roots, end points, terminal clusters (:217-262)
    t_marks = [time.perf_counter()]
    labels = _Labels(adata.obs[cluster_key])            # str(label) per cell, through the categories: no per-cell Python work
    adata.uns["run_info"] = {"cluster_key": cluster_key}
    if not root_cells:
        if not root_clusters:
            try:
                root_cells = np.asarray(np.where(np.asarray(adata.obs["root_cells"]) >= root_cell_probability))[0]
            except Exception:
                raise ValueError('Root cells must be provided in adata.obs["root_cells"]. Run cytopath.terminal_states')
        else:
            root_cells = labels.cells_in(adata.obs[cluster_key], root_clusters)
    adata.uns["run_info"]["root_cells"] = root_cells
    if not end_points:
        if not end_clusters:
            try:
                end_points = np.asarray(np.where(np.asarray(adata.obs["end_points"]) >= end_point_probability))[0]
            except Exception:
                raise ValueError('End points must be provided in adata.obs["end_points"]. Run cytopath.terminal_states')
        else:
            end_points = labels.cells_in(adata.obs[cluster_key], end_clusters)
    adata.uns["run_info"]["end_points"] = end_points
    end_cells = np.asarray(end_points, dtype=np.int64).reshape(-1)
    # terminal clusters: the distinct labels of the end points, sorted, and for every end point which of them it has
    end_clusters_, end_point_cluster = labels.groups_of(end_cells)
    if len(end_clusters_) == 0:
        raise ValueError("Terminal cluster set is empty.")
    roots = np.asarray(root_cells, dtype=np.int64).reshape(-1)
    n_roots = len(roots)
    if n_roots == 0:
        raise ZeroDivisionError("division by zero")          # what :285 raises when no root cell qualifies

    # -- auto adjust (:265-279)
    if auto_adjust:
        print("Adjusting simulation parameters based on dataset properties. Set auto_adjust=False if this is unwanted.")
        traj_number = math.ceil(np.log10(adata.shape[0]) * traj_number)
        print("Number of required simulations per end point (traj_number) set to {}".format(traj_number))
        sim_number = math.ceil(traj_number * len(end_clusters_) * np.log2(n_roots + 1))
        print("Number of initial simulations (sim_number) set to {}".format(sim_number))
        max_steps = _auto_max_steps(group, adata.uns[matrix_key], pristine,
                                    (adata.obs["root_cells"] / adata.obs["root_cells"].sum()).values,
                                    (adata.obs["end_points"] / adata.obs["end_points"].sum()).values, max_iter, tol)
        print("Number of initial simulation steps (max_steps) set to {}".format(max_steps))
    max_steps = int(max_steps)
    t_marks.append(time.perf_counter())

    # -- device-side state
    trunc, cell_code, n_codes = labels.truncated()               # labels are stored into a U8 array (:39,:51)
    # end-point slots: one slot per distinct end-point cell; the reference scans end points in
    # list order (:348-349, :390-391), duplicates in the list count twice.
    slot_cells, end_slot_of_point = np.unique(end_cells, return_inverse=True)
    end_slot_of_cell = np.full(n, -1, dtype=np.int32)
    end_slot_of_cell[slot_cells] = np.arange(len(slot_cells), dtype=np.int32)

    group.create_session(roots, cell_code, n_codes, int(min_clusters), end_slot_of_cell,
                         len(slot_cells), max_steps, bool(unique), int(seed))
    t_marks.append(time.perf_counter())
    timing["host_ms"].update(resolve_and_auto_adjust=1e3 * (t_marks[1] - t_marks[0]), labels_and_session=1e3 * (t_marks[2] - t_marks[1]))
    return _run_rounds(adata, copy, group, roots, end_slot_of_point, end_point_cluster, end_clusters_, trunc,
                       max_steps, traj_number, sim_number, min_sim_ratio, rounds_limit, timing, slot_cells, want_rows)


class _Labels:
    """The cluster label of every cell as the reference sees it (`adata.obs[cluster_key].astype(str)`, :217) and cut
    to 8 characters (what markov_sim stores, :39,:51), computed per CATEGORY and expanded through the category codes:
    a million cells cost a few integer gathers instead of a million Python string operations."""

    def __init__(self, col):
        self.cats = self.codes = None
        try:
            cats = np.array([str(c) for c in col.cat.categories], dtype=object)
            codes = np.asarray(col.cat.codes)
            if cats.size and codes.size and codes.min() >= 0:
                self.cats, self.codes = cats, codes
        except Exception:
            pass
        if self.cats is None:                                    # not categorical after all, or missing values: per cell
            self.full = np.array([str(x) for x in col.astype(str).values], dtype=object)    # (a missing value reads 'nan')

    def cells_in(self, col, clusters) -> np.ndarray:
        """np.where(col.isin(clusters))[0] (:225,:236).  pandas matches `clusters` against the categories and then looks
        the codes up; with every cell labelled that is one table lookup per cell."""
        if self.cats is not None:
            in_cat = np.asarray(col.cat.categories.isin(clusters), dtype=bool)
            return np.flatnonzero(in_cat[self.codes])
        return np.asarray(np.where(np.asarray(col.isin(clusters))))[0]

    def of(self, cells) -> np.ndarray:
        out = self.cats[self.codes[cells]] if self.cats is not None else self.full[cells]
        return out.astype(str)

    def groups_of(self, cells):
        """(np.unique(labels of `cells`), index into it for every cell of `cells`)."""
        if cells.size == 0:
            return np.empty(0, dtype=str), np.empty(0, dtype=np.int64)
        if self.cats is None:
            return np.unique(self.of(cells), return_inverse=True)
        names, cat_group = np.unique(self.cats.astype(str), return_inverse=True)    # categories may share a str()
        group = cat_group[self.codes[cells]]
        present = np.flatnonzero(np.bincount(group, minlength=len(names)))
        local = np.full(len(names), -1, dtype=np.int64)
        local[present] = np.arange(len(present))
        return np.array([str(x) for x in names[present]], dtype=str), local[group]

    def truncated(self):
        """(lookup cells -> '<U8' labels, code of that label per cell, number of codes) -- codes number the distinct
        truncated labels in sorted order, like np.unique(..., return_inverse=True)."""
        if self.cats is not None:
            cut = np.array([c[:8] for c in self.cats], dtype="U8")
            names, cat_code = np.unique(cut, return_inverse=True)
            present = np.zeros(len(names), dtype=bool)
            present[cat_code[np.flatnonzero(np.bincount(self.codes, minlength=len(cut)))]] = True   # categories without cells do not count as labels
            renum = np.cumsum(present) - 1
            codes = self.codes
            return (lambda cells, width="U8": _gather_labels(cut, codes, cells, width)), renum[cat_code].astype(np.int32)[codes], int(present.sum())
        trunc = np.array([c[:8] for c in self.full], dtype="U8")
        names, code = np.unique(trunc, return_inverse=True)
        return (lambda cells, width="U8": trunc[cells].astype(width)), code.astype(np.int32), len(names)


def _gather_labels(cut, codes, cells, width):
    """cut[codes[cells]].astype(width) -- `cut` holds one '<U8' label per category.  The [sequences, steps] result of a
    1M-cell run is 250 MB of '<U32' (:312): most of the time goes into touching that memory for the first time, so the
    rows are gathered as plain uint32 words (no per-element unicode cast, the GIL is released) by a few threads."""
    cells = np.asarray(cells)
    n_chars = np.dtype(width).itemsize // 4
    if n_chars < 8 or cells.ndim == 0:
        return cut[codes[cells]].astype(width)
    table = np.zeros((len(cut), n_chars), dtype=np.uint32)
    table[:, :8] = np.ascontiguousarray(cut).view(np.uint32).reshape(len(cut), 8)
    flat = cells.reshape(-1)
    out = np.empty((flat.shape[0], n_chars), dtype=np.uint32)
    n_threads = min(8, os.cpu_count() or 1) if out.nbytes >= (32 << 20) else 1

    def work(bounds):
        a, b = bounds
        np.take(table, codes[flat[a:b]], axis=0, out=out[a:b], mode="clip")      # (codes are category codes: always in range)

    if n_threads == 1:
        work((0, flat.shape[0]))
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(n_threads) as pool:
            list(pool.map(work, [(flat.shape[0] * i // n_threads, flat.shape[0] * (i + 1) // n_threads) for i in range(n_threads)]))
    return out.view(width).reshape(cells.shape)


def _run_rounds(adata, copy, group, roots, end_slot_of_point, end_point_cluster, end_clusters_, trunc, max_steps,
                traj_number, sim_number, min_sim_ratio, rounds_limit, timing, slot_cells, want_rows=True):
    n_roots = len(roots)
    n_slots = int(end_slot_of_point.max()) + 1
    sim_number_ = max(int(sim_number / n_roots), 1)                    # :285
    # per terminal cluster: the end-point slots in the order the reference scans them
    cluster_slots = [end_slot_of_point[end_point_cluster == k] for k in range(len(end_clusters_))]

    slot_counts = np.zeros(n_slots, dtype=np.int64)                   # kept sequences per end point, all rounds
    t_rounds = time.perf_counter()
    traj_num = [0]
    count = 0
    stats = []
    while min(traj_num) < traj_number:                                # :293
        print()
        print("Sampling round: {}".format(count))
        rnd = count
        count += 1
        sim_number_ *= 2                                              # :299
        # a2 + a3 + a6: K2, K3, the cross-rank first-occurrence prune; what comes back is the histogram of the
        # end points of the sequences this round kept (:341-354) -- the sequences stay on the devices as records
        st, counts = group.round(rnd, sim_number_)
        stats.append(st)
        slot_counts += counts
        traj_num = [int(slot_counts[cs].sum()) for cs in cluster_slots]   # list duplicates count twice, like :348-349

        ratio_obtained = np.round(sum(traj_num) / (sim_number_ * n_roots), 6)      # :357-358
        ratio_min = np.round(min(traj_num) / traj_number, 4)
        lagging = end_clusters_[int(np.argmin(traj_num))]
        if count == rounds_limit + 1 and min(traj_num) < min_sim_ratio * traj_number:    # :360-367
            print("{} % of required simulations obtained for lagging end point {}.".format(np.round(ratio_min * 100, 2), lagging))
            print("{} % of simulations reached atleast one endpoint after {} rounds.".format(np.round(ratio_obtained * 100, 4), rounds_limit))
            raise ValueError("Sampling failed. Try lowering min_sim_ratio or increase rounds_limit.")
        if min(traj_num) < traj_number:                               # :370-374
            print("{} % of required simulations obtained for lagging end point {}.".format(np.round(ratio_min * 100, 2), lagging))
            print("{} % of simulations reached atleast one endpoint. Generating {} additional samples.".format(np.round(ratio_obtained * 100, 4), sim_number_))
    print("Sampling done.")

    # -- final selection (:378-403).  The candidates of a terminal cluster are ordered by end point (list order), then
    # by accumulation order; the draw is np.random.randint with replacement from the global numpy RNG.  Candidate
    # idx of a cluster is the (idx - first candidate of its end point)-th kept sequence of that end point: the
    # devices resolve that against their index and replay the sequences.  With several processes rank 0 draws and
    # the others receive its numbers (their own global RNGs are not seeded alike).
    pick_slot, pick_off = [], []
    for cs in cluster_slots:
        ends = np.cumsum(slot_counts[cs])
        draw = np.empty(traj_number, dtype=np.int64)
        if group.rank == 0:
            draw[:] = np.random.randint(int(ends[-1]) if len(ends) else 0, size=traj_number)    # :395
        if group.world > 1:
            group.broadcast(draw)
        k = np.searchsorted(ends, draw, side="right")
        pick_slot.append(cs[k])
        pick_off.append(draw - (ends[k] - slot_counts[cs][k]))
    pick_slot, pick_off = np.concatenate(pick_slot), np.concatenate(pick_off)
    timing["host_ms"]["rounds_and_draw"] = 1e3 * (time.perf_counter() - t_rounds)
    adata.uns["run_info"]["simulation_steps"] = max_steps             # :406-407
    adata.uns["run_info"]["didrun"] = True
    if not want_rows:
        # a caller that only looks at where the selected sequences END (utils.undirected_simulations, utils.py:92):
        # the last cell of a pick is its end point, so no sequence has to be produced at all
        adata.uns["run_info"]["b200"] = {"rounds": stats, "world_size": group.world, "timing": timing,
                                         "final_cells": slot_cells[pick_slot].astype(np.int32),
                                         "kernel": group.graph.variant_name() if hasattr(group.graph, "variant_name") else ""}
        return adata if copy else None
    t_sel = time.perf_counter()
    cell_sequences, scores, sel = group.select(pick_slot, pick_off)
    timing["host_ms"]["select"] = 1e3 * (time.perf_counter() - t_sel)

    width = "U8" if n_roots == 1 else "U32"                           # markov_sim's own array (:39) vs the copy at :312
    adata.uns["samples"] = {
        "cell_sequences": cell_sequences,
        "transition_score": scores,
        "cluster_sequences": trunc(cell_sequences, width),
    }
    adata.uns["run_info"]["b200"] = {"rounds": stats, "world_size": group.world, "select": sel, "timing": timing,
                                     "kernel": group.graph.variant_name() if hasattr(group.graph, "variant_name") else ""}
    return adata if copy else None
