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
import sys
import warnings

import numpy as np
from scipy import sparse
from scipy.spatial.distance import cosine

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


def iterate_state_probability(adata, matrix_key="T_forward", init=None, stationary=None, max_iter=1000, tol=1e-5):
    """Power iteration p <- p @ T with cosine distance to `stationary` (:62-87).  Returns
    (history[:converged], history, converged) like the reference."""
    T = adata.uns[matrix_key]
    history = np.empty((max_iter, adata.shape[0]))
    eps_history = np.empty(max_iter)
    p = init
    for i in range(max_iter):
        history[i] = p
        p = p @ T
        eps_history[i] = cosine(p, stationary)
    converged = check_convergence_criteria(eps_history, tol=tol)
    if isinstance(converged, int) and converged < max_iter:
        print("Tolerance reached after {} iterations of {}.".format(converged, max_iter))
    else:
        print("Max number ({}) of iterations reached.".format(max_iter))
        converged = max_iter
    return history[:converged], history, converged


def _auto_max_steps(T, init, stationary, max_iter, tol, backend=None, device=0):
    """The only thing `sampling` keeps from iterate_state_probability is its last element
    (:275-278): max_steps.  The power iteration runs on the GPU (K4, cyp_power_iteration) without
    the [max_iter, n] history buffer (1.6 GB at 1M cells); test backends without a device run the
    reference's scipy expression."""
    if backend is not None and hasattr(backend, "power_iteration"):
        eps_history = backend.power_iteration(T, init, stationary, max_iter, device)
    else:
        eps_history = np.empty(max_iter)
        p = np.asarray(init, dtype=np.float64)
        for i in range(max_iter):
            p = p @ T
            eps_history[i] = cosine(p, stationary)
    converged = check_convergence_criteria(eps_history, tol=tol)
    if isinstance(converged, int) and converged < max_iter:
        print("Tolerance reached after {} iterations of {}.".format(converged, max_iter))
        return converged
    print("Max number ({}) of iterations reached.".format(max_iter))
    return max_iter


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


_NCCL_CONNECTED = False


class _World:
    """torch.distributed plumbing for walker sharding (one process per GPU).  World size 1
    when torch.distributed is not initialised."""

    def __init__(self, use_dist):
        self.rank, self.size, self.dist, self.torch = 0, 1, None, None
        if use_dist is False:
            return
        if use_dist == "auto" and "torch.distributed" not in sys.modules:
            return                      # nobody initialised a process group: do not pay `import torch` (seconds)
        try:
            import torch
            import torch.distributed as dist
        except Exception:
            if use_dist is True:
                raise
            return
        if dist.is_available() and dist.is_initialized():
            self.rank, self.size, self.dist, self.torch = dist.get_rank(), dist.get_world_size(), dist, torch
            global _NCCL_CONNECTED
            if self.size > 1 and not _NCCL_CONNECTED and dist.get_backend() == "nccl":
                # NCCL connects its peer-to-peer transports at the first collective; if that happens after the graph
                # and the trajectory buffers are allocated, the step kernel runs ~25 % slower (measured on 2 x B200,
                # profiles/r01_multi_gpu_probe.md).  So: one tiny collective before anything is allocated.
                dist.all_reduce(torch.zeros(1, device=torch.device("cuda", torch.cuda.current_device())))
                torch.cuda.synchronize()
                _NCCL_CONNECTED = True
        elif use_dist is True:
            raise RuntimeError("dist=True but torch.distributed is not initialised")

    def all_gather_object(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        self.dist.all_gather_object(out, obj)
        return out

    def device(self, idx):
        backend = self.dist.get_backend() if self.size > 1 else "nccl"
        return self.torch.device("cuda", idx) if backend == "nccl" else self.torch.device("cpu")


class _DevPtr:
    """Zero-copy view of raw device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr, n_u64):
        self.__cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _shard(n_walkers, rank, size):
    """Contiguous block of walker ids for `rank` (lower ranks hold lower ids)."""
    return (n_walkers * rank) // size, (n_walkers * (rank + 1)) // size


# ------------------------------------------------------------------------------------ sampling
def sampling(data, auto_adjust=True, matrix_key="T_forward", cluster_key="louvain", max_steps=10, min_sim_ratio=0.6,
             rounds_limit=10, traj_number=200, sim_number=500, end_point_probability=0.99, root_cell_probability=0.99,
             end_points=None, root_cells=None, end_clusters=None, root_clusters=None, min_clusters=2, max_iter=200,
             tol=1e-3, normalize=False, unique=True, num_cores=1, copy=False, *, seed=0, cdf_mode="reference",
             device=None, dist="auto", _backend=None):
    """Markov sampling of cell sequences from root cells to terminal regions on a B200.

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
        CUDA device; defaults to LOCAL_RANK under torchrun, else 0.
    dist : 'auto' | bool
        Shard every round's walkers over the ranks of an initialised torch.distributed
        group (one process per GPU).  All ranks must call with the same arguments; all ranks
        receive the same outputs.

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

    world = _World(dist)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if world.size > 1 else 0
    backend = _backend if _backend is not None else _lib

    C = _canonical_csr(M)
    n = C.shape[0]
    if normalize:                                                          # :176-180
        row_sums = _row_sums(C)
        with np.errstate(divide="ignore", invalid="ignore"):
            C = sparse.csr_matrix((C.data.astype(np.float64, copy=False) / np.repeat(row_sums, np.diff(C.indptr)),
                                   C.indices, C.indptr), shape=C.shape)
        C.eliminate_zeros()
    # a1 (K1): upload + CDF build.  The row sums of the normalisation check (:181-187) come back from the
    # same kernel, so the host never walks the 50M-entry value array.
    graph = backend.Graph(C, _CDF_MODES[cdf_mode], device)
    session = None
    try:
        if not normalize:
            lo, hi = graph.row_sum_range() if hasattr(graph, "row_sum_range") else (lambda r: (r.min(), r.max()))(_row_sums(C))
            if not (hi <= 1.0 + 1e-3 and lo >= 1.0 - 1e-3):
                raise ValueError("Transition matrix has not been normalized. Either pass normalize==True or provide a row sum normalized matrix.")
        return _sampling_on_graph(adata, copy, world, backend, graph, C, device, auto_adjust, matrix_key, cluster_key, max_steps,
                                  min_sim_ratio, rounds_limit, traj_number, sim_number, end_point_probability,
                                  root_cell_probability, end_points, root_cells, end_clusters, root_clusters, min_clusters,
                                  max_iter, tol, unique, seed)
    finally:
        graph.close()


def _sampling_on_graph(adata, copy, world, backend, graph, C, device, auto_adjust, matrix_key, cluster_key, max_steps,
                       min_sim_ratio, rounds_limit, traj_number, sim_number, end_point_probability, root_cell_probability,
                       end_points, root_cells, end_clusters, root_clusters, min_clusters, max_iter, tol, unique, seed):
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
    clusters = np.asarray(adata.obs[cluster_key].astype(str).values, dtype=object)
    adata.uns["run_info"] = {"cluster_key": cluster_key}
    if not root_cells:
        if not root_clusters:
            try:
                root_cells = np.asarray(np.where(np.asarray(adata.obs["root_cells"]) >= root_cell_probability))[0]
            except Exception:
                raise ValueError('Root cells must be provided in adata.obs["root_cells"]. Run cytopath.terminal_states')
        else:
            root_cells = np.asarray(np.where(np.asarray(adata.obs[cluster_key].isin(root_clusters))))[0]
    adata.uns["run_info"]["root_cells"] = root_cells
    if not end_points:
        if not end_clusters:
            try:
                end_points = np.asarray(np.where(np.asarray(adata.obs["end_points"]) >= end_point_probability))[0]
            except Exception:
                raise ValueError('End points must be provided in adata.obs["end_points"]. Run cytopath.terminal_states')
        else:
            end_points = np.asarray(np.where(np.asarray(adata.obs[cluster_key].isin(end_clusters))))[0]
    adata.uns["run_info"]["end_points"] = end_points
    end_cells = np.asarray(end_points, dtype=np.int64).reshape(-1)
    end_point_clusters = clusters[end_cells].astype(str) if end_cells.size else np.empty(0, dtype=str)
    end_clusters_ = np.unique(end_point_clusters)
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
        max_steps = _auto_max_steps(adata.uns[matrix_key],
                                    (adata.obs["root_cells"] / adata.obs["root_cells"].sum()).values,
                                    (adata.obs["end_points"] / adata.obs["end_points"].sum()).values, max_iter, tol,
                                    backend, device)
        print("Number of initial simulation steps (max_steps) set to {}".format(max_steps))
    max_steps = int(max_steps)

    # -- device-side state
    trunc = np.array([c[:8] for c in clusters], dtype="U8")     # labels are stored into a U8 array (:39,:51)
    code_names, cell_code = np.unique(trunc, return_inverse=True)
    # end-point slots: one slot per distinct end-point cell; the reference scans end points in
    # list order (:348-349, :390-391), duplicates in the list count twice.
    slot_cells, end_slot_of_point = np.unique(end_cells, return_inverse=True)
    end_slot_of_cell = np.full(n, -1, dtype=np.int32)
    end_slot_of_cell[slot_cells] = np.arange(len(slot_cells), dtype=np.int32)

    session = backend.Session(graph, roots, cell_code, len(code_names), int(min_clusters), end_slot_of_cell,
                              len(slot_cells), max_steps, bool(unique), int(seed))
    try:
        return _run_rounds(adata, copy, world, graph, session, roots, end_cells, end_slot_of_point, end_point_clusters,
                           end_clusters_, trunc, max_steps, traj_number, sim_number, min_sim_ratio, rounds_limit,
                           unique, device)
    finally:
        session.close()


def _run_rounds(adata, copy, world, graph, session, roots, end_cells, end_slot_of_point, end_point_clusters,
                end_clusters_, trunc, max_steps, traj_number, sim_number, min_sim_ratio, rounds_limit, unique, device):
    n_roots = len(roots)
    n_slots = int(end_slot_of_point.max()) + 1
    sim_number_ = max(int(sim_number / n_roots), 1)                    # :285
    # per terminal cluster: the end-point slots in the order the reference scans them
    cluster_slots = [end_slot_of_point[end_point_clusters == name] for name in end_clusters_]

    # global index of the kept rows, accumulation order (:335-340): per round, per rank, walker order
    g_owner, g_local_row, g_slot = [], [], []
    slot_counts = np.zeros(n_slots, dtype=np.int64)
    traj_num = [0]
    count = 0
    stats = []
    while min(traj_num) < traj_number:                                # :293
        print()
        print("Sampling round: {}".format(count))
        rnd = count
        count += 1
        sim_number_ *= 2                                              # :299
        total = n_roots * sim_number_
        w0, w1 = _shard(total, world.rank, world.size)
        row0 = session.rows()
        st = session.round(rnd, sim_number_, w0, w1)
        if world.size > 1 and unique:
            _prune_across_ranks(world, session, device)
        _, walker, slot = session.fetch_index(row0)
        st["n_kept"] = int(len(walker))
        stats.append(st)
        round_slots = []
        for r, (r0, sl) in enumerate(world.all_gather_object((row0, slot))):   # rank order = walker order
            g_owner.append(np.full(len(sl), r, dtype=np.int32))
            g_local_row.append(np.arange(r0, r0 + len(sl), dtype=np.int64))
            g_slot.append(sl)
            round_slots.append(sl)
        round_slots = np.concatenate(round_slots)
        slot_counts += np.bincount(round_slots, minlength=n_slots)
        traj_num = [int(slot_counts[cs].sum()) for cs in cluster_slots]   # :341-354 (list duplicates count twice)

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

    # -- final selection (:378-403).  Candidates of a terminal cluster are ordered by end point
    # (list order), then by accumulation order; the draw is np.random.randint with replacement.
    owner = np.concatenate(g_owner) if g_owner else np.empty(0, np.int32)
    local_row = np.concatenate(g_local_row) if g_local_row else np.empty(0, np.int64)
    slot_all = np.concatenate(g_slot) if g_slot else np.empty(0, np.int32)
    order = np.argsort(slot_all, kind="stable")                       # accumulation order inside each slot
    starts = np.concatenate([[0], np.cumsum(np.bincount(slot_all, minlength=n_slots))])
    picked = []
    for cs in cluster_slots:
        cand = np.concatenate([order[starts[s]:starts[s + 1]] for s in cs]) if len(cs) else np.empty(0, np.int64)
        idx_range = np.random.randint(cand.shape[0], size=traj_number)    # :395
        picked.append(cand[idx_range])
    picked = np.concatenate(picked)

    cell_sequences = _collect_rows(world, session, owner[picked], local_row[picked], max_steps, device)
    adata.uns["run_info"]["simulation_steps"] = max_steps             # :406-407
    adata.uns["run_info"]["didrun"] = True
    width = "U8" if n_roots == 1 else "U32"                           # markov_sim's own array (:39) vs the copy at :312
    adata.uns["samples"] = {
        "cell_sequences": cell_sequences,
        "transition_score": graph.transition_scores(cell_sequences),
        "cluster_sequences": trunc[cell_sequences].astype(width),
    }
    adata.uns["run_info"]["b200"] = {"rounds": stats, "world_size": world.size,
                                     "kernel": graph.variant_name() if hasattr(graph, "variant_name") else ""}
    return adata if copy else None


def _prune_across_ranks(world, session, device):
    """Global first-occurrence de-duplication: all-gather the (hash, walker) records of the
    rows every rank kept this round and drop local rows that lose to a lower walker id."""
    torch, dist = world.torch, world.dist
    ptr, n_local = session.round_records()
    counts = world.all_gather_object(n_local)
    n_max = max(counts)
    if n_max == 0:
        return
    dev = world.device(device)
    if dev.type == "cuda":
        local = torch.zeros(n_max * 3, dtype=torch.int64, device=dev)
        if n_local:
            local[: n_local * 3] = torch.as_tensor(_DevPtr(ptr, n_local * 3), device=dev)
    else:                                                              # gloo test backends hand out host arrays
        local = torch.zeros(n_max * 3, dtype=torch.int64)
        if n_local:
            local[: n_local * 3] = torch.from_numpy(np.asarray(ptr).view(np.int64).reshape(-1))
    gathered = [torch.empty_like(local) for _ in range(world.size)]
    dist.all_gather(gathered, local)
    flat = torch.cat([g[: c * 3] for g, c in zip(gathered, counts)])
    n_all = int(sum(counts))
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
        session.round_prune(flat.data_ptr(), n_all)
    else:
        session.round_prune(flat.numpy().view(np.uint64), n_all)


def _collect_rows(world, session, owner, local_row, max_steps, device):
    """Every rank gathers the selected rows it owns from its device store; the pieces are
    exchanged with an all-gather so every rank ends up with the full output."""
    mine = np.flatnonzero(owner == world.rank)
    part = session.fetch_rows(local_row[mine])
    if world.size == 1:
        return part
    torch, dist = world.torch, world.dist
    dev = world.device(device)
    counts = world.all_gather_object(len(mine))
    n_max = max(max(counts), 1)
    local = torch.zeros((n_max, max_steps), dtype=torch.int32, device=dev)
    if len(mine):
        local[: len(mine)] = torch.from_numpy(part).to(dev)
    gathered = [torch.empty_like(local) for _ in range(world.size)]
    dist.all_gather(gathered, local)
    out = np.empty((len(owner), max_steps), dtype=np.int32)
    for r in range(world.size):
        pos = np.flatnonzero(owner == r)
        out[pos] = gathered[r][: counts[r]].cpu().numpy()
    return out
