"""ctypes binding of libcytopath_b200.so (include/cytopath_b200.h).

There is no CPU implementation behind this module.  If the library is missing, or no
CUDA device is present when a compute entry point is called, it raises: the product
never silently falls back to numpy or to the test oracle.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcytopath_b200.so")
if os.environ.get("CYP_LIB_PATH"):          # developer A/B runs of two builds on one GPU box
    LIB_PATH = os.environ["CYP_LIB_PATH"]

CDF_REFERENCE = 0
CDF_EXACT = 1
E_NOCROSSING = -4

EXPORTS = [
    "cyp_version", "cyp_last_error", "cyp_device_count", "cyp_release_cached_memory",
    "cyp_graph_create", "cyp_graph_create_f32", "cyp_graph_row_sum_range", "cyp_graph_canonical", "cyp_graph_destroy", "cyp_graph_info", "cyp_graph_layout", "cyp_graph_fetch_fused", "cyp_graph_fetch_cdf", "cyp_graph_build_ms", "cyp_graph_placement",
    "cyp_power_iteration", "cyp_walk", "cyp_walk_device", "cyp_walk_variant_name", "cyp_transition_scores",
    "cyp_session_create", "cyp_session_destroy", "cyp_session_round", "cyp_session_rows",
    "cyp_session_fetch_index", "cyp_session_fetch_rows", "cyp_session_slot_counts", "cyp_session_round_records",
    "cyp_session_round_prune", "cyp_session_set_hash_bits", "cyp_session_set_chunk_walkers",
    "cyp_hausdorff_matrix", "cyp_directionality_scores", "cyp_graph_power_iteration",
    "cyp_group_create_local", "cyp_group_unique_id", "cyp_group_create_rank", "cyp_group_destroy", "cyp_group_release", "cyp_group_info",
    "cyp_group_graph", "cyp_group_session", "cyp_group_graph_create", "cyp_group_graph_create_sharded", "cyp_group_session_create", "cyp_group_round",
    "cyp_group_rows", "cyp_group_fetch_index", "cyp_group_digest", "cyp_group_select", "cyp_group_set_chunk_walkers",
    "cyp_group_broadcast_bytes", "cyp_gather_ceiling",
]


class RoundStats(ctypes.Structure):
    _fields_ = [
        ("n_walkers", ctypes.c_int64), ("n_pass_clusters", ctypes.c_int64), ("n_terminal", ctypes.c_int64),
        ("n_kept", ctypes.c_int64), ("n_nocrossing", ctypes.c_int64), ("store_rows", ctypes.c_int64),
        ("dedup_passes", ctypes.c_int32), ("ms_walk", ctypes.c_float), ("ms_filter", ctypes.c_float),
        ("n_chunks", ctypes.c_int32), ("n_replayed", ctypes.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GroupRoundStats(ctypes.Structure):
    _fields_ = [
        ("n_walkers", ctypes.c_int64), ("n_pass_clusters", ctypes.c_int64), ("n_terminal", ctypes.c_int64),
        ("n_kept", ctypes.c_int64), ("n_nocrossing", ctypes.c_int64), ("n_replayed", ctypes.c_int64),
        ("total_rows", ctypes.c_int64), ("exchange_bytes", ctypes.c_int64),
        ("ms_walk", ctypes.c_float), ("ms_filter", ctypes.c_float), ("ms_local", ctypes.c_float), ("ms_exchange", ctypes.c_float),
        ("n_chunks", ctypes.c_int32), ("dedup_passes", ctypes.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GroupSelectStats(ctypes.Structure):
    _fields_ = [("gather_bytes", ctypes.c_int64), ("ms_locate_replay", ctypes.c_float), ("ms_gather", ctypes.c_float),
                ("ms_fetch", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CytopathB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cytopath_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load the shared library (building nothing: run `python -m cytopath_b200.build` or
    `__graft_entry__.build()` first).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a source checkout without the built library: compile it in tree (nvcc, sm_100a) — a build
        # step, not a fallback; if nvcc is missing this raises and so does every entry point
        try:
            from .build import build_library
            build_library()
        except Exception as exc:
            raise ImportError(
                "cytopath_b200: %s not found and building it failed (%s). Build it with "
                "`python -m cytopath_b200.build` (needs nvcc; sm_100a). There is no CPU fallback." % (LIB_PATH, exc))
    L = ctypes.CDLL(LIB_PATH)
    i32, i64, u32, u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    vp, ci = ctypes.c_void_p, ctypes.c_int
    P = ctypes.POINTER
    L.cyp_version.restype = ci
    L.cyp_last_error.restype = ctypes.c_char_p
    L.cyp_device_count.argtypes = [P(ci)]
    L.cyp_release_cached_memory.argtypes = [ci]
    L.cyp_graph_create.argtypes = [ci, i64, i64, vp, vp, vp, ci, P(vp)]
    L.cyp_graph_create_f32.argtypes = [ci, i64, i64, vp, vp, vp, ci, P(vp)]
    L.cyp_graph_row_sum_range.argtypes = [vp, P(ctypes.c_double), P(ctypes.c_double)]
    L.cyp_graph_canonical.argtypes = [vp, P(ci), P(ci)]
    L.cyp_graph_destroy.argtypes = [vp]
    L.cyp_graph_info.argtypes = [vp, P(i64), P(i64), P(i32), P(ci), P(ci), P(i64)]
    L.cyp_graph_layout.argtypes = [vp, P(ci), P(i64), P(i64), P(i32)]
    L.cyp_graph_fetch_fused.argtypes = [vp, vp, vp, vp]
    L.cyp_graph_fetch_cdf.argtypes = [vp, vp]
    L.cyp_graph_build_ms.argtypes = [vp, P(ctypes.c_float)]
    L.cyp_graph_placement.argtypes = [vp, P(ctypes.c_float), P(ctypes.c_float), P(ci)]
    L.cyp_power_iteration.argtypes = [ci, i64, i64, vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.cyp_walk.argtypes = [vp, vp, i64, i64, i64, i64, i32, u64, u32, vp, vp, P(i64)]
    L.cyp_walk_device.argtypes = [vp, vp, i64, i64, i64, i64, i32, u64, u32, vp, vp, vp, ci, vp]
    L.cyp_walk_variant_name.argtypes = [vp, ci]
    L.cyp_walk_variant_name.restype = ctypes.c_char_p
    L.cyp_transition_scores.argtypes = [vp, vp, i64, i32, vp]
    L.cyp_session_create.argtypes = [vp, vp, i64, vp, i32, i32, vp, i32, i32, ci, u64, P(vp)]
    L.cyp_session_destroy.argtypes = [vp]
    L.cyp_session_round.argtypes = [vp, u32, i64, i64, i64, P(RoundStats)]
    L.cyp_session_rows.argtypes = [vp, P(i64)]
    L.cyp_session_fetch_index.argtypes = [vp, i64, i64, vp, vp, vp]
    L.cyp_session_fetch_rows.argtypes = [vp, vp, i64, vp]
    L.cyp_session_slot_counts.argtypes = [vp, i64, i64, vp]
    L.cyp_session_round_records.argtypes = [vp, P(vp), P(i64)]
    L.cyp_session_round_prune.argtypes = [vp, vp, i64, P(i64)]
    L.cyp_session_set_hash_bits.argtypes = [vp, ci]
    L.cyp_session_set_chunk_walkers.argtypes = [vp, i64]
    L.cyp_directionality_scores.argtypes = [ci, i64, i32, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, P(ctypes.c_float)]
    L.cyp_hausdorff_matrix.argtypes = [ci, vp, i64, i32, i32, ci, vp, P(ctypes.c_float)]
    L.cyp_graph_power_iteration.argtypes = [vp, vp, vp, i32, vp, vp]
    L.cyp_group_create_local.argtypes = [ci, vp, P(vp)]
    L.cyp_group_unique_id.argtypes = [vp]
    L.cyp_group_create_rank.argtypes = [vp, ci, ci, ci, P(vp)]
    L.cyp_group_destroy.argtypes = [vp]
    L.cyp_group_release.argtypes = [vp]
    L.cyp_group_info.argtypes = [vp, P(ci), P(ci), P(ci), P(ci)]
    L.cyp_group_graph.argtypes = [vp, ci]
    L.cyp_group_graph.restype = vp
    L.cyp_group_session.argtypes = [vp, ci]
    L.cyp_group_session.restype = vp
    L.cyp_group_graph_create.argtypes = [vp, i64, i64, vp, vp, vp, ci, ci, ci, P(ctypes.c_float), P(ctypes.c_float)]
    L.cyp_group_graph_create_sharded.argtypes = [vp, i64, i64, vp, vp, vp, ci, ci, P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_float)]
    L.cyp_group_session_create.argtypes = [vp, vp, i64, vp, i32, i32, vp, i32, i32, ci, u64]
    L.cyp_group_round.argtypes = [vp, u32, i64, vp, P(GroupRoundStats)]
    L.cyp_group_rows.argtypes = [vp, P(i64)]
    L.cyp_group_fetch_index.argtypes = [vp, i64, i64, vp, vp, vp]
    L.cyp_group_digest.argtypes = [vp, vp]
    L.cyp_group_select.argtypes = [vp, i64, vp, vp, vp, vp, P(GroupSelectStats)]
    L.cyp_group_set_chunk_walkers.argtypes = [vp, i64]
    L.cyp_group_broadcast_bytes.argtypes = [vp, vp, i64, ci]
    L.cyp_gather_ceiling.argtypes = [ci, i64, i32, P(ctypes.c_double)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("cyp_last_error", "cyp_walk_variant_name", "cyp_group_graph", "cyp_group_session"):
            fn.restype = ci
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise CytopathB200Error(rc, load().cyp_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = ctypes.c_int(0)
    check(load().cyp_device_count(ctypes.byref(n)))
    return n.value


def release_cached_memory(device: int = -1) -> None:
    """Give the device memory the library keeps for reuse back to the driver (device < 0: all devices)."""
    check(load().cyp_release_cached_memory(device))


def gather_ceiling(working_set_bytes: int, gather_bytes: int = 128, device: int = 0) -> float:
    """Independent random gathers/s of `gather_bytes` out of `working_set_bytes` (cyp_gather_ceiling)."""
    out = ctypes.c_double()
    check(load().cyp_gather_ceiling(device, int(working_set_bytes), int(gather_bytes), ctypes.byref(out)))
    return out.value


def require_device(device: int = 0):
    """Fail loudly when there is no GPU: this package has no CPU path."""
    L = load()
    n = ctypes.c_int(0)
    rc = L.cyp_device_count(ctypes.byref(n))
    if rc != 0 or n.value <= device:
        raise RuntimeError(
            "cytopath_b200 needs CUDA device %d (found %d; %s). There is no CPU fallback."
            % (device, n.value, L.cyp_last_error().decode("utf-8", "replace")))


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(int(a))          # raw address (e.g. torch.Tensor.data_ptr())


def power_iteration(T, init, stationary, max_iter: int, device: int = 0, return_state: bool = False, history=None):
    """eps history of iterate_state_probability (markov_chain_sampler.py:62-87) computed on the GPU (K4).
    T: scipy sparse or dense n x n matrix as stored in adata.uns[matrix_key] (not normalised, like :73)."""
    from scipy import sparse
    require_device(device)
    M = T if sparse.isspmatrix_csr(T) or isinstance(T, sparse.csr_array) else sparse.csr_matrix(T)
    n = M.shape[0]
    col_ptr = np.ascontiguousarray(M.indptr, dtype=np.int64)          # CSR row pointers / column indices:
    row_idx = np.ascontiguousarray(M.indices, dtype=np.int32)         # the library transposes on the device
    val = np.ascontiguousarray(M.data, dtype=np.float64)
    init = np.ascontiguousarray(init, dtype=np.float64)
    stationary = np.ascontiguousarray(stationary, dtype=np.float64)
    eps = np.empty(max_iter, dtype=np.float64)
    state = np.empty(n, dtype=np.float64) if return_state else None
    if history is not None:
        assert history.shape == (max_iter, n) and history.dtype == np.float64 and history.flags["C_CONTIGUOUS"]
    check(load().cyp_power_iteration(device, n, val.shape[0], ptr(col_ptr), ptr(row_idx), ptr(val), ptr(init),
                                     ptr(stationary), max_iter, ptr(eps), ptr(state), ptr(history)))
    return (eps, state) if return_state else eps


HAUSDORFF_METRICS = {"euclidean": 0, "manhattan": 1, "chebyshev": 2, "cosine": 3}


def directionality_scores(pos, indptr, indices, w, dirs, cell, fwd, bwd, device: int = 0, return_ms: bool = False):
    """f4 kernel (cyp_directionality_scores): one score per (cell, forward direction, backward direction) instance."""
    require_device(device)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, pos.shape[1])
    cell, fwd, bwd = (np.ascontiguousarray(a, dtype=np.int32) for a in (cell, fwd, bwd))
    out = np.empty(cell.shape[0], dtype=np.float64)
    ms = ctypes.c_float(0.0)
    check(load().cyp_directionality_scores(device, pos.shape[0], pos.shape[1], ptr(pos), ptr(indptr), ptr(indices), ptr(w),
                                           dirs.shape[0], ptr(dirs), cell.shape[0], ptr(cell), ptr(fwd), ptr(bwd), ptr(out),
                                           ctypes.byref(ms)))
    return (out, ms.value) if return_ms else out


def hausdorff_matrix(coords, metric: str = "euclidean", device: int = 0, return_ms: bool = False):
    """[S, S] Hausdorff distances between the S chains of `coords` [S, L, d] (K5, cyp_hausdorff_matrix)."""
    require_device(device)
    if metric not in HAUSDORFF_METRICS:
        raise ValueError("distance must be one of %s" % sorted(HAUSDORFF_METRICS))
    c = np.ascontiguousarray(coords, dtype=np.float64)
    if c.ndim != 3:
        raise ValueError("sequence_coordinates must have shape [n_sequences, n_steps, n_dims]")
    out = np.empty((c.shape[0], c.shape[0]), dtype=np.float64)
    ms = ctypes.c_float(0)
    if c.shape[0]:
        check(load().cyp_hausdorff_matrix(device, ptr(c), c.shape[0], c.shape[1], c.shape[2], HAUSDORFF_METRICS[metric],
                                          ptr(out), ctypes.byref(ms)))
    return (out, ms.value) if return_ms else out


class Graph:
    """Device-resident transition graph: CSR + per-row CDF (K1).  Mirrors what
    `matrix_prep` (markov_chain_sampler.py:11-31) returns, minus the ELL padding."""

    _owner = True

    @classmethod
    def view(cls, handle, n: int, nnz: int, cdf_mode: int, device: int):
        """A Graph object over a handle somebody else owns (a group's graph): close() does not destroy it."""
        g = cls.__new__(cls)
        g._h, g.n, g.nnz, g.cdf_mode, g.device, g._owner, g.indptr = ctypes.c_void_p(handle), n, nnz, cdf_mode, device, False, None
        return g

    def __init__(self, C, cdf_mode: int = CDF_REFERENCE, device: int = 0):
        """C: canonical scipy CSR (sorted columns, no duplicates, no explicit zeros), float64 or float32
        (float32 values are uploaded as they are and widened on the device: no host copy)."""
        require_device(device)
        self._h = ctypes.c_void_p()
        self.indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(C.indices, dtype=np.int32)
        f32 = C.data.dtype == np.float32
        data = np.ascontiguousarray(C.data, dtype=np.float32 if f32 else np.float64)
        self.n, self.nnz, self.cdf_mode, self.device = int(C.shape[0]), int(data.shape[0]), cdf_mode, device
        create = load().cyp_graph_create_f32 if f32 else load().cyp_graph_create
        check(create(device, self.n, self.nnz, ptr(self.indptr), ptr(indices), ptr(data), cdf_mode, ctypes.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def info(self):
        n, nnz, mx, mode, dev, b = (ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int(),
                                    ctypes.c_int(), ctypes.c_int64())
        check(load().cyp_graph_info(self._h, n, nnz, mx, mode, dev, b))
        return dict(n_cells=n.value, nnz=nnz.value, max_row_nnz=mx.value, cdf_mode=mode.value,
                    device=dev.value, device_bytes=b.value)

    def layout(self):
        """Fused-row layout: dict(fused, row_block_bytes, target_word_bytes, max_block_bytes)."""
        f, rb, tb, mb = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        check(load().cyp_graph_layout(self._h, ctypes.byref(f), ctypes.byref(rb), ctypes.byref(tb), ctypes.byref(mb)))
        return dict(fused=bool(f.value), row_block_bytes=rb.value, target_word_bytes=tb.value, max_block_bytes=mb.value)

    def fused(self):
        """(row blocks uint8, row words uint32 [n], target words uint32) of the fused-row layout (test hook)."""
        lay = self.layout()
        rows = np.empty(lay["row_block_bytes"], dtype=np.uint8)
        words = np.empty(self.n, dtype=np.uint32)
        targets = np.empty(lay["target_word_bytes"] // 4, dtype=np.uint32)
        check(load().cyp_graph_fetch_fused(self._h, ptr(rows), ptr(words), ptr(targets)))
        return rows, words, targets

    def row_sum_range(self):
        lo, hi = ctypes.c_double(), ctypes.c_double()
        check(load().cyp_graph_row_sum_range(self._h, ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    def canonical(self) -> bool:
        """Did K1 find the CSR canonical (every row's columns strictly increasing, no explicit zero)?"""
        a, b = ctypes.c_int(), ctypes.c_int()
        check(load().cyp_graph_canonical(self._h, ctypes.byref(a), ctypes.byref(b)))
        return bool(a.value and b.value)

    def placement(self) -> dict:
        """Probe times of the placement calibration (0: the graph is small, not calibrated) and the number of moves."""
        a, b, m = ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
        check(load().cyp_graph_placement(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(m)))
        return {"probe_ms_before": a.value, "probe_ms_after": b.value, "moves": m.value}

    def build_ms(self) -> float:
        ms = ctypes.c_float()
        check(load().cyp_graph_build_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def cdf(self) -> np.ndarray:
        out = np.empty(self.nnz, dtype=np.float64 if self.cdf_mode == CDF_EXACT else np.float32)
        check(load().cyp_graph_fetch_cdf(self._h, ptr(out)))
        return out

    def variant_name(self, variant: int = 0) -> str:
        return load().cyp_walk_variant_name(self._h, variant).decode()

    def walk(self, root_cells, sims_per_root: int, n_transitions: int, seed: int = 0, round_idx: int = 0,
             walker_begin: int = 0, n_walkers: int | None = None, uniforms=None, strict: bool = True):
        """markov_sim for walkers [walker_begin, walker_begin + n_walkers): int32 [n_walkers, n_transitions + 1]."""
        roots = np.ascontiguousarray(root_cells, dtype=np.int32)
        if n_walkers is None:
            n_walkers = roots.shape[0] * sims_per_root - walker_begin
        out = np.empty((n_walkers, n_transitions + 1), dtype=np.int32)
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert uniforms.shape == (n_walkers, n_transitions)
        miss = ctypes.c_int64(0)
        rc = load().cyp_walk(self._h, ptr(roots), roots.shape[0], sims_per_root, walker_begin, n_walkers,
                             n_transitions, seed, round_idx, ptr(uniforms), ptr(out), ctypes.byref(miss))
        if rc == E_NOCROSSING:
            if strict:      # the reference's accidental failure mode (markov_chain_sampler.py:49)
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")
            return out
        check(rc)
        return out

    def transition_scores(self, seq) -> np.ndarray:
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        out = np.empty(seq.shape[0], dtype=np.float32)
        if seq.shape[0]:
            check(load().cyp_transition_scores(self._h, ptr(seq), seq.shape[0], seq.shape[1], ptr(out)))
        return out

    def power_iteration(self, init, stationary, max_iter: int, return_state: bool = False):
        """eps history of iterate_state_probability (:62-87) on the matrix this graph already holds in HBM (K4)."""
        init = np.ascontiguousarray(init, dtype=np.float64)
        stationary = np.ascontiguousarray(stationary, dtype=np.float64)
        eps = np.empty(max_iter, dtype=np.float64)
        state = np.empty(self.n, dtype=np.float64) if return_state else None
        check(load().cyp_graph_power_iteration(self._h, ptr(init), ptr(stationary), max_iter, ptr(eps), ptr(state)))
        return (eps, state) if return_state else eps

    def close(self):
        if self._h and self._owner:
            load().cyp_graph_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Session:
    """One `sampling()` run on one device: the per-run inputs of the round loop and the
    device-resident store of kept sequences (markov_chain_sampler.py:281-354)."""

    def __init__(self, graph: Graph, root_cells, cell_code, n_codes: int, min_clusters: int, end_slot_of_cell,
                 n_end_slots: int, max_steps: int, unique: bool, seed: int):
        self.graph = graph
        self.max_steps = int(max_steps)
        self.n_end_slots = int(n_end_slots)
        self._h = ctypes.c_void_p()
        roots = np.ascontiguousarray(root_cells, dtype=np.int32)
        code = np.ascontiguousarray(cell_code, dtype=np.int32)
        slot = np.ascontiguousarray(end_slot_of_cell, dtype=np.int32)
        assert code.shape[0] == graph.n and slot.shape[0] == graph.n
        check(load().cyp_session_create(graph.handle, ptr(roots), roots.shape[0], ptr(code), n_codes, min_clusters,
                                        ptr(slot), n_end_slots, max_steps, int(bool(unique)), seed,
                                        ctypes.byref(self._h)))

    def round(self, round_idx: int, sims_per_root: int, walker_begin: int, walker_end: int) -> dict:
        st = RoundStats()
        rc = load().cyp_session_round(self._h, round_idx, sims_per_root, walker_begin, walker_end, ctypes.byref(st))
        if rc == E_NOCROSSING:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        check(rc)
        return st.as_dict()

    def rows(self) -> int:
        n = ctypes.c_int64()
        check(load().cyp_session_rows(self._h, ctypes.byref(n)))
        return n.value

    def fetch_index(self, row_begin: int = 0, n_rows: int | None = None):
        if n_rows is None:
            n_rows = self.rows() - row_begin
        rnd = np.empty(n_rows, dtype=np.int32)
        walker = np.empty(n_rows, dtype=np.int64)
        slot = np.empty(n_rows, dtype=np.int32)
        if n_rows:
            check(load().cyp_session_fetch_index(self._h, row_begin, n_rows, ptr(rnd), ptr(walker), ptr(slot)))
        return rnd, walker, slot

    def slot_counts(self, row_begin: int = 0, n_rows: int | None = None) -> np.ndarray:
        """How many of the kept rows [row_begin, row_begin + n_rows) end on each end-point slot."""
        if n_rows is None:
            n_rows = self.rows() - row_begin
        out = np.zeros(self.n_end_slots, dtype=np.int64)
        check(load().cyp_session_slot_counts(self._h, row_begin, n_rows, ptr(out)))
        return out

    def fetch_rows(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((rows.shape[0], self.max_steps), dtype=np.int32)
        if rows.shape[0]:
            check(load().cyp_session_fetch_rows(self._h, ptr(rows), rows.shape[0], ptr(out)))
        return out

    def round_records(self):
        """(device pointer, count) of this round's (hash_lo, hash_hi, walker) uint64 triples."""
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        check(load().cyp_session_round_records(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value or 0, n.value

    def round_prune(self, d_all_records, n_all: int) -> int:
        dropped = ctypes.c_int64()
        check(load().cyp_session_round_prune(self._h, ptr(d_all_records), n_all, ctypes.byref(dropped)))
        return dropped.value

    def set_hash_bits(self, bits: int):
        check(load().cyp_session_set_hash_bits(self._h, bits))

    def set_chunk_walkers(self, walkers: int):
        check(load().cyp_session_set_chunk_walkers(self._h, walkers))

    def close(self):
        if self._h:
            load().cyp_session_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_path():
    """The NCCL the group layer should bind when the process has none loaded yet: torch's bundled copy if it is
    installed (found without importing torch), else whatever `libnccl.so.2` resolves to."""
    if os.environ.get("CYP_NCCL_PATH"):
        return os.environ["CYP_NCCL_PATH"]
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return ""


class Group:
    """The devices that share one `sampling()` run (include/cytopath_b200.h, "multi-GPU"): either N devices driven by
    this process (`Group(devices=[...])`) or this process's single device as rank `rank` of `world` processes
    (`Group(unique_id=..., rank=, world=, device=)`).  Owns the replicated graph and the per-device sessions."""

    def __init__(self, devices=None, *, unique_id: bytes | None = None, rank: int = 0, world: int = 1, device: int = 0):
        L = load()
        self._h = ctypes.c_void_p()
        self.graph = None
        if (devices is not None and len(devices) > 1) or world > 1:
            path = nccl_path()
            if path:
                os.environ.setdefault("CYP_NCCL_PATH", path)
        if unique_id is not None or world > 1:
            require_device(device)
            buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
            check(L.cyp_group_create_rank(buf, rank, world, device, ctypes.byref(self._h)))
        else:
            devs = np.ascontiguousarray([0] if devices is None else list(devices), dtype=np.int32)
            require_device(int(devs.max()))
            check(L.cyp_group_create_local(len(devs), ptr(devs), ctypes.byref(self._h)))
        w, nl, fr, ver = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(L.cyp_group_info(self._h, ctypes.byref(w), ctypes.byref(nl), ctypes.byref(fr), ctypes.byref(ver)))
        self.world, self.n_local, self.first_rank, self.nccl_version = w.value, nl.value, fr.value, ver.value
        self.n_end_slots = self.max_steps = 0
        self._cached = False

    @staticmethod
    def unique_id() -> bytes:
        path = nccl_path()
        if path:
            os.environ.setdefault("CYP_NCCL_PATH", path)
        buf = (ctypes.c_uint8 * 128)()
        check(load().cyp_group_unique_id(buf))
        return bytes(buf)

    @property
    def rank(self) -> int:
        return self.first_rank

    def create_graph(self, C, cdf_mode: int, root_rank: int = 0, shape=None, sharded: bool = False) -> dict:
        """K1 on `root_rank` from its canonical CSR `C`, then a broadcast of the device arrays.  Processes that do not
        host the root rank may pass C=None with shape=(n_cells, nnz).
        sharded=True (EVERY rank passes the same C): each rank uploads 1/world of the matrix, an all-gather over NVLink
        completes it and every device builds the graph itself (cyp_group_graph_create_sharded)."""
        L = load()
        bms, xms = ctypes.c_float(), ctypes.c_float()
        if sharded and self.world > 1:
            if C is None:
                raise ValueError("create_graph(sharded=True): every rank passes the matrix")
            indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(C.indices, dtype=np.int32)
            f32 = C.data.dtype == np.float32
            data = np.ascontiguousarray(C.data, dtype=np.float32 if f32 else np.float64)
            n, nnz = int(C.shape[0]), int(data.shape[0])
            tms = ctypes.c_float()
            check(L.cyp_group_graph_create_sharded(self._h, n, nnz, ptr(indptr), ptr(indices), ptr(data), int(f32), cdf_mode,
                                                   ctypes.byref(bms), ctypes.byref(xms), ctypes.byref(tms)))
            h = L.cyp_group_graph(self._h, 0)
            dev = ctypes.c_int()
            check(L.cyp_graph_info(h, None, None, None, None, ctypes.byref(dev), None))
            self.graph = Graph.view(h, n, nnz, cdf_mode, dev.value)
            return {"ms_upload": bms.value, "ms_gather": xms.value, "ms_total": tms.value, "sharded": True}
        if C is not None:
            indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(C.indices, dtype=np.int32)
            f32 = C.data.dtype == np.float32
            data = np.ascontiguousarray(C.data, dtype=np.float32 if f32 else np.float64)
            n, nnz = int(C.shape[0]), int(data.shape[0])
            args = (ptr(indptr), ptr(indices), ptr(data), int(f32))
        else:
            n, nnz = int(shape[0]), int(shape[1])
            args = (None, None, None, 0)
        check(L.cyp_group_graph_create(self._h, n, nnz, *args, cdf_mode, root_rank, ctypes.byref(bms), ctypes.byref(xms)))
        h = L.cyp_group_graph(self._h, 0)
        dev = ctypes.c_int()
        check(L.cyp_graph_info(h, None, None, None, None, ctypes.byref(dev), None))
        self.graph = Graph.view(h, n, nnz, cdf_mode, dev.value)
        return {"ms_build": bms.value, "ms_broadcast": xms.value}

    def create_session(self, root_cells, cell_code, n_codes: int, min_clusters: int, end_slot_of_cell, n_end_slots: int,
                       max_steps: int, unique: bool, seed: int):
        roots = np.ascontiguousarray(root_cells, dtype=np.int32)
        code = np.ascontiguousarray(cell_code, dtype=np.int32)
        slot = np.ascontiguousarray(end_slot_of_cell, dtype=np.int32)
        assert code.shape[0] == self.graph.n and slot.shape[0] == self.graph.n
        self.n_end_slots, self.max_steps = int(n_end_slots), int(max_steps)
        check(load().cyp_group_session_create(self._h, ptr(roots), roots.shape[0], ptr(code), n_codes, min_clusters, ptr(slot),
                                              n_end_slots, max_steps, int(bool(unique)), seed))

    def power_iteration_matrix(self, T, init, stationary, max_iter: int):
        """K4 on an upload of `T` (the stored matrix differs from the one the graph was built from)."""
        return power_iteration(T, init, stationary, max_iter, self.graph.device)

    def session_handle(self, local_index: int = 0):
        return load().cyp_group_session(self._h, local_index)

    def graph_view(self, local_index: int = 0) -> Graph:
        """The graph of the i-th local device (owned by the group)."""
        L = load()
        h = L.cyp_group_graph(self._h, local_index)
        if not h:
            raise CytopathB200Error(-1, "the group has no graph on local device %d" % local_index)
        dev = ctypes.c_int()
        check(L.cyp_graph_info(h, None, None, None, None, ctypes.byref(dev), None))
        return Graph.view(h, self.graph.n, self.graph.nnz, self.graph.cdf_mode, dev.value)

    def round(self, round_idx: int, sims_per_root: int, want_slot_counts: bool = True):
        """One sampling round over the whole group.  Returns (stats dict, per-slot counts of the sequences it added)."""
        st = GroupRoundStats()
        counts = np.zeros(self.n_end_slots, dtype=np.int64) if want_slot_counts else None
        rc = load().cyp_group_round(self._h, round_idx, sims_per_root, ptr(counts), ctypes.byref(st))
        if rc == E_NOCROSSING:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        check(rc)
        return st.as_dict(), counts

    def rows(self) -> int:
        n = ctypes.c_int64()
        check(load().cyp_group_rows(self._h, ctypes.byref(n)))
        return n.value

    def fetch_index(self, row_begin: int = 0, n_rows: int | None = None):
        if n_rows is None:
            n_rows = self.rows() - row_begin
        rnd, walker, slot = np.empty(n_rows, np.int32), np.empty(n_rows, np.int64), np.empty(n_rows, np.int32)
        if n_rows:
            check(load().cyp_group_fetch_index(self._h, row_begin, n_rows, ptr(rnd), ptr(walker), ptr(slot)))
        return rnd, walker, slot

    def digest(self) -> tuple:
        d = np.zeros(2, dtype=np.uint64)
        check(load().cyp_group_digest(self._h, ptr(d)))
        return int(d[0]), int(d[1])

    def select(self, slots, offsets, scores: bool = True):
        """Rows (and transition scores) of the picks: pick k = the offsets[k]-th kept sequence ending on slots[k]."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        seq = np.empty((slots.shape[0], self.max_steps), dtype=np.int32)
        score = np.empty(slots.shape[0], dtype=np.float32) if scores else None
        st = GroupSelectStats()
        if slots.shape[0]:
            check(load().cyp_group_select(self._h, slots.shape[0], ptr(slots), ptr(offsets), ptr(seq), ptr(score), ctypes.byref(st)))
        return seq, score, st.as_dict()

    def set_chunk_walkers(self, walkers: int):
        check(load().cyp_group_set_chunk_walkers(self._h, walkers))

    def broadcast(self, array: np.ndarray, root_rank: int = 0) -> np.ndarray:
        """In place: the bytes of `array` on `root_rank` reach every process of the group."""
        assert array.flags["C_CONTIGUOUS"]
        check(load().cyp_group_broadcast_bytes(self._h, ptr(array), array.nbytes, root_rank))
        return array

    def close(self):
        """Releases the device memory of the run.  A multi-GPU group keeps its NCCL communicators (forming them costs a
        few hundred ms) and stays in the module's cache for the next `sampling()` call; `destroy()` ends it."""
        if self._h:
            if self.graph is not None:
                self.graph.close()
                self.graph = None
            if self._cached:
                check(load().cyp_group_release(self._h))
            else:
                self.destroy()

    def destroy(self):
        if self._h:
            if self.graph is not None:
                self.graph.close()
                self.graph = None
            load().cyp_group_destroy(self._h)
            self._h = ctypes.c_void_p()
            for k in [k for k, v in _GROUPS.items() if v is self]:
                del _GROUPS[k]

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_GROUPS = {}          # multi-GPU groups kept between sampling() calls: (kind, ...) -> Group


def close_groups():
    """Ends the cached multi-GPU groups (their NCCL communicators)."""
    for g in list(_GROUPS.values()):
        g.destroy()


def open_group(device=None, devices=None, dist="auto") -> Group:
    """The group `sampling()` runs on: `devices=[...]` -> this process drives them all; an initialised
    torch.distributed group (one process per GPU, `dist` True or 'auto') -> one rank per process, the NCCL id travels
    through torch's store; else one device."""
    import sys
    if devices is not None:
        devs = [int(d) for d in devices]
        if not devs:
            raise ValueError("devices must name at least one CUDA device")
        if len(devs) == 1:
            return Group(devices=devs)
        key = ("local", tuple(devs))
        if key not in _GROUPS:
            _GROUPS[key] = Group(devices=devs)
            _GROUPS[key]._cached = True
        return _GROUPS[key]
    if dist is True or (dist == "auto" and "torch.distributed" in sys.modules):
        try:
            import torch.distributed as td
        except Exception:
            if dist is True:
                raise
            td = None
        if td is not None and td.is_available() and td.is_initialized() and td.get_world_size() > 1:
            rank, world = td.get_rank(), td.get_world_size()
            if device is None:
                device = int(os.environ.get("LOCAL_RANK", "0"))
            key = ("rank", rank, world, int(device))
            if key not in _GROUPS:
                box = [Group.unique_id() if rank == 0 else None]
                td.broadcast_object_list(box, src=0)
                _GROUPS[key] = Group(unique_id=box[0], rank=rank, world=world, device=int(device))
                _GROUPS[key]._cached = True
            return _GROUPS[key]
        if dist is True:
            raise RuntimeError("dist=True but torch.distributed is not initialised with more than one rank")
    return Group(devices=[0 if device is None else int(device)])
