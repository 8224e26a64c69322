"""A stand-in for cytopath_b200._lib.{Graph,Session,Group,open_group} built on the CPU oracle.

TEST INFRASTRUCTURE ONLY.  It lets the `-m "not gpu"` suite exercise the host logic of
`cytopath_b200.sampling` (argument resolution, round loop, multi-rank sharding over gloo,
final selection) where there is no GPU.  The product never imports this: `sampling()` uses
the CUDA library unless a test passes `_backend=` explicitly.
"""
import hashlib

import numpy as np
from scipy.spatial.distance import cosine

from oracle import cwalk, sampler_port as sp


class Graph:
    def __init__(self, C, cdf_mode, device=0):
        self.C, self.cdf_mode, self.n = C, cdf_mode, C.shape[0]
        self._cdf = cwalk.build_cdf(C, cdf_mode)

    def cdf(self):
        return self._cdf

    def canonical(self):
        return bool(self.C.has_canonical_format and np.count_nonzero(self.C.data) == self.C.nnz)

    def row_sum_range(self):
        r = np.asarray(self.C.sum(axis=1)).ravel()
        return float(r.min()), float(r.max())

    def power_iteration(self, init, stationary, max_iter):
        return power_iteration(self.C, init, stationary, max_iter)

    def transition_scores(self, seq):
        return sp.transition_scores(self.C, np.asarray(seq))

    def variant_name(self, variant=0):
        return "oracle"

    def close(self):
        pass


def power_iteration(T, init, stationary, max_iter):
    eps = np.empty(max_iter)
    p = np.asarray(init, dtype=np.float64)
    for i in range(max_iter):
        p = p @ T                                           # :73
        eps[i] = cosine(p, stationary)
    return eps


class Session:
    """Per-rank state: like the CUDA session it keeps (round, walker, slot) records, not rows."""

    def __init__(self, graph, root_cells, cell_code, n_codes, min_clusters, end_slot_of_cell, n_end_slots,
                 max_steps, unique, seed):
        self.g, self.roots = graph, np.asarray(root_cells, dtype=np.int64)
        self.code, self.min_clusters = np.asarray(cell_code), min_clusters
        self.end_slot, self.max_steps, self.unique, self.seed = np.asarray(end_slot_of_cell), max_steps, unique, seed
        self.st_round = np.empty(0, np.int32)
        self.st_walker = np.empty(0, np.int64)
        self.st_slot = np.empty(0, np.int32)
        self.records = np.empty((0, 3), np.uint64)
        self.row0 = 0
        self.sims_by_round = {}

    def replay(self, rounds, walkers):
        out = np.empty((len(walkers), self.max_steps), dtype=np.int32)
        for i, (r, w) in enumerate(zip(rounds, walkers)):
            start = self.roots[int(w) // self.sims_by_round[int(r)]]
            out[i] = cwalk.walk(self.g.C, self.g._cdf, np.array([start]), self.max_steps - 1, seed=self.seed, round_idx=int(r),
                                walker_begin=int(w))[0]
        return out

    def round(self, round_idx, sims_per_root, walker_begin, walker_end):
        W = walker_end - walker_begin
        self.sims_by_round[int(round_idx)] = int(sims_per_root)
        starts = self.roots[np.arange(walker_begin, walker_end) // sims_per_root]
        seq = cwalk.walk(self.g.C, self.g._cdf, starts, self.max_steps - 1, seed=self.seed, round_idx=round_idx,
                         walker_begin=walker_begin)
        cs = np.sort(self.code[seq], axis=1)
        distinct = 1 + (np.diff(cs, axis=1) != 0).sum(axis=1)
        ok = distinct >= self.min_clusters
        slot = self.end_slot[seq[:, -1]]
        cand = np.flatnonzero(ok & (slot >= 0))
        if self.unique and cand.size:
            first = np.unique(seq[cand], axis=0, return_index=True)[1]
            cand = cand[np.sort(first)]
        rec = np.empty((cand.size, 3), dtype=np.uint64)
        for i, w in enumerate(cand):
            d = hashlib.blake2b(seq[w].tobytes(), digest_size=16).digest()
            rec[i, 0] = int.from_bytes(d[:8], "little")
            rec[i, 1] = int.from_bytes(d[8:], "little")
            rec[i, 2] = walker_begin + w
        self.row0 = self.st_round.shape[0]
        self.st_round = np.concatenate([self.st_round, np.full(cand.size, round_idx, np.int32)])
        self.st_walker = np.concatenate([self.st_walker, (walker_begin + cand).astype(np.int64)])
        self.st_slot = np.concatenate([self.st_slot, slot[cand].astype(np.int32)])
        self.records = rec
        return dict(n_walkers=W, n_pass_clusters=int(ok.sum()), n_terminal=int((ok & (slot >= 0)).sum()),
                    n_kept=int(cand.size), n_nocrossing=0, store_rows=self.st_round.shape[0], dedup_passes=1,
                    ms_walk=0.0, ms_filter=0.0)

    def rows(self):
        return self.st_round.shape[0]

    def round_prune(self, all_records):
        best = {}
        for lo, hi, wid in all_records:
            k = (int(lo), int(hi))
            best[k] = min(best.get(k, int(wid)), int(wid))
        keep = np.array([best[(int(lo), int(hi))] == int(wid) for lo, hi, wid in self.records], dtype=bool)
        sel = np.concatenate([np.arange(self.row0), self.row0 + np.flatnonzero(keep)]).astype(np.int64)
        self.st_round, self.st_walker, self.st_slot = self.st_round[sel], self.st_walker[sel], self.st_slot[sel]
        self.records = self.records[keep]


class Group:
    """The ranks of one run: torch.distributed (gloo) moves what NCCL moves in the CUDA library."""

    def __init__(self, td=None):
        self.td = td
        self.rank = td.get_rank() if td else 0
        self.world = td.get_world_size() if td else 1
        self.graph = self.session = None
        self.g_round, self.g_walker, self.g_slot = (np.empty(0, np.int32), np.empty(0, np.int64), np.empty(0, np.int32))

    def _gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.td.all_gather_object(out, obj)
        return out

    def create_graph(self, C, cdf_mode, root_rank=0, shape=None, sharded=False):
        self.graph = Graph(C, cdf_mode)
        return {}

    def power_iteration_matrix(self, T, init, stationary, max_iter):
        return power_iteration(T, init, stationary, max_iter)

    def create_session(self, root_cells, cell_code, n_codes, min_clusters, end_slot_of_cell, n_end_slots, max_steps, unique, seed):
        self.session = Session(self.graph, root_cells, cell_code, n_codes, min_clusters, end_slot_of_cell, n_end_slots, max_steps, unique, seed)
        self.n_end_slots, self.unique, self.n_roots = n_end_slots, unique, len(root_cells)

    def round(self, round_idx, sims_per_root, want_slot_counts=True):
        total = self.n_roots * sims_per_root
        w0, w1 = (total * self.rank) // self.world, (total * (self.rank + 1)) // self.world
        s = self.session
        st = s.round(round_idx, sims_per_root, w0, w1)
        if self.world > 1 and self.unique:
            s.round_prune(np.concatenate(self._gather(s.records)))
        parts = self._gather((s.st_walker[s.row0:], s.st_slot[s.row0:]))          # rank order = walker order
        new_walker = np.concatenate([p[0] for p in parts])
        new_slot = np.concatenate([p[1] for p in parts])
        self.g_round = np.concatenate([self.g_round, np.full(len(new_slot), round_idx, np.int32)])
        self.g_walker = np.concatenate([self.g_walker, new_walker])
        self.g_slot = np.concatenate([self.g_slot, new_slot])
        st["n_kept"] = int(len(new_slot))
        st["total_rows"] = int(len(self.g_slot))
        return st, np.bincount(new_slot, minlength=self.n_end_slots).astype(np.int64)

    def rows(self):
        return len(self.g_slot)

    def select(self, slots, offsets, scores=True):
        order = np.argsort(self.g_slot, kind="stable")
        first = np.searchsorted(self.g_slot[order], slots, side="left")
        rows = order[first + np.asarray(offsets)]
        assert np.array_equal(self.g_slot[rows], slots)
        seq = self.session.replay(self.g_round[rows], self.g_walker[rows])
        return seq, (self.graph.transition_scores(seq) if scores else None), {}

    def broadcast(self, array, root_rank=0):
        if self.world > 1:
            import torch
            t = torch.from_numpy(array)
            self.td.broadcast(t, src=root_rank)
        return array

    def close(self):
        pass


def open_group(device=None, devices=None, dist="auto"):
    if dist is False:
        return Group(None)
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
            return Group(td)
    except ImportError:
        pass
    return Group(None)
