"""A stand-in for cytopath_b200._lib.{Graph,Session} built on the CPU oracle.

TEST INFRASTRUCTURE ONLY.  It lets the `-m "not gpu"` suite exercise the host logic of
`cytopath_b200.sampling` (argument resolution, round loop, multi-rank sharding over gloo,
final selection) where there is no GPU.  The product never imports this: `sampling()` uses
the CUDA library unless a test passes `_backend=` explicitly.
"""
import hashlib

import numpy as np

from oracle import cwalk, sampler_port as sp


class Graph:
    def __init__(self, C, cdf_mode, device=0):
        self.C, self.cdf_mode, self.n = C, cdf_mode, C.shape[0]
        self._cdf = cwalk.build_cdf(C, cdf_mode)

    def cdf(self):
        return self._cdf

    def transition_scores(self, seq):
        return sp.transition_scores(self.C, np.asarray(seq))

    def variant_name(self, variant=0):
        return "oracle"

    def close(self):
        pass


class Session:
    def __init__(self, graph, root_cells, cell_code, n_codes, min_clusters, end_slot_of_cell, n_end_slots,
                 max_steps, unique, seed):
        self.g, self.roots = graph, np.asarray(root_cells, dtype=np.int64)
        self.code, self.min_clusters = np.asarray(cell_code), min_clusters
        self.end_slot, self.max_steps, self.unique, self.seed = np.asarray(end_slot_of_cell), max_steps, unique, seed
        self.store = np.empty((0, max_steps), dtype=np.int32)
        self.st_round = np.empty(0, np.int32)
        self.st_walker = np.empty(0, np.int64)
        self.st_slot = np.empty(0, np.int32)
        self.records = np.empty((0, 3), np.uint64)
        self.row0 = 0

    def round(self, round_idx, sims_per_root, walker_begin, walker_end):
        W = walker_end - walker_begin
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
        self.row0 = self.store.shape[0]
        self.store = np.vstack([self.store, seq[cand]])
        self.st_round = np.concatenate([self.st_round, np.full(cand.size, round_idx, np.int32)])
        self.st_walker = np.concatenate([self.st_walker, (walker_begin + cand).astype(np.int64)])
        self.st_slot = np.concatenate([self.st_slot, slot[cand].astype(np.int32)])
        self.records = rec
        return dict(n_walkers=W, n_pass_clusters=int(ok.sum()), n_terminal=int((ok & (slot >= 0)).sum()),
                    n_kept=int(cand.size), n_nocrossing=0, store_rows=self.store.shape[0], dedup_passes=1,
                    ms_walk=0.0, ms_filter=0.0)

    def rows(self):
        return self.store.shape[0]

    def fetch_index(self, row_begin=0, n_rows=None):
        end = self.rows() if n_rows is None else row_begin + n_rows
        return self.st_round[row_begin:end], self.st_walker[row_begin:end], self.st_slot[row_begin:end]

    def fetch_rows(self, rows):
        return self.store[np.asarray(rows, dtype=np.int64)].astype(np.int32).reshape(-1, self.max_steps)

    def round_records(self):
        return self.records.reshape(-1), self.records.shape[0]

    def round_prune(self, all_records, n_all):
        allr = np.asarray(all_records).view(np.uint64).reshape(-1, 3)[:n_all]
        best = {}
        for lo, hi, wid in allr:
            k = (int(lo), int(hi))
            best[k] = min(best.get(k, int(wid)), int(wid))
        keep = np.array([best[(int(lo), int(hi))] == int(wid) for lo, hi, wid in self.records], dtype=bool)
        m = self.records.shape[0]
        sel = np.concatenate([np.arange(self.row0), self.row0 + np.flatnonzero(keep)]).astype(np.int64)
        self.store, self.st_round = self.store[sel], self.st_round[sel]
        self.st_walker, self.st_slot = self.st_walker[sel], self.st_slot[sel]
        self.records = self.records[keep]
        return int(m - keep.sum())

    def set_hash_bits(self, bits):
        pass

    def close(self):
        pass
