"""TEST INFRASTRUCTURE (see oracle/__init__.py): numpy restatement of the fused-row layout that K1b builds on
the device (cytopath_b200/csrc/rows_v2.cu, layout in csrc/common.cuh).  It derives everything from the CDF of
the oracle's matrix_prep restatement (reference cytopath/markov_chain_sampler.py:11-31), so a byte-for-byte match
pins K1b to the reference's CDF: 12-bit codes, inlined targets, row words.  Never imported by the product."""
import numpy as np

HEADER = 20
OFF_BITS = 27


def used_bytes(length):
    return (HEADER + length + (length + 1) // 2 + 3) & ~3


def plan(indptr, extra=None):
    """Row words (sectors << 27 | first sector), target bases (units of 4 words), extra sectors per row."""
    lens = np.diff(indptr).astype(np.int64)
    nsect = (np.array([used_bytes(int(l)) for l in lens]) + 31) // 32
    if extra is None:
        extra = 1 if (int(nsect.sum()) + len(lens)) * 32 <= (32 << 20) else 0
    nsect = nsect + np.where((lens > 0) & (nsect + extra <= 31), extra, 0)
    first = np.concatenate([[0], np.cumsum(nsect)[:-1]])
    words = ((np.minimum(nsect, 31).astype(np.uint64) << OFF_BITS) | first.astype(np.uint64)).astype(np.uint32)   # the sector field saturates
    ebase = np.concatenate([[0], np.cumsum((lens + 3) // 4)[:-1]]).astype(np.int64)
    return words, ebase, int(nsect.sum()), int(((lens + 3) // 4).sum()), extra


def codes12(cdf_row, exact):
    if exact:
        return np.minimum(4095, np.floor(np.asarray(cdf_row, dtype=np.float64) * 4096.0)).astype(np.int64)
    return np.rint(np.asarray(cdf_row, dtype=np.float32) * np.float32(1000.0)).astype(np.int64) << 2     # 4 m, see rows_v2.cu


def build(C, cdf, exact, extra=None):
    """C: canonical CSR; cdf: per-nnz CDF of the oracle (float64 in exact mode, float32 in reference mode).
    Returns (row blocks uint8, row words uint32, target words uint32) as cyp_graph_fetch_fused delivers them."""
    indptr, indices, data = C.indptr, C.indices, np.asarray(C.data, dtype=np.float64)
    n = C.shape[0]
    words, ebase, sectors, units, extra = plan(indptr, extra)
    first_sector = (words & np.uint32((1 << OFF_BITS) - 1)).astype(np.int64)
    true_nsect = np.diff(np.concatenate([first_sector, [sectors]]))
    rows = np.zeros(sectors * 32, dtype=np.uint8)
    targets = np.zeros(units * 4, dtype=np.uint32)
    for r in range(n):
        a, b = int(indptr[r]), int(indptr[r + 1])
        length = b - a
        first = int(first_sector[r])
        nsect = int(true_nsect[r])
        blk = rows[first * 32:(first + nsect) * 32]
        used = used_bytes(length)
        m = min(length, 64, (nsect * 32 - used) // 4)
        code = codes12(cdf[a:b], exact)
        blk[HEADER:HEADER + length] = (code >> 4).astype(np.uint8)
        lo = code & 15
        lo_pad = np.concatenate([lo, [0]]) if length % 2 else lo
        blk[HEADER + length:HEADER + length + (length + 1) // 2] = (lo_pad[0::2] | (lo_pad[1::2] << 4)).astype(np.uint8)
        tw = words[indices[a:b]]
        targets[ebase[r] * 4:ebase[r] * 4 + length] = tw
        cand = data[a:min(b, a + 64)]
        order = np.lexsort((np.arange(len(cand)), -cand))[:m]          # most probable first, ties to the earlier entry
        mask = 0
        for j in order:
            mask |= 1 << int(j)
        inl = tw[np.sort(order)]
        blk[0:4] = np.frombuffer(np.int32(r).tobytes(), dtype=np.uint8)
        blk[4:6] = np.frombuffer(np.uint16(length).tobytes(), dtype=np.uint8)
        blk[6] = 0
        blk[7] = m
        blk[8:16] = np.frombuffer(np.uint64(mask).tobytes(), dtype=np.uint8)
        blk[16:20] = np.frombuffer(np.uint32(ebase[r]).tobytes(), dtype=np.uint8)
        blk[used:used + 4 * m] = np.frombuffer(inl.astype(np.uint32).tobytes(), dtype=np.uint8)
    return rows, words, targets
