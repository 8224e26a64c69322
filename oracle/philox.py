"""Philox4x32-10 in numpy and the walker uniform stream.  TEST INFRASTRUCTURE ONLY.

The reference draws `np.random.rand(sim_number, max_steps-1)` per root cell
(cytopath/markov_chain_sampler.py:43), which is neither reproducible across worker
processes nor partition independent.  The stream below (ours, frozen here and in
include/cytopath_b200.h) replaces it; the reference is driven with the same numbers by
patching `np.random.rand` (oracle/make_golden.py).

Stream definition
-----------------
    key  = (seed & 0xffffffff, seed >> 32)
    ctr  = (walker & 0xffffffff, walker >> 32, t >> 1, round)
    x    = philox4x32_10(ctr, key)
    a, b = (x[0], x[1]) if t even else (x[2], x[3])
    u    = ((a >> 5) * 2**26 + (b >> 6)) * 2**-53          # numpy's 53-bit double, in [0, 1)

`walker` is the 0-based row of the round's sample matrix (root-major, matching
markov_chain_sampler.py:314-317: walker = j * sim_number_ + k), `t` the 0-based
transition index and `round` the 0-based sampling round (markov_chain_sampler.py:296).
Constants match Random123 / curand_philox4x32_x.h.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; all inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def walker_uniforms(seed, round_idx, walker_begin, n_walkers, n_transitions):
    """float64 [n_walkers, n_transitions] uniforms of the frozen stream."""
    seed = int(seed)
    w = np.arange(walker_begin, walker_begin + n_walkers, dtype=np.uint64)[:, None]
    t = np.arange(n_transitions, dtype=np.uint64)[None, :]
    pair = (t >> np.uint64(1)).astype(np.uint32)
    x0, x1, x2, x3 = philox4x32_10(
        (w & _MASK).astype(np.uint32), (w >> _S32).astype(np.uint32), pair, np.uint32(round_idx),
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (t & np.uint64(1)).astype(bool)
    odd = np.broadcast_to(odd, x0.shape)
    a = np.where(odd, x2, x0).astype(np.uint64)
    b = np.where(odd, x3, x1).astype(np.uint64)
    mant = (a >> np.uint64(5)) * np.uint64(67108864) + (b >> np.uint64(6))
    return mant.astype(np.float64) * (1.0 / 9007199254740992.0)
