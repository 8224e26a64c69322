// a3: one sampling round and its filters (reference cytopath/markov_chain_sampler.py:299-354).
//
// Round 2 design: KEPT SEQUENCES ARE NOT STORED.  The uniform stream is a counter-based Philox keyed by
// (seed; walker, step, round), so walking a walker again reproduces its trajectory bit for bit.  A kept
// sequence is therefore 16 bytes -- (round, walker id, end-point slot) -- plus, while its round is being
// de-duplicated, its 128-bit hash.  Rows are materialised by REPLAY only where the reference's semantics
// need them: to verify that two sequences with the same hash are equal (np.unique(axis=0), :329-333) and
// for the rows the final draw selects (:395-403).  What this buys on a B200: the step kernel of a round
// writes no trajectories (4 bytes of HBM stores per walker-step gone, a 100M-walker round needs 2.4 GB
// instead of 80 GB), nothing of size [walkers, max_steps] is ever allocated, and a multi-GPU round
// exchanges 24-byte records instead of 2 KB rows.
//
// K2 (walk*.cu, filter mode, no trajectory buffer) leaves per walker a 128-bit hash of its sequence and a
// slot code (-2: fewer than min_clusters labels, :322; -1: does not end on an end point; >= 0: index of
// the end-point cell it ends on, :343-349).  K3, in this file:
//   compact_*        order-preserving stream compaction (warp ballots + block offsets) of the walkers that
//                    passed both filters ("candidates"), in walker order;
//   dedup_insert     open-addressing hash table keyed by the sequence hash; every item does atomicMin(its
//                    position) on its key's slot, so the winner is the first occurrence whatever the
//                    thread schedule (np.unique(..., return_index) + sorted, :330-331);
//   dedup_classify / compare_pairs
//                    a non-winner is dropped only after its REPLAYED row compared equal to the winner's
//                    replayed row; a true hash collision marks it unresolved, and the items that share an
//                    unresolved item's key are replayed, re-hashed with another salt and resolved among
//                    themselves until none is left, so the result is exact for any hash quality;
//   append_kept      survivors are appended, in walker order, to the session's index of kept sequences.
// The same de-duplication runs over the records of all chunks of a round and over the records gathered
// from all ranks (cyp_session_round_prune): a rank can replay any walker of any rank.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "session.cuh"

namespace cyp {

constexpr int kThreads = 256;
constexpr int kItems = 4;                  // elements per thread in the compaction kernels
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ uint64_t fmix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// ---- order-preserving compaction: out[rank(i)] = i for every i with pred(i) ---------------
struct PredSlotGE {            // slot[i] >= threshold
    const int32_t *slot;
    int32_t thr;
    __device__ bool operator()(int64_t i) const { return slot[i] >= thr; }
};
struct PredFlagEq {            // flag[i] == value
    const uint8_t *flag;
    uint8_t value;
    __device__ bool operator()(int64_t i) const { return flag[i] == value; }
};
struct PredSlotFlagged {       // list position i: the table slot of its key is flagged, and it is not an own item already decided
    const uint64_t *slot_of;
    const uint8_t *slot_flag;
    const int64_t *ids;        // original item of list position i (nullptr: identity)
    const uint8_t *keep;       // by original item
    int64_t own_begin, own_end;
    __device__ bool operator()(int64_t i) const {
        if (slot_of[i] == ~0ull || !slot_flag[slot_of[i]]) return false;
        const int64_t o = ids ? ids[i] : i;
        return !(o >= own_begin && o < own_end && keep[o] == 0);       // verified duplicates are gone for good
    }
};

template <typename Pred>
__global__ void __launch_bounds__(kThreads) compact_count_kernel(int64_t n, Pred pred, int32_t *__restrict__ tile_count) {
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kTile;
    int c = 0;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = base + it * kThreads + threadIdx.x;
        c += (i < n && pred(i)) ? 1 : 0;
    }
    __shared__ int s[kThreads / 32];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += s[w];
        tile_count[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of tile counts into 64-bit offsets, total in offsets[n_tiles]
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int64_t n_tiles, const int32_t *__restrict__ tile_count,
                                                          int64_t *__restrict__ offsets) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = (i < n_tiles) ? tile_count[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t wv = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, wv, d);
                if (lane >= d) wv += y;
            }
            s_warp[lane] = wv;                 // inclusive over warps
        }
        __syncthreads();
        const int64_t warp_excl = (warp == 0) ? 0 : s_warp[warp - 1];
        const int64_t carry = s_carry;
        if (i < n_tiles) offsets[i] = carry + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_tiles] = s_carry;
}

template <typename Pred>
__global__ void __launch_bounds__(kThreads) compact_write_kernel(int64_t n, Pred pred, const int64_t *__restrict__ offsets,
                                                                 int64_t *__restrict__ out) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_run;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    const int64_t tile_off = offsets[blockIdx.x];
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = base + it * kThreads + threadIdx.x;
        const bool f = (i < n) && pred(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (f) out[tile_off + before + __popc(b & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < kThreads / 32; ++w) t += s_warp[w];
            s_run += t;
        }
        __syncthreads();
    }
}

// ---- records, replay items --------------------------------------------------------------------
// record of candidate c (walker-local index cand[c] of the chunk): (hash_lo, hash_hi, global walker id)
__global__ void __launch_bounds__(kThreads)
make_records_kernel(int64_t n_cand, const int64_t *__restrict__ cand, const uint64_t *__restrict__ hash, int64_t walker_begin,
                    uint64_t *__restrict__ rec) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (c >= n_cand) return;
    const int64_t w = cand[c];
    rec[3 * c] = hash[2 * w];
    rec[3 * c + 1] = hash[2 * w + 1];
    rec[3 * c + 2] = static_cast<uint64_t>(walker_begin + w);
}

// replay item of list position k: item ids[k] (or k), whose walker id is walker[stride * item] and whose round is
// round_of[item] (or `round`); root = roots[walker / sims_per_root of that round]
__global__ void __launch_bounds__(kThreads)
make_items_kernel(int64_t n, const int64_t *__restrict__ ids, const uint64_t *__restrict__ walker, int walker_stride,
                  const int32_t *__restrict__ round_of, uint32_t round, const int64_t *__restrict__ sims_by_round,
                  const int32_t *__restrict__ roots, ReplayItem *__restrict__ items) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n) return;
    const int64_t o = ids ? ids[k] : k;
    const int64_t w = static_cast<int64_t>(walker[static_cast<int64_t>(walker_stride) * o]);
    const uint32_t r = round_of ? static_cast<uint32_t>(round_of[o]) : round;
    items[k] = ReplayItem{w, r, roots[w / sims_by_round[r]]};
}

// ---- de-duplication ---------------------------------------------------------------------
struct DedupTable {
    uint64_t *key_lo;     // 0 = empty
    uint64_t *key_hi;     // 0 = not yet written
    uint64_t *min_pos;    // lowest list position with this key
    uint64_t mask;        // capacity - 1
};

// List position k holds the hash hash[stride * k], hash[stride * k + 1] (pass 0: the records, stride 3; later passes:
// the salted re-hash of the replayed rows, stride 2).  Positions ascend with the walker id, so the lowest position
// with a key is its first occurrence.
__global__ void __launch_bounds__(kThreads)
dedup_insert_kernel(int64_t first, int64_t count, const uint64_t *__restrict__ hash, int hash_stride, uint64_t lo_mask, uint64_t hi_mask,
                    DedupTable t, uint64_t *__restrict__ slot_of) {
    const int64_t k = first + static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= first + count) return;
    const uint64_t lo = (hash[static_cast<int64_t>(hash_stride) * k] & lo_mask) | 1ull;          // never 0
    const uint64_t hi = (hash[static_cast<int64_t>(hash_stride) * k + 1] & hi_mask) | 1ull;
    uint64_t idx = fmix(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & t.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_lo + idx), 0ull, lo);
        if (prev == 0ull || prev == lo) {
            const unsigned long long ph = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_hi + idx), 0ull, hi);
            if (ph == 0ull || ph == hi) {
                atomicMin(reinterpret_cast<unsigned long long *>(t.min_pos + idx), static_cast<unsigned long long>(k));
                slot_of[k] = idx;
                return;
            }
        }
        idx = (idx + 1) & t.mask;
    }
}

// Positions [0, count) that are NOT in the table's owner range (records of lower ranks): a position whose key is in the table
// competes for it (atomicMin) and remembers the slot; the others get no slot.  Read-only but for the matches, and the table
// only holds this rank's keys, so it stays cache resident: a rank does not insert the records of all ranks to find its own losers.
__global__ void __launch_bounds__(kThreads)
dedup_probe_kernel(int64_t count, const uint64_t *__restrict__ hash, int hash_stride, uint64_t lo_mask, uint64_t hi_mask, DedupTable t,
                   uint64_t *__restrict__ slot_of) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= count) return;
    const uint64_t lo = (hash[static_cast<int64_t>(hash_stride) * k] & lo_mask) | 1ull;
    const uint64_t hi = (hash[static_cast<int64_t>(hash_stride) * k + 1] & hi_mask) | 1ull;
    uint64_t idx = fmix(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & t.mask;
    while (true) {
        const uint64_t kl = t.key_lo[idx];
        if (kl == 0ull) { slot_of[k] = ~0ull; return; }
        if (kl == lo && t.key_hi[idx] == hi) {
            atomicMin(reinterpret_cast<unsigned long long *>(t.min_pos + idx), static_cast<unsigned long long>(k));
            slot_of[k] = idx;
            return;
        }
        idx = (idx + 1) & t.mask;
    }
}

// keep[item]: 1 = first occurrence, 0 = verified duplicate, 2 = unresolved (hash collision), 3 = has a lower
// position with the same key (winner_of[item]) and waits for the row compare.  Only this rank's own items
// [own_begin, own_end) are classified; in later passes only those still unresolved.
__global__ void __launch_bounds__(kThreads)
dedup_classify_kernel(int64_t n_list, const int64_t *__restrict__ ids, int64_t own_begin, int64_t own_end, int pass, DedupTable t,
                      const uint64_t *__restrict__ slot_of, uint8_t *__restrict__ keep, int64_t *__restrict__ winner_of) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_list) return;
    const int64_t o = ids ? ids[k] : k;
    if (o < own_begin || o >= own_end) return;
    if (pass > 0 && keep[o] != 2) return;
    const int64_t wpos = static_cast<int64_t>(t.min_pos[slot_of[k]]);
    if (wpos == k) {
        keep[o] = 1;
    } else {
        keep[o] = 3;
        winner_of[o] = ids ? ids[wpos] : wpos;
    }
}

// replay list of a batch of m pending items: positions 0..m-1 the items, m..2m-1 their winners
__global__ void __launch_bounds__(kThreads)
make_pair_ids_kernel(int64_t m, const int64_t *__restrict__ pending, const int64_t *__restrict__ winner_of, int64_t *__restrict__ ids) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= m) return;
    ids[k] = pending[k];
    ids[m + k] = winner_of[pending[k]];
}

// one warp per pending item: rows k and m + k of the replay buffer
__global__ void __launch_bounds__(kThreads)
compare_pairs_kernel(int64_t m, const int64_t *__restrict__ pending, const int32_t *__restrict__ rows, int32_t n_cols,
                     uint8_t *__restrict__ keep, unsigned long long *__restrict__ n_unresolved) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= m) return;
    const int32_t *a = rows + k * static_cast<int64_t>(n_cols);
    const int32_t *b = rows + (m + k) * static_cast<int64_t>(n_cols);
    bool diff = false;
    for (int i = lane; i < n_cols; i += 32) diff |= (a[i] != b[i]);
    diff = __any_sync(0xffffffffu, diff);
    if (lane == 0) {
        keep[pending[k]] = diff ? 2 : 0;
        if (diff) atomicAdd(n_unresolved, 1ull);
    }
}

// flags the table slots that hold the key of an unresolved own item
__global__ void __launch_bounds__(kThreads)
mark_slots_kernel(int64_t n_list, const int64_t *__restrict__ ids, int64_t own_begin, int64_t own_end, const uint8_t *__restrict__ keep,
                  const uint64_t *__restrict__ slot_of, uint8_t *__restrict__ slot_flag) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_list) return;
    const int64_t o = ids ? ids[k] : k;
    if (o >= own_begin && o < own_end && keep[o] == 2 && slot_of[k] != ~0ull) slot_flag[slot_of[k]] = 1;
}

// out[k] = ids ? ids[sel[k]] : sel[k]  (list positions -> original items)
__global__ void __launch_bounds__(kThreads)
map_ids_kernel(int64_t n, const int64_t *__restrict__ sel, const int64_t *__restrict__ ids, int64_t *__restrict__ out) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k < n) out[k] = ids ? ids[sel[k]] : sel[k];
}

// salted 128-bit hash of replayed rows (passes after the first)
__global__ void __launch_bounds__(kThreads)
rehash_rows_kernel(int64_t n, const int32_t *__restrict__ rows, int32_t n_cols, uint64_t salt, uint64_t *__restrict__ hash) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n) return;
    const int32_t *row = rows + k * static_cast<int64_t>(n_cols);
    uint64_t h1 = fmix(salt ^ 0x9E3779B97F4A7C15ull), h2 = fmix(salt + 0xD6E8FEB86659FD93ull);
    for (int i = 0; i < n_cols; ++i) {
        const uint64_t x = static_cast<uint32_t>(row[i]);
        h1 = fmix(h1 ^ (x + salt));
        h2 = (h2 ^ x) * 0x9FB21C651E98DF25ull + salt;
        h2 ^= h2 >> 31;
    }
    hash[2 * k] = h1;
    hash[2 * k + 1] = fmix(h2);
}

// ---- counting, appending -------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
count_slots_kernel(int64_t n, const int32_t *__restrict__ slot, unsigned long long *__restrict__ counts) {
    // counts[0] = #(slot >= -1) (passed min_clusters), counts[1] = #(slot >= 0) (and terminal)
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int32_t s = (i < n) ? slot[i] : -3;
    const unsigned b1 = __ballot_sync(0xffffffffu, s >= -1);
    const unsigned b2 = __ballot_sync(0xffffffffu, s >= 0);
    if ((threadIdx.x & 31) == 0) {
        if (b1) atomicAdd(counts, static_cast<unsigned long long>(__popc(b1)));
        if (b2) atomicAdd(counts + 1, static_cast<unsigned long long>(__popc(b2)));
    }
}

// kept candidate k (rank kept[k] or k among the chunk's candidates) -> index record + round record
__global__ void __launch_bounds__(kThreads)
append_kept_kernel(int64_t n_kept, const int64_t *__restrict__ kept, const int64_t *__restrict__ cand, const int32_t *__restrict__ slot,
                   const uint64_t *__restrict__ cand_rec, int32_t round, int32_t *__restrict__ k_round, int64_t *__restrict__ k_walker,
                   int32_t *__restrict__ k_slot, uint64_t *__restrict__ rec_out) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_kept) return;
    const int64_t c = kept ? kept[k] : k;
    k_round[k] = round;
    k_walker[k] = static_cast<int64_t>(cand_rec[3 * c + 2]);
    k_slot[k] = slot[cand[c]];
    rec_out[3 * k] = cand_rec[3 * c];
    rec_out[3 * k + 1] = cand_rec[3 * c + 1];
    rec_out[3 * k + 2] = cand_rec[3 * c + 2];
}

// order-preserving compaction of the last round's segment of the index and of its records (out of place)
__global__ void __launch_bounds__(kThreads)
move_index_kernel(int64_t n_keep, const int64_t *__restrict__ src, const int32_t *__restrict__ r_in, const int64_t *__restrict__ w_in,
                  const int32_t *__restrict__ s_in, const uint64_t *__restrict__ rec_in, int32_t *__restrict__ r_out,
                  int64_t *__restrict__ w_out, int32_t *__restrict__ s_out, uint64_t *__restrict__ rec_out) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_keep) return;
    const int64_t s = src[k];
    r_out[k] = r_in[s];
    w_out[k] = w_in[s];
    s_out[k] = s_in[s];
    rec_out[3 * k] = rec_in[3 * s];
    rec_out[3 * k + 1] = rec_in[3 * s + 1];
    rec_out[3 * k + 2] = rec_in[3 * s + 2];
}

// position of this rank's first record (walker id `first_walker`) in the records gathered from all ranks
__global__ void __launch_bounds__(kThreads)
find_walker_kernel(int64_t n, const uint64_t *__restrict__ rec, uint64_t first_walker, long long *__restrict__ pos) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k < n && rec[3 * k + 2] == first_walker) *pos = k;
}

__global__ void __launch_bounds__(kThreads)
slot_histogram_kernel(int64_t n, const int32_t *__restrict__ slot, unsigned long long *__restrict__ counts) {
    // walkers pile up on few end points: one atomic per distinct slot of a warp, not per row (a hot counter serialises)
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    const bool in = i < n;
    const unsigned active = __ballot_sync(0xffffffffu, in);
    if (!in) return;
    const int32_t s = slot[i];
    const unsigned same = __match_any_sync(active, s);
    if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(counts + s, static_cast<unsigned long long>(__popc(same)));
}

// the label code lives in byte 6 of every fused row block's header (walk_v2.cuh reads it with the header)
__global__ void __launch_bounds__(256)
label_rows2_kernel(int64_t n, const uint32_t *__restrict__ word, unsigned char *__restrict__ rows2, const uint8_t *__restrict__ code8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) rows2[static_cast<size_t>(word[i] & kWordOffMask) * 32 + 6] = code8[i];
}

// generic min_clusters filter for label sets too large for the in-kernel bitmap (n_codes > 256):
// one warp per walker, bitmap of n_codes bits in shared memory.  The only path that needs the trajectories of a
// whole chunk: K2 writes them for it.
__global__ void __launch_bounds__(128)
distinct_codes_kernel(int64_t n_walkers, const int32_t *__restrict__ seq, int64_t ld, int32_t n_cols,
                      const int32_t *__restrict__ code, int32_t n_codes, int32_t min_clusters, int32_t *__restrict__ slot) {
    extern __shared__ unsigned s_bits[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words = (n_codes + 31) / 32;
    unsigned *bits = s_bits + warp * words;
    for (int64_t w = static_cast<int64_t>(blockIdx.x) * 4 + warp; w < n_walkers; w += static_cast<int64_t>(gridDim.x) * 4) {
        for (int i = lane; i < words; i += 32) bits[i] = 0u;
        __syncwarp();
        const int32_t *row = seq + w * ld;
        for (int i = lane; i < n_cols; i += 32) {
            const int c = code[row[i]];
            atomicOr(bits + (c >> 5), 1u << (c & 31));
        }
        __syncwarp();
        int cnt = 0;
        for (int i = lane; i < words; i += 32) cnt += __popc(bits[i]);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt < min_clusters) slot[w] = -2;
        __syncwarp();
    }
}

// writes the session's label code of every edge's target into the edge records, so the step
// loop of walk_tma.cu / walk.cu gets the label with the 16 bytes it loads anyway
__global__ void __launch_bounds__(kThreads)
label_edges_kernel(int64_t n_edges, EdgeRec *__restrict__ edge, const uint8_t *__restrict__ code8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n_edges) return;
    const int32_t c = edge[i].col;
    edge[i].code = (c >= 0) ? code8[c] : 0u;
}

}  // namespace cyp

using namespace cyp;

namespace {

struct Scratch {                   // frees everything it allocated when it goes out of scope
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) dev_free(p);
    }
    template <typename T>
    cudaError_t alloc(T **p, size_t count) {
        *p = nullptr;
        cudaError_t e = dev_alloc(reinterpret_cast<void **>(p), std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

#define S_CUDA(expr)                                                                                           \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            return fail(_e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA,                            \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                                   \
    } while (0)

inline unsigned blocks_for(int64_t n, int per_block) { return static_cast<unsigned>((n + per_block - 1) / per_block); }

// out[0..count) = indices i in [0, n) with pred(i), ascending.  Returns count through *count.
template <typename Pred>
int compact_indices(int64_t n, Pred pred, Scratch &sc, int64_t **d_out, int64_t *count, cudaStream_t st) {
    *count = 0;
    *d_out = nullptr;
    if (n == 0) return CYP_OK;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    int32_t *d_tile = nullptr;
    int64_t *d_off = nullptr;
    S_CUDA(sc.alloc(&d_tile, n_tiles));
    S_CUDA(sc.alloc(&d_off, n_tiles + 1));
    compact_count_kernel<<<static_cast<unsigned>(n_tiles), kThreads, 0, st>>>(n, pred, d_tile);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(n_tiles, d_tile, d_off);
    S_CUDA(cudaGetLastError());
    int64_t total = 0;
    S_CUDA(cudaMemcpyAsync(&total, d_off + n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    S_CUDA(sc.alloc(d_out, total));
    if (total) {
        compact_write_kernel<<<static_cast<unsigned>(n_tiles), kThreads, 0, st>>>(n, pred, d_off, *d_out);
        S_CUDA(cudaGetLastError());
    }
    *count = total;
    return CYP_OK;
}

int alloc_table(int64_t n_keys, Scratch &sc, DedupTable *t, cudaStream_t st) {
    uint64_t cap = 1024;
    while (cap < static_cast<uint64_t>(n_keys) * 2) cap <<= 1;
    S_CUDA(sc.alloc(&t->key_lo, cap));
    S_CUDA(sc.alloc(&t->key_hi, cap));
    S_CUDA(sc.alloc(&t->min_pos, cap));
    S_CUDA(cudaMemsetAsync(t->key_lo, 0, cap * 8, st));
    S_CUDA(cudaMemsetAsync(t->key_hi, 0, cap * 8, st));
    S_CUDA(cudaMemsetAsync(t->min_pos, 0xFF, cap * 8, st));
    t->mask = cap - 1;
    return CYP_OK;
}

void hash_masks(int bits, uint64_t *lo, uint64_t *hi) {
    if (bits >= 128) { *lo = ~0ull; *hi = ~0ull; }
    else if (bits >= 64) { *lo = ~0ull; *hi = (bits == 64) ? 0ull : ((1ull << (bits - 64)) - 1ull); }
    else { *lo = (bits <= 0) ? 0ull : ((1ull << bits) - 1ull); *hi = 0ull; }
}

size_t free_device_bytes(int device) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    // memory the stream-ordered pool holds but is not using is available too
    cudaMemPool_t pool;
    uint64_t reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        free_b += static_cast<size_t>(reserved - used);
    return free_b + dev_cached_bytes(device);        // dev_alloc gives cached blocks back when the pool runs dry
}

// CYP_TRACE=1: host wall-clock between marks, to stderr (developer aid)
struct Tracer {
    bool on;
    std::chrono::steady_clock::time_point t;
    Tracer() : on(getenv("CYP_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cyp trace] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

int grow_index(cyp_session *s, int64_t need_rows) {
    if (need_rows <= s->kept_cap) return CYP_OK;
    int64_t cap = std::max<int64_t>(need_rows, s->kept_cap + s->kept_cap / 2);
    cap = std::max<int64_t>(cap, 4096);
    int32_t *n_round = nullptr, *n_slot = nullptr;
    int64_t *n_walker = nullptr;
    cudaError_t e = dev_alloc(&n_round, sizeof(int32_t) * cap);
    if (e == cudaSuccess) e = dev_alloc(&n_walker, sizeof(int64_t) * cap);
    if (e == cudaSuccess) e = dev_alloc(&n_slot, sizeof(int32_t) * cap);
    if (e == cudaSuccess && s->kept_rows) {
        e = cudaMemcpy(n_round, s->d_k_round, sizeof(int32_t) * s->kept_rows, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(n_walker, s->d_k_walker, sizeof(int64_t) * s->kept_rows, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(n_slot, s->d_k_slot, sizeof(int32_t) * s->kept_rows, cudaMemcpyDeviceToDevice);
    }
    if (e != cudaSuccess) {
        dev_free(n_round); dev_free(n_walker); dev_free(n_slot);
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("grow_index: ") + cudaGetErrorString(e));
    }
    dev_free(s->d_k_round); dev_free(s->d_k_walker); dev_free(s->d_k_slot);
    s->d_k_round = n_round; s->d_k_walker = n_walker; s->d_k_slot = n_slot;
    s->kept_cap = cap;
    return CYP_OK;
}

int ensure_records(cyp_session *s, int64_t need) {
    if (need <= s->records_cap) return CYP_OK;
    const int64_t cap = std::max<int64_t>(std::max<int64_t>(need, 4096), s->records_cap + s->records_cap / 2);
    uint64_t *n_rec = nullptr;
    S_CUDA(dev_alloc(&n_rec, sizeof(uint64_t) * 3 * cap));
    if (s->n_records) S_CUDA(cudaMemcpy(n_rec, s->d_records, sizeof(uint64_t) * 3 * s->n_records, cudaMemcpyDeviceToDevice));
    dev_free(s->d_records);
    s->d_records = n_rec;
    s->records_cap = cap;
    return CYP_OK;
}

int note_round(cyp_session *s, uint32_t round, int64_t sims_per_root) {
    if (round >= (1u << 20)) return fail(CYP_E_INVALID, "cyp_session_round: round index out of range");
    if (s->sims_by_round.size() <= round) s->sims_by_round.resize(round + 1, 1);
    if (s->sims_by_round[round] == sims_per_root && s->d_sims_by_round && s->sims_uploaded > round) return CYP_OK;
    s->sims_by_round[round] = sims_per_root;
    dev_free(s->d_sims_by_round);
    s->d_sims_by_round = nullptr;
    S_CUDA(dev_alloc(&s->d_sims_by_round, sizeof(int64_t) * s->sims_by_round.size()));
    S_CUDA(cudaMemcpy(s->d_sims_by_round, s->sims_by_round.data(), sizeof(int64_t) * s->sims_by_round.size(), cudaMemcpyHostToDevice));
    s->sims_uploaded = s->sims_by_round.size();
    return CYP_OK;
}

}  // namespace

namespace cyp {

// Rows of n replay items into d_rows [n, max_steps] (plain mode of the graph's default step kernel).
int session_replay(const cyp_session *s, int64_t n, const ReplayItem *d_items, int32_t *d_rows, cudaStream_t st) {
    if (n == 0) return CYP_OK;
    WalkParams p{};
    p.rows = s->g->d_row; p.cdf = s->g->d_cdf; p.edge = s->g->d_edge;
    p.roots = s->d_roots; p.sims_per_root = 1;
    p.walker_begin = 0; p.n_walkers = n; p.nt = s->max_steps - 1; p.round = 0; p.seed = s->seed;
    p.uniforms = nullptr; p.replay = d_items; p.out_seq = d_rows; p.ld = s->max_steps; p.nocross = nullptr;
    return launch_walk(s->g, p, kModePlain, 0, st);
}

int64_t session_replay_batch_rows(const cyp_session *s) {
    // rows per replay batch: at most 1 GB of trajectories, at most a quarter of the free HBM
    const size_t row_bytes = sizeof(int32_t) * static_cast<size_t>(s->max_steps);
    size_t budget = std::min<size_t>(1ull << 30, free_device_bytes(s->device) / 4);
    int64_t rows = static_cast<int64_t>(budget / row_bytes);
    return rows < 64 ? 64 : rows;
}

// First-occurrence de-duplication of n items given as records (hash_lo, hash_hi, walker id), ascending in the walker
// id, all of round `round`.  Only the items [own_begin, own_end) are decided: keep[item] = 1 (first occurrence) or 0
// (an equal sequence has a lower walker id somewhere in the list).  Exact for any hash quality (see the file header).
static int session_dedup(cyp_session *s, Scratch &sc, int64_t n, const uint64_t *d_rec, int64_t own_begin, int64_t own_end, uint32_t round,
                  cudaStream_t st, uint8_t **d_keep_out, int *passes_out, int64_t *replayed_out) {
    uint8_t *d_keep = nullptr;
    int64_t *d_winner = nullptr;
    unsigned long long *d_unres = nullptr;
    S_CUDA(sc.alloc(&d_keep, n));
    S_CUDA(sc.alloc(&d_winner, n));
    S_CUDA(sc.alloc(&d_unres, 1));
    S_CUDA(cudaMemsetAsync(d_keep, 1, n, st));
    const int32_t ncol = s->max_steps;
    const int64_t batch_rows = session_replay_batch_rows(s);
    uint64_t lo_mask, hi_mask;
    hash_masks(s->hash_bits, &lo_mask, &hi_mask);
    // the list of the current pass: positions -> items (nullptr: all items), hashes by position
    int64_t n_list = n;
    int64_t *d_ids = nullptr;
    const uint64_t *d_hash = d_rec;
    int hash_stride = 3;
    int passes = 0;
    int64_t replayed = 0;
    while (true) {
        Scratch ps;                                   // per pass
        DedupTable tab{};
        uint64_t *d_slot_of = nullptr;
        // first pass over records gathered from several ranks: the table holds this rank's keys, the records of the lower ranks
        // only probe it (those of the higher ranks cannot win against an own record at all)
        const bool probe = passes == 0 && d_ids == nullptr && (own_begin > 0 || own_end < n_list);
        int rc = alloc_table(probe ? own_end - own_begin : n_list, ps, &tab, st);
        if (rc != CYP_OK) return rc;
        S_CUDA(ps.alloc(&d_slot_of, n_list));
        if (probe) {
            S_CUDA(cudaMemsetAsync(d_slot_of, 0xFF, sizeof(uint64_t) * n_list, st));
            dedup_insert_kernel<<<blocks_for(own_end - own_begin, kThreads), kThreads, 0, st>>>(own_begin, own_end - own_begin, d_hash, hash_stride, lo_mask,
                                                                                                hi_mask, tab, d_slot_of);
            if (own_begin > 0)
                dedup_probe_kernel<<<blocks_for(own_begin, kThreads), kThreads, 0, st>>>(own_begin, d_hash, hash_stride, lo_mask, hi_mask, tab, d_slot_of);
        } else {
            dedup_insert_kernel<<<blocks_for(n_list, kThreads), kThreads, 0, st>>>(0, n_list, d_hash, hash_stride, lo_mask, hi_mask, tab, d_slot_of);
        }
        dedup_classify_kernel<<<blocks_for(n_list, kThreads), kThreads, 0, st>>>(n_list, d_ids, own_begin, own_end, passes, tab, d_slot_of,
                                                                                 d_keep, d_winner);
        S_CUDA(cudaGetLastError());
        ++passes;
        // verify the pending items against their winners by replaying both
        int64_t *d_pending = nullptr, n_pending = 0;
        rc = compact_indices(n, PredFlagEq{d_keep, 3}, ps, &d_pending, &n_pending, st);
        if (rc != CYP_OK) return rc;
        S_CUDA(cudaMemsetAsync(d_unres, 0, sizeof(unsigned long long), st));
        if (n_pending) {
            const int64_t bm = std::max<int64_t>(1, std::min<int64_t>(n_pending, batch_rows / 2));
            int64_t *d_pair_ids = nullptr;
            ReplayItem *d_items = nullptr;
            int32_t *d_rows = nullptr;
            S_CUDA(ps.alloc(&d_pair_ids, 2 * bm));
            S_CUDA(ps.alloc(&d_items, 2 * bm));
            S_CUDA(ps.alloc(&d_rows, static_cast<size_t>(2 * bm) * ncol));
            for (int64_t b0 = 0; b0 < n_pending; b0 += bm) {
                const int64_t m = std::min(bm, n_pending - b0);
                make_pair_ids_kernel<<<blocks_for(m, kThreads), kThreads, 0, st>>>(m, d_pending + b0, d_winner, d_pair_ids);
                make_items_kernel<<<blocks_for(2 * m, kThreads), kThreads, 0, st>>>(2 * m, d_pair_ids, d_rec + 2, 3, nullptr, round,
                                                                                  s->d_sims_by_round, s->d_roots, d_items);
                S_CUDA(cudaGetLastError());
                rc = session_replay(s, 2 * m, d_items, d_rows, st);
                if (rc != CYP_OK) return rc;
                compare_pairs_kernel<<<blocks_for(m * 32, kThreads), kThreads, 0, st>>>(m, d_pending + b0, d_rows, ncol, d_keep, d_unres);
                S_CUDA(cudaGetLastError());
                replayed += 2 * m;
            }
        }
        unsigned long long unresolved = 0;
        S_CUDA(cudaMemcpyAsync(&unresolved, d_unres, sizeof(unresolved), cudaMemcpyDeviceToHost, st));
        S_CUDA(cudaStreamSynchronize(st));
        if (unresolved == 0) break;
        if (passes > 64) return fail(CYP_E_CUDA, "cyp_session_round: de-duplication did not converge");
        // next pass: every item that shares the key of an unresolved item, re-hashed from its replayed row with a new salt
        uint8_t *d_flag = nullptr;
        S_CUDA(ps.alloc(&d_flag, tab.mask + 1));
        S_CUDA(cudaMemsetAsync(d_flag, 0, tab.mask + 1, st));
        mark_slots_kernel<<<blocks_for(n_list, kThreads), kThreads, 0, st>>>(n_list, d_ids, own_begin, own_end, d_keep, d_slot_of, d_flag);
        int64_t *d_sel = nullptr, n_next = 0;
        rc = compact_indices(n_list, PredSlotFlagged{d_slot_of, d_flag, d_ids, d_keep, own_begin, own_end}, ps, &d_sel, &n_next, st);
        if (rc != CYP_OK) return rc;
        int64_t *d_next_ids = nullptr;
        uint64_t *d_next_hash = nullptr;
        S_CUDA(sc.alloc(&d_next_ids, n_next));
        S_CUDA(sc.alloc(&d_next_hash, 2 * n_next));
        map_ids_kernel<<<blocks_for(n_next, kThreads), kThreads, 0, st>>>(n_next, d_sel, d_ids, d_next_ids);
        {
            const int64_t bm = std::max<int64_t>(1, std::min<int64_t>(n_next, batch_rows));
            ReplayItem *d_items = nullptr;
            int32_t *d_rows = nullptr;
            S_CUDA(ps.alloc(&d_items, bm));
            S_CUDA(ps.alloc(&d_rows, static_cast<size_t>(bm) * ncol));
            for (int64_t b0 = 0; b0 < n_next; b0 += bm) {
                const int64_t m = std::min(bm, n_next - b0);
                make_items_kernel<<<blocks_for(m, kThreads), kThreads, 0, st>>>(m, d_next_ids + b0, d_rec + 2, 3, nullptr, round,
                                                                              s->d_sims_by_round, s->d_roots, d_items);
                S_CUDA(cudaGetLastError());
                rc = session_replay(s, m, d_items, d_rows, st);
                if (rc != CYP_OK) return rc;
                rehash_rows_kernel<<<blocks_for(m, kThreads), kThreads, 0, st>>>(m, d_rows, ncol, 0x5851F42D4C957F2Dull * static_cast<uint64_t>(passes),
                                                                                 d_next_hash + 2 * b0);
                S_CUDA(cudaGetLastError());
                replayed += m;
            }
        }
        S_CUDA(cudaStreamSynchronize(st));              // the per-pass scratch goes away
        d_ids = d_next_ids;
        n_list = n_next;
        d_hash = d_next_hash;
        hash_stride = 2;
        lo_mask = hi_mask = ~0ull;
    }
    *d_keep_out = d_keep;
    if (passes_out) *passes_out = passes;
    if (replayed_out) *replayed_out = replayed;
    return CYP_OK;
}

// Keeps the items of the last round's segment whose flag is 1 (order preserved); keep is indexed from `keep_begin`.
static int session_compact_round(cyp_session *s, const uint8_t *d_keep, cudaStream_t st, int64_t *n_keep_out) {
    const int64_t m = s->n_records;
    Scratch sc;
    int64_t *d_src = nullptr, n_keep = 0;
    int rc = compact_indices(m, PredFlagEq{d_keep, 1}, sc, &d_src, &n_keep, st);
    if (rc != CYP_OK) return rc;
    if (n_keep < m) {
        int32_t *t_round = nullptr, *t_slot = nullptr;
        int64_t *t_walker = nullptr;
        uint64_t *t_rec = nullptr;
        S_CUDA(sc.alloc(&t_round, n_keep));
        S_CUDA(sc.alloc(&t_walker, n_keep));
        S_CUDA(sc.alloc(&t_slot, n_keep));
        S_CUDA(sc.alloc(&t_rec, 3 * n_keep));
        if (n_keep) {
            const int64_t row0 = s->last_row0;
            move_index_kernel<<<blocks_for(n_keep, kThreads), kThreads, 0, st>>>(n_keep, d_src, s->d_k_round + row0, s->d_k_walker + row0,
                                                                                s->d_k_slot + row0, s->d_records, t_round, t_walker, t_slot, t_rec);
            S_CUDA(cudaGetLastError());
            S_CUDA(cudaMemcpyAsync(s->d_k_round + row0, t_round, sizeof(int32_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_k_walker + row0, t_walker, sizeof(int64_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_k_slot + row0, t_slot, sizeof(int32_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_records, t_rec, sizeof(uint64_t) * 3 * n_keep, cudaMemcpyDeviceToDevice, st));
        }
        S_CUDA(cudaStreamSynchronize(st));
        s->kept_rows = s->last_row0 + n_keep;
        s->n_records = n_keep;
    }
    *n_keep_out = n_keep;
    return CYP_OK;
}

// per-slot counts of a segment of an index (rows ending on every end point, :341-354)
int slot_histogram(const int32_t *d_slot, int64_t n, int32_t n_slots, int64_t *h_counts, cudaStream_t st) {
    Scratch sc;
    unsigned long long *d_counts = nullptr;
    S_CUDA(sc.alloc(&d_counts, n_slots));
    S_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * n_slots, st));
    if (n) slot_histogram_kernel<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, d_slot, d_counts);
    S_CUDA(cudaGetLastError());
    S_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(int64_t) * n_slots, cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    return CYP_OK;
}

void launch_make_items(int64_t n, const int64_t *d_ids, const uint64_t *d_walker, int walker_stride, const int32_t *d_round_of,
                       uint32_t round, const cyp_session *s, ReplayItem *d_items, cudaStream_t st) {
    if (n > 0)
        make_items_kernel<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, d_ids, d_walker, walker_stride, d_round_of, round,
                                                                      s->d_sims_by_round, s->d_roots, d_items);
}

}  // namespace cyp

extern "C" int cyp_session_create(const cyp_graph *g, const int32_t *h_root_cells, int64_t n_roots,
                                  const int32_t *h_cell_code, int32_t n_codes, int32_t min_clusters,
                                  const int32_t *h_end_slot_of_cell, int32_t n_end_slots, int32_t max_steps, int unique,
                                  uint64_t seed, cyp_session **out) {
    CYP_REQUIRE(out != nullptr, "cyp_session_create: out is NULL");
    *out = nullptr;
    CYP_REQUIRE(g != nullptr && h_root_cells && h_cell_code && h_end_slot_of_cell, "cyp_session_create: NULL argument");
    CYP_REQUIRE(n_roots > 0 && max_steps >= 1 && n_codes >= 1, "cyp_session_create: bad sizes");
    CYP_REQUIRE(n_codes <= 65536 || min_clusters <= 1, "cyp_session_create: more than 65536 distinct cluster labels");
    for (int64_t i = 0; i < n_roots; ++i)
        CYP_REQUIRE(h_root_cells[i] >= 0 && h_root_cells[i] < g->n, "cyp_session_create: root cell out of range");
    for (int64_t i = 0; i < g->n; ++i) {
        CYP_REQUIRE(h_cell_code[i] >= 0 && h_cell_code[i] < n_codes, "cyp_session_create: cell code out of range");
        CYP_REQUIRE(h_end_slot_of_cell[i] >= -1 && h_end_slot_of_cell[i] < n_end_slots, "cyp_session_create: end slot out of range");
    }
    DeviceGuard guard(g->device);
    static std::atomic<uint64_t> next_id{0};
    cyp_session *s = new cyp_session();
    s->id = ++next_id;
    s->g = g;
    s->device = g->device;
    s->n_roots = n_roots;
    s->n_codes = n_codes;
    s->min_clusters = min_clusters;
    s->n_end_slots = n_end_slots;
    s->max_steps = max_steps;
    s->unique = unique;
    s->seed = seed;
    cudaError_t e = dev_alloc(&s->d_roots, sizeof(int32_t) * n_roots);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_roots, h_root_cells, sizeof(int32_t) * n_roots, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = dev_alloc(&s->d_end_slot, sizeof(int32_t) * g->n);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_end_slot, h_end_slot_of_cell, sizeof(int32_t) * g->n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && min_clusters > 1) {
        if (n_codes <= 256) {
            std::vector<uint8_t> c8(static_cast<size_t>(g->n));
            for (int64_t i = 0; i < g->n; ++i) c8[i] = static_cast<uint8_t>(h_cell_code[i]);
            e = dev_alloc(&s->d_code8, g->n);
            if (e == cudaSuccess) e = cudaMemcpy(s->d_code8, c8.data(), g->n, cudaMemcpyHostToDevice);
        } else {
            e = dev_alloc(&s->d_code32, sizeof(int32_t) * g->n);
            if (e == cudaSuccess) e = cudaMemcpy(s->d_code32, h_cell_code, sizeof(int32_t) * g->n, cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        cyp_session_destroy(s);
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_session_create: ") + cudaGetErrorString(e));
    }
    *out = s;
    return CYP_OK;
}

extern "C" int cyp_session_destroy(cyp_session *s) {
    if (!s) return CYP_OK;
    DeviceGuard guard(s->device);
    dev_free(s->d_roots); dev_free(s->d_code8); dev_free(s->d_code32); dev_free(s->d_end_slot);
    dev_free(s->d_k_round); dev_free(s->d_k_walker); dev_free(s->d_k_slot);
    dev_free(s->d_records); dev_free(s->d_sims_by_round);
    delete s;
    return CYP_OK;
}

extern "C" int cyp_session_set_hash_bits(cyp_session *s, int bits) {
    CYP_REQUIRE(s != nullptr && bits >= 0 && bits <= 128, "cyp_session_set_hash_bits: bad argument");
    s->hash_bits = bits;
    return CYP_OK;
}

namespace {

// One chunk of a round: walkers [wb, we).  K2 in filter mode (no trajectories), K3 within the chunk, survivors
// appended to the index and to the round's record list.  Statistics are accumulated into *stats.
int run_chunk(cyp_session *s, uint32_t round, int64_t sims_per_root, int64_t wb, int64_t we, cyp_round_stats *stats) {
    cudaStream_t st = nullptr;
    const int64_t W = we - wb;
    const int32_t ncol = s->max_steps, nt = s->max_steps - 1;
    Tracer tr;
    Scratch sc;
    int32_t *d_seq = nullptr, *d_slot = nullptr;
    uint64_t *d_hash = nullptr;
    unsigned long long *d_counters = nullptr;     // [0] nocrossing [1] pass [2] terminal
    const bool need_rows = s->min_clusters > 1 && s->d_code32;       // > 256 labels: the generic filter reads the trajectories
    if (need_rows) S_CUDA(sc.alloc(&d_seq, static_cast<size_t>(W) * ncol));
    S_CUDA(sc.alloc(&d_slot, W));
    S_CUDA(sc.alloc(&d_hash, 2 * W));
    S_CUDA(sc.alloc(&d_counters, 4));
    S_CUDA(cudaMemsetAsync(d_counters, 0, 4 * sizeof(unsigned long long), st));
    tr.mark("chunk allocs");
    cudaEvent_t ev[3];
    for (auto &e : ev) S_CUDA(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 3; ++i) cudaEventDestroy(e[i]); } } evg{ev};

    // ---- K2
    if (s->d_code8 && walk_auto_is_v2(s->g)) {           // label the fused rows' headers for this session (once)
        if (s->g->labels2_owner != s->id) {
            label_rows2_kernel<<<blocks_for(s->g->n, kThreads), kThreads, 0, st>>>(s->g->n, s->g->d_word, s->g->d_rows2, s->d_code8);
            S_CUDA(cudaGetLastError());
            s->g->labels2_owner = s->id;
        }
    } else if (s->d_code8 && s->g->labels_owner != s->id) {     // label the edge records for this session (once)
        const int rce = ensure_edge_records(s->g);
        if (rce != CYP_OK) return rce;
        label_edges_kernel<<<blocks_for(s->g->padded, kThreads), kThreads, 0, st>>>(s->g->padded, s->g->d_edge, s->d_code8);
        S_CUDA(cudaGetLastError());
        s->g->labels_owner = s->id;
    }
    WalkParams p{};
    p.rows = s->g->d_row; p.cdf = s->g->d_cdf; p.edge = s->g->d_edge;
    p.roots = s->d_roots; p.sims_per_root = sims_per_root;
    p.walker_begin = wb; p.n_walkers = W; p.nt = nt; p.round = round; p.seed = s->seed;
    p.uniforms = nullptr; p.out_seq = d_seq; p.ld = ncol; p.nocross = d_counters;
    p.code8 = s->d_code8; p.min_clusters = s->min_clusters; p.end_slot = s->d_end_slot;
    p.out_slot = d_slot; p.out_hash = d_hash;
    WalkMode mode = kModeFilter0;
    if (s->min_clusters > 1 && s->d_code8) mode = (s->n_codes <= 64) ? kModeFilter1 : kModeFilter4;
    S_CUDA(cudaEventRecord(ev[0], st));
    int rc = launch_walk(s->g, p, mode, 0, st);
    if (rc != CYP_OK) return rc;
    S_CUDA(cudaEventRecord(ev[1], st));
    tr.mark("K2 launched");

    // ---- K3
    if (need_rows) {
        const int smem = 4 * ((s->n_codes + 31) / 32) * static_cast<int>(sizeof(unsigned));
        S_CUDA(cudaFuncSetAttribute(distinct_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((W + 3) / 4, 148 * 16));
        distinct_codes_kernel<<<grid, 128, smem, st>>>(W, d_seq, ncol, ncol, s->d_code32, s->n_codes, s->min_clusters, d_slot);
        S_CUDA(cudaGetLastError());
    }
    count_slots_kernel<<<blocks_for(W, kThreads), kThreads, 0, st>>>(W, d_slot, d_counters + 1);
    S_CUDA(cudaGetLastError());
    int64_t *d_cand = nullptr, n_cand = 0;
    rc = compact_indices(W, PredSlotGE{d_slot, 0}, sc, &d_cand, &n_cand, st);
    if (rc != CYP_OK) return rc;
    if (need_rows) {                                  // the trajectories have served their purpose
        dev_free(d_seq);
        sc.ptrs.erase(std::find(sc.ptrs.begin(), sc.ptrs.end(), static_cast<void *>(d_seq)));
        d_seq = nullptr;
    }
    tr.mark("compact candidates");
    uint64_t *d_cand_rec = nullptr;
    S_CUDA(sc.alloc(&d_cand_rec, 3 * n_cand));
    if (n_cand) make_records_kernel<<<blocks_for(n_cand, kThreads), kThreads, 0, st>>>(n_cand, d_cand, d_hash, wb, d_cand_rec);
    S_CUDA(cudaGetLastError());
    int64_t *d_kept = nullptr, n_kept = n_cand;      // candidate ranks that survive; nullptr = all
    int passes = 0;
    int64_t replayed = 0;
    if (s->unique && n_cand > 1) {
        uint8_t *d_keep = nullptr;
        rc = session_dedup(s, sc, n_cand, d_cand_rec, 0, n_cand, round, st, &d_keep, &passes, &replayed);
        if (rc != CYP_OK) return rc;
        rc = compact_indices(n_cand, PredFlagEq{d_keep, 1}, sc, &d_kept, &n_kept, st);
        if (rc != CYP_OK) return rc;
    }
    tr.mark("dedup");
    // ---- append survivors to the index
    rc = grow_index(s, s->kept_rows + n_kept);
    if (rc != CYP_OK) return rc;
    rc = ensure_records(s, s->n_records + n_kept);
    if (rc != CYP_OK) return rc;
    if (n_kept) {
        append_kept_kernel<<<blocks_for(n_kept, kThreads), kThreads, 0, st>>>(
            n_kept, d_kept, d_cand, d_slot, d_cand_rec, static_cast<int32_t>(round), s->d_k_round + s->kept_rows, s->d_k_walker + s->kept_rows,
            s->d_k_slot + s->kept_rows, s->d_records + 3 * s->n_records);
        S_CUDA(cudaGetLastError());
    }
    S_CUDA(cudaEventRecord(ev[2], st));
    unsigned long long counters[4] = {0, 0, 0, 0};
    S_CUDA(cudaMemcpyAsync(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    tr.mark("append + sync");
    float ms_walk = 0.f, ms_filter = 0.f;
    S_CUDA(cudaEventElapsedTime(&ms_walk, ev[0], ev[1]));
    S_CUDA(cudaEventElapsedTime(&ms_filter, ev[1], ev[2]));
    s->kept_rows += n_kept;
    s->n_records += n_kept;
    stats->ms_walk += ms_walk;
    stats->ms_filter += ms_filter;
    stats->n_nocrossing += static_cast<int64_t>(counters[0]);
    stats->n_pass_clusters += static_cast<int64_t>(counters[1]);
    stats->n_terminal += static_cast<int64_t>(counters[2]);
    stats->dedup_passes = std::max(stats->dedup_passes, passes);
    stats->n_replayed += replayed;
    return CYP_OK;
}

}  // namespace

extern "C" int cyp_session_set_chunk_walkers(cyp_session *s, int64_t walkers) {
    CYP_REQUIRE(s != nullptr && walkers >= 0, "cyp_session_set_chunk_walkers: bad argument");
    s->chunk_walkers = walkers;
    return CYP_OK;
}

extern "C" int cyp_session_round(cyp_session *s, uint32_t round, int64_t sims_per_root, int64_t walker_begin,
                                 int64_t walker_end, cyp_round_stats *stats) {
    CYP_REQUIRE(s != nullptr && stats != nullptr, "cyp_session_round: NULL argument");
    CYP_REQUIRE(sims_per_root > 0 && walker_begin >= 0 && walker_end >= walker_begin &&
                    walker_end <= s->n_roots * sims_per_root, "cyp_session_round: bad walker range");
    DeviceGuard guard(s->device);
    cudaStream_t st = nullptr;
    *stats = cyp_round_stats{};
    const int64_t W = walker_end - walker_begin;
    const int32_t ncol = s->max_steps;
    stats->n_walkers = W;
    s->n_records = 0;
    s->last_row0 = s->kept_rows;
    s->last_round = round;
    stats->store_rows = s->kept_rows;
    int rc = note_round(s, round, sims_per_root);
    if (rc != CYP_OK) return rc;
    if (W == 0) return CYP_OK;

    // A round is walked in chunks that fit the free HBM: per walker the hash and the slot, the K3 scratch if every
    // walker were a candidate (ranks, records, table, flags), and the index entries of the chunk's survivors; the
    // trajectories only for label sets above 256 (the generic min_clusters filter reads them).
    const double per_walker = 20.0 + 8.0 + 24.0 + 48.0 + 24.0 + 2.0 * 40.0 + ((s->min_clusters > 1 && s->d_code32) ? 4.0 * ncol : 0.0);
    int64_t chunk = s->chunk_walkers;
    if (chunk <= 0) {
        const double free_b = static_cast<double>(free_device_bytes(s->device));
        chunk = static_cast<int64_t>(0.7 * free_b / per_walker);
        if (chunk < 1024)
            return fail(CYP_E_NOMEM, "cyp_session_round: " + std::to_string(static_cast<long long>(free_b / (1 << 20))) +
                                         " MiB free is not enough for 1024 walkers x " + std::to_string(ncol) + " steps");
    }
    int n_chunks = 0;
    for (int64_t wb = walker_begin; wb < walker_end; wb += chunk, ++n_chunks) {
        const int64_t we = std::min(walker_end, wb + chunk);
        rc = run_chunk(s, round, sims_per_root, wb, we, stats);
        if (rc != CYP_OK) return rc;
    }
    // first occurrence across chunks: de-duplicate the round's records (chunks are in walker order)
    if (n_chunks > 1 && s->unique && s->n_records > 1) {
        Scratch sc;
        uint8_t *d_keep = nullptr;
        int passes = 0;
        int64_t replayed = 0;
        rc = session_dedup(s, sc, s->n_records, s->d_records, 0, s->n_records, round, st, &d_keep, &passes, &replayed);
        if (rc != CYP_OK) return rc;
        int64_t n_keep = 0;
        rc = session_compact_round(s, d_keep, st, &n_keep);
        if (rc != CYP_OK) return rc;
        stats->dedup_passes = std::max(stats->dedup_passes, passes);
        stats->n_replayed += replayed;
    }
    stats->n_kept = s->n_records;
    stats->store_rows = s->kept_rows;
    stats->n_chunks = n_chunks;
    if (stats->n_nocrossing)
        return fail(CYP_E_NOCROSSING, "cyp_session_round: a uniform exceeded its row's last CDF value (reference: IndexError at markov_chain_sampler.py:49)");
    return CYP_OK;
}

extern "C" int cyp_session_rows(const cyp_session *s, int64_t *n_rows) {
    CYP_REQUIRE(s != nullptr && n_rows != nullptr, "cyp_session_rows: NULL argument");
    *n_rows = s->kept_rows;
    return CYP_OK;
}

extern "C" int cyp_session_fetch_index(const cyp_session *s, int64_t row_begin, int64_t n_rows, int32_t *h_round,
                                       int64_t *h_walker, int32_t *h_end_slot) {
    CYP_REQUIRE(s != nullptr, "cyp_session_fetch_index: NULL session");
    CYP_REQUIRE(row_begin >= 0 && n_rows >= 0 && row_begin + n_rows <= s->kept_rows, "cyp_session_fetch_index: bad row range");
    DeviceGuard guard(s->device);
    if (n_rows == 0) return CYP_OK;
    if (h_round) S_CUDA(cudaMemcpy(h_round, s->d_k_round + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_walker) S_CUDA(cudaMemcpy(h_walker, s->d_k_walker + row_begin, sizeof(int64_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_end_slot) S_CUDA(cudaMemcpy(h_end_slot, s->d_k_slot + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    return CYP_OK;
}

extern "C" int cyp_session_slot_counts(const cyp_session *s, int64_t row_begin, int64_t n_rows, int64_t *h_counts) {
    CYP_REQUIRE(s != nullptr && h_counts != nullptr, "cyp_session_slot_counts: NULL argument");
    CYP_REQUIRE(row_begin >= 0 && n_rows >= 0 && row_begin + n_rows <= s->kept_rows, "cyp_session_slot_counts: bad row range");
    DeviceGuard guard(s->device);
    return slot_histogram(s->d_k_slot + row_begin, n_rows, s->n_end_slots, h_counts, nullptr);
}

extern "C" int cyp_session_fetch_rows(const cyp_session *s, const int64_t *h_rows, int64_t n_rows, int32_t *h_seq) {
    CYP_REQUIRE(s != nullptr && (n_rows == 0 || (h_rows && h_seq)), "cyp_session_fetch_rows: NULL argument");
    for (int64_t i = 0; i < n_rows; ++i)
        CYP_REQUIRE(h_rows[i] >= 0 && h_rows[i] < s->kept_rows, "cyp_session_fetch_rows: row out of range");
    if (n_rows == 0) return CYP_OK;
    DeviceGuard guard(s->device);
    cudaStream_t st = nullptr;
    Scratch sc;
    const int32_t ncol = s->max_steps;
    const int64_t bm = std::min<int64_t>(n_rows, session_replay_batch_rows(s));
    int64_t *d_ids = nullptr;
    ReplayItem *d_items = nullptr;
    int32_t *d_out = nullptr;
    S_CUDA(sc.alloc(&d_ids, n_rows));
    S_CUDA(sc.alloc(&d_items, bm));
    S_CUDA(sc.alloc(&d_out, static_cast<size_t>(bm) * ncol));
    S_CUDA(cudaMemcpyAsync(d_ids, h_rows, sizeof(int64_t) * n_rows, cudaMemcpyHostToDevice, st));
    for (int64_t b0 = 0; b0 < n_rows; b0 += bm) {
        const int64_t m = std::min(bm, n_rows - b0);
        make_items_kernel<<<blocks_for(m, kThreads), kThreads, 0, st>>>(m, d_ids + b0, reinterpret_cast<const uint64_t *>(s->d_k_walker), 1,
                                                                      s->d_k_round, 0u, s->d_sims_by_round, s->d_roots, d_items);
        S_CUDA(cudaGetLastError());
        const int rc = session_replay(s, m, d_items, d_out, st);
        if (rc != CYP_OK) return rc;
        S_CUDA(cudaMemcpyAsync(h_seq + b0 * ncol, d_out, sizeof(int32_t) * static_cast<size_t>(m) * ncol, cudaMemcpyDeviceToHost, st));
        S_CUDA(cudaStreamSynchronize(st));
    }
    return CYP_OK;
}

extern "C" int cyp_session_round_records(const cyp_session *s, uint64_t **d_records, int64_t *n_records) {
    CYP_REQUIRE(s != nullptr && d_records != nullptr && n_records != nullptr, "cyp_session_round_records: NULL argument");
    *d_records = s->d_records;
    *n_records = s->n_records;
    return CYP_OK;
}

extern "C" int cyp_session_round_prune(cyp_session *s, const uint64_t *d_all_records, int64_t n_all, int64_t *n_dropped) {
    CYP_REQUIRE(s != nullptr && n_dropped != nullptr, "cyp_session_round_prune: NULL argument");
    *n_dropped = 0;
    if (!s->unique || s->n_records == 0 || n_all == 0) return CYP_OK;
    CYP_REQUIRE(d_all_records != nullptr, "cyp_session_round_prune: NULL records");
    CYP_REQUIRE(n_all >= s->n_records, "cyp_session_round_prune: the gathered records do not include this rank's");
    DeviceGuard guard(s->device);
    cudaStream_t st = nullptr;
    Scratch sc;
    // where this rank's records sit in the gathered list (ranks hold ascending, contiguous walker blocks)
    long long *d_pos = nullptr;
    S_CUDA(sc.alloc(&d_pos, 1));
    S_CUDA(cudaMemsetAsync(d_pos, 0xFF, sizeof(long long), st));
    uint64_t first_walker = 0;
    S_CUDA(cudaMemcpyAsync(&first_walker, s->d_records + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    find_walker_kernel<<<blocks_for(n_all, kThreads), kThreads, 0, st>>>(n_all, d_all_records, first_walker, d_pos);
    long long own_begin = -1;
    S_CUDA(cudaMemcpyAsync(&own_begin, d_pos, sizeof(long long), cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    CYP_REQUIRE(own_begin >= 0 && own_begin + s->n_records <= n_all, "cyp_session_round_prune: this rank's records are not in the gathered list");
    uint8_t *d_keep = nullptr;
    int passes = 0;
    int rc = session_dedup(s, sc, n_all, d_all_records, own_begin, own_begin + s->n_records, s->last_round, st, &d_keep, &passes, nullptr);
    if (rc != CYP_OK) return rc;
    const int64_t m = s->n_records;
    int64_t n_keep = 0;
    rc = session_compact_round(s, d_keep + own_begin, st, &n_keep);
    if (rc != CYP_OK) return rc;
    *n_dropped = m - n_keep;
    return CYP_OK;
}
