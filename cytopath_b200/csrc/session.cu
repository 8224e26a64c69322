// a3: one sampling round and its filters (reference cytopath/markov_chain_sampler.py:299-354).
//
// K2 (walk.cu, filter mode) leaves per walker: the trajectory, a 128-bit hash of it and a
// slot code (-2: fewer than min_clusters labels, :322; -1: does not end on an end point;
// >= 0: index of the end-point cell it ends on, :343-349).  K3, in this file:
//   compact_*      order-preserving stream compaction (warp ballots + block offsets) of the
//                  walkers that passed both filters ("candidates"), in walker order;
//   dedup_insert   open-addressing hash table keyed by the sequence hash; every candidate
//                  does atomicMin(walker id) on its key's slot, so the winner is the first
//                  occurrence whatever the thread schedule (np.unique(..., return_index) +
//                  sorted, :330-331);
//   dedup_resolve  a non-winner is dropped only after its row compared equal to the
//                  winner's row; a true hash collision marks it unresolved and the
//                  unresolved set is re-hashed with another salt until none is left, so
//                  the result is exact for any hash quality;
//   gather_rows    survivors are appended, in walker order, to the device-resident store.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "walk.cuh"

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

// ---- de-duplication ---------------------------------------------------------------------
struct DedupTable {
    uint64_t *key_lo;     // 0 = empty
    uint64_t *key_hi;     // 0 = not yet written
    uint64_t *min_id;     // candidate rank of the lowest walker with this key
    uint64_t mask;        // capacity - 1
};

// pass 0 uses the hash K2 accumulated; later passes re-hash the stored rows with a salt
__global__ void __launch_bounds__(kThreads)
rehash_rows_kernel(int64_t n_items, const int64_t *__restrict__ items, const int64_t *__restrict__ cand,
                   const int32_t *__restrict__ seq, int64_t ld, int32_t n_cols, uint64_t salt, uint64_t *__restrict__ hash) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_items) return;
    const int64_t c = items[k];
    const int64_t rw = cand ? cand[c] : c;
    const int32_t *row = seq + rw * ld;
    uint64_t h1 = fmix(salt ^ 0x9E3779B97F4A7C15ull), h2 = fmix(salt + 0xD6E8FEB86659FD93ull);
    for (int i = 0; i < n_cols; ++i) {
        const uint64_t x = static_cast<uint32_t>(row[i]);
        h1 = fmix(h1 ^ (x + salt));
        h2 = (h2 ^ x) * 0x9FB21C651E98DF25ull + salt;
        h2 ^= h2 >> 31;
    }
    hash[2 * rw] = h1;
    hash[2 * rw + 1] = fmix(h2);
}

// items[k] = candidate rank (or nullptr: identity); cand[c] = row of candidate c in `seq`/`hash`
// (or nullptr: identity); the hash of row w sits at hash[stride*w], hash[stride*w + 1].
__global__ void __launch_bounds__(kThreads)
dedup_insert_kernel(int64_t n_items, const int64_t *__restrict__ items, const int64_t *__restrict__ cand,
                    const uint64_t *__restrict__ hash, int hash_stride, uint64_t lo_mask, uint64_t hi_mask, DedupTable t,
                    uint64_t *__restrict__ slot_of) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_items) return;
    const int64_t c = items ? items[k] : k;
    const int64_t w = cand ? cand[c] : c;
    const uint64_t lo = (hash[hash_stride * w] & lo_mask) | 1ull;          // never 0
    const uint64_t hi = (hash[hash_stride * w + 1] & hi_mask) | 1ull;
    uint64_t idx = fmix(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & t.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_lo + idx), 0ull, lo);
        if (prev == 0ull || prev == lo) {
            const unsigned long long ph = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_hi + idx), 0ull, hi);
            if (ph == 0ull || ph == hi) {
                atomicMin(reinterpret_cast<unsigned long long *>(t.min_id + idx), static_cast<unsigned long long>(c));
                slot_of[c] = idx;
                return;
            }
        }
        idx = (idx + 1) & t.mask;
    }
}

// One warp per item.  keep[c]: 1 = first occurrence, 0 = verified duplicate, 2 = unresolved.
__global__ void __launch_bounds__(kThreads)
dedup_resolve_kernel(int64_t n_items, const int64_t *__restrict__ items, const int64_t *__restrict__ cand,
                     const int32_t *__restrict__ seq, int64_t ld, int32_t n_cols, DedupTable t,
                     const uint64_t *__restrict__ slot_of, uint8_t *__restrict__ keep, unsigned long long *n_unresolved) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_items) return;
    const int64_t c = items ? items[k] : k;
    const int64_t winner = static_cast<int64_t>(t.min_id[slot_of[c]]);
    if (winner == c) {
        if (lane == 0) keep[c] = 1;
        return;
    }
    const int32_t *a = seq + (cand ? cand[c] : c) * ld;
    const int32_t *b = seq + (cand ? cand[winner] : winner) * ld;
    bool diff = false;
    for (int i = lane; i < n_cols; i += 32) diff |= (a[i] != b[i]);
    diff = __any_sync(0xffffffffu, diff);
    if (lane == 0) {
        keep[c] = diff ? 2 : 0;
        if (diff) atomicAdd(n_unresolved, 1ull);
    }
}

// ---- counting, gathering -------------------------------------------------------------------
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

// one warp per kept row: coalesced copy of the trajectory into the store + its index record
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(int64_t n_kept, const int64_t *__restrict__ kept /* candidate ranks or nullptr */,
                   const int64_t *__restrict__ cand, const int32_t *__restrict__ seq, int64_t ld, int32_t n_cols,
                   const int32_t *__restrict__ slot, const uint64_t *__restrict__ hash, int64_t walker_begin, int32_t round,
                   int32_t *__restrict__ store, int64_t store_row0, int32_t *__restrict__ st_round,
                   int64_t *__restrict__ st_walker, int32_t *__restrict__ st_slot, uint64_t *__restrict__ records) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_kept) return;
    const int64_t w = cand[kept ? kept[k] : k];
    const int32_t *src = seq + w * ld;
    int32_t *dst = store + (store_row0 + k) * static_cast<int64_t>(n_cols);
    for (int i = lane; i < n_cols; i += 32) dst[i] = src[i];
    if (lane == 0) {
        st_round[store_row0 + k] = round;
        st_walker[store_row0 + k] = walker_begin + w;
        st_slot[store_row0 + k] = slot[w];
        records[3 * k] = hash[2 * w];
        records[3 * k + 1] = hash[2 * w + 1];
        records[3 * k + 2] = static_cast<uint64_t>(walker_begin + w);
    }
}

__global__ void __launch_bounds__(kThreads)
fetch_rows_kernel(int64_t n_rows, const int64_t *__restrict__ rows, const int32_t *__restrict__ store, int32_t n_cols,
                  int32_t *__restrict__ out) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_rows) return;
    const int32_t *src = store + rows[k] * static_cast<int64_t>(n_cols);
    int32_t *dst = out + k * static_cast<int64_t>(n_cols);
    for (int i = lane; i < n_cols; i += 32) dst[i] = src[i];
}

// same for the fused rows (walk_v2.cu): the label code lives in byte 6 of every row block's header
__global__ void __launch_bounds__(256)
label_rows2_kernel(int64_t n, const uint32_t *__restrict__ word, unsigned char *__restrict__ rows2, const uint8_t *__restrict__ code8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) rows2[static_cast<size_t>(word[i] & kWordOffMask) * 32 + 6] = code8[i];
}

// generic min_clusters filter for label sets too large for the in-kernel bitmap (n_codes > 256):
// one warp per walker, bitmap of n_codes bits in shared memory.
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
// loop gets the label with the 16 bytes it loads anyway (no per-step code lookup)
__global__ void __launch_bounds__(kThreads)
label_edges_kernel(int64_t n_edges, EdgeRec *__restrict__ edge, const uint8_t *__restrict__ code8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n_edges) return;
    const int32_t c = edge[i].col;
    edge[i].code = (c >= 0) ? code8[c] : 0u;
}

// prune against records gathered from all ranks: drop[k] = 1 if another record with the same
// key has a lower walker id
__global__ void __launch_bounds__(kThreads)
records_insert_kernel(int64_t n, const uint64_t *__restrict__ rec, DedupTable t) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n) return;
    const uint64_t lo = rec[3 * k] | 1ull, hi = rec[3 * k + 1] | 1ull, id = rec[3 * k + 2];
    uint64_t idx = fmix(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & t.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_lo + idx), 0ull, lo);
        if (prev == 0ull || prev == lo) {
            const unsigned long long ph = atomicCAS(reinterpret_cast<unsigned long long *>(t.key_hi + idx), 0ull, hi);
            if (ph == 0ull || ph == hi) {
                atomicMin(reinterpret_cast<unsigned long long *>(t.min_id + idx), static_cast<unsigned long long>(id));
                return;
            }
        }
        idx = (idx + 1) & t.mask;
    }
}

__global__ void __launch_bounds__(kThreads)
records_lookup_kernel(int64_t n, const uint64_t *__restrict__ rec, DedupTable t, uint8_t *__restrict__ keep) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n) return;
    const uint64_t lo = rec[3 * k] | 1ull, hi = rec[3 * k + 1] | 1ull, id = rec[3 * k + 2];
    uint64_t idx = fmix(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & t.mask;
    while (true) {
        const uint64_t kl = t.key_lo[idx];
        if (kl == 0ull) { keep[k] = 1; return; }              // cannot happen: own record was inserted
        if (kl == lo && t.key_hi[idx] == hi) { keep[k] = (t.min_id[idx] == id) ? 1 : 0; return; }
        idx = (idx + 1) & t.mask;
    }
}

// in-place order-preserving compaction of the last round's store segment (rows only move down)
__global__ void __launch_bounds__(kThreads)
move_rows_kernel(int64_t n_keep, const int64_t *__restrict__ src_rank, int64_t row0, const int32_t *__restrict__ store_in,
                 int32_t n_cols, const int32_t *__restrict__ r_in, const int64_t *__restrict__ w_in,
                 const int32_t *__restrict__ s_in, const uint64_t *__restrict__ rec_in, int32_t *__restrict__ store_out,
                 int32_t *__restrict__ r_out, int64_t *__restrict__ w_out, int32_t *__restrict__ s_out,
                 uint64_t *__restrict__ rec_out) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_keep) return;
    const int64_t s = src_rank[k];
    const int32_t *src = store_in + s * static_cast<int64_t>(n_cols);
    int32_t *dst = store_out + k * static_cast<int64_t>(n_cols);
    for (int i = lane; i < n_cols; i += 32) dst[i] = src[i];
    if (lane == 0) {
        r_out[k] = r_in[row0 + s];
        w_out[k] = w_in[row0 + s];
        s_out[k] = s_in[row0 + s];
        rec_out[3 * k] = rec_in[3 * s];
        rec_out[3 * k + 1] = rec_in[3 * s + 1];
        rec_out[3 * k + 2] = rec_in[3 * s + 2];
    }
}

}  // namespace cyp

using namespace cyp;

struct cyp_session {
    uint64_t id = 0;
    const cyp_graph *g = nullptr;
    int device = 0;
    int64_t n_roots = 0;
    int32_t n_codes = 0, min_clusters = 0, n_end_slots = 0, max_steps = 0;
    int unique = 1;
    uint64_t seed = 0;
    int hash_bits = 128;
    int32_t *d_roots = nullptr;
    uint8_t *d_code8 = nullptr;
    int32_t *d_code32 = nullptr;
    int32_t *d_end_slot = nullptr;
    // device-resident store of kept sequences (accumulation order of :335-340)
    int32_t *d_store = nullptr;
    int32_t *d_st_round = nullptr;
    int64_t *d_st_walker = nullptr;
    int32_t *d_st_slot = nullptr;
    int64_t store_rows = 0, store_cap = 0;
    // records of the rows appended by the last round
    uint64_t *d_records = nullptr;
    int64_t n_records = 0, records_cap = 0;
    int64_t last_row0 = 0;
    int64_t chunk_walkers = 0;   // 0 = as many walkers per chunk as the free HBM allows
};

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
    S_CUDA(sc.alloc(&t->min_id, cap));
    S_CUDA(cudaMemsetAsync(t->key_lo, 0, cap * 8, st));
    S_CUDA(cudaMemsetAsync(t->key_hi, 0, cap * 8, st));
    S_CUDA(cudaMemsetAsync(t->min_id, 0xFF, cap * 8, st));
    t->mask = cap - 1;
    return CYP_OK;
}

int grow_store(cyp_session *s, int64_t need_rows) {
    if (need_rows <= s->store_cap) return CYP_OK;
    int64_t cap = std::max<int64_t>(need_rows, s->store_cap + s->store_cap / 2);
    cap = std::max<int64_t>(cap, 1024);
    int32_t *n_store = nullptr, *n_round = nullptr, *n_slot = nullptr;
    int64_t *n_walker = nullptr;
    const size_t row_bytes = sizeof(int32_t) * static_cast<size_t>(s->max_steps);
    cudaError_t e = dev_alloc(&n_store, row_bytes * cap);
    if (e == cudaSuccess) e = dev_alloc(&n_round, sizeof(int32_t) * cap);
    if (e == cudaSuccess) e = dev_alloc(&n_walker, sizeof(int64_t) * cap);
    if (e == cudaSuccess) e = dev_alloc(&n_slot, sizeof(int32_t) * cap);
    if (e == cudaSuccess && s->store_rows) {
        e = cudaMemcpy(n_store, s->d_store, row_bytes * s->store_rows, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(n_round, s->d_st_round, sizeof(int32_t) * s->store_rows, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(n_walker, s->d_st_walker, sizeof(int64_t) * s->store_rows, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(n_slot, s->d_st_slot, sizeof(int32_t) * s->store_rows, cudaMemcpyDeviceToDevice);
    }
    if (e != cudaSuccess) {
        dev_free(n_store); dev_free(n_round); dev_free(n_walker); dev_free(n_slot);
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("grow_store: ") + cudaGetErrorString(e));
    }
    dev_free(s->d_store); dev_free(s->d_st_round); dev_free(s->d_st_walker); dev_free(s->d_st_slot);
    s->d_store = n_store; s->d_st_round = n_round; s->d_st_walker = n_walker; s->d_st_slot = n_slot;
    s->store_cap = cap;
    return CYP_OK;
}

void hash_masks(int bits, uint64_t *lo, uint64_t *hi) {
    if (bits >= 128) { *lo = ~0ull; *hi = ~0ull; }
    else if (bits >= 64) { *lo = ~0ull; *hi = (bits == 64) ? 0ull : ((1ull << (bits - 64)) - 1ull); }
    else { *lo = (bits <= 0) ? 0ull : ((1ull << bits) - 1ull); *hi = 0ull; }
}

}  // namespace

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
    static uint64_t next_id = 0;
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
    dev_free(s->d_store); dev_free(s->d_st_round); dev_free(s->d_st_walker); dev_free(s->d_st_slot);
    dev_free(s->d_records);
    delete s;
    return CYP_OK;
}

extern "C" int cyp_session_set_hash_bits(cyp_session *s, int bits) {
    CYP_REQUIRE(s != nullptr && bits >= 0 && bits <= 128, "cyp_session_set_hash_bits: bad argument");
    s->hash_bits = bits;
    return CYP_OK;
}

namespace {

// First-occurrence de-duplication of n_cand rows (candidate c = row `cand ? cand[c] : c` of `seq`, hash at
// hash[stride*row]).  keep[c]: 1 = first occurrence, 0 = verified duplicate.  Exact for any hash quality:
// non-winners are compared against the winner's row, true collisions are re-hashed with another salt.
int dedup_rows(cyp_session *s, Scratch &sc, int64_t n_cand, const int64_t *d_cand, const int32_t *seq, int64_t ld,
               int32_t ncol, const uint64_t *hash, int hash_stride, int64_t hash_rows, unsigned long long *d_unresolved,
               cudaStream_t st, uint8_t **d_keep_out, int *passes_out) {
    uint8_t *d_keep = nullptr;
    uint64_t *d_slot_of = nullptr;
    S_CUDA(sc.alloc(&d_keep, n_cand));
    S_CUDA(sc.alloc(&d_slot_of, n_cand));
    int64_t *d_items = nullptr, n_items = n_cand;      // nullptr = all candidates
    const uint64_t *hash_pass = hash;                   // later passes re-hash into their own buffer
    uint64_t *d_hash_pass = nullptr;
    int stride = hash_stride;
    uint64_t lo_mask, hi_mask;
    hash_masks(s->hash_bits, &lo_mask, &hi_mask);
    int passes = 0;
    while (n_items > 0) {
        Scratch pass_sc;
        DedupTable tab{};
        int rc = alloc_table(n_items, pass_sc, &tab, st);
        if (rc != CYP_OK) return rc;
        if (passes > 0) {
            if (!d_hash_pass) S_CUDA(sc.alloc(&d_hash_pass, 2 * hash_rows));
            rehash_rows_kernel<<<blocks_for(n_items, kThreads), kThreads, 0, st>>>(
                n_items, d_items, d_cand, seq, ld, ncol, 0x5851F42D4C957F2Dull * static_cast<uint64_t>(passes), d_hash_pass);
            hash_pass = d_hash_pass;
            stride = 2;
            lo_mask = hi_mask = ~0ull;
        }
        S_CUDA(cudaMemsetAsync(d_unresolved, 0, sizeof(unsigned long long), st));
        dedup_insert_kernel<<<blocks_for(n_items, kThreads), kThreads, 0, st>>>(n_items, d_items, d_cand, hash_pass, stride,
                                                                               lo_mask, hi_mask, tab, d_slot_of);
        dedup_resolve_kernel<<<blocks_for(n_items * 32, kThreads), kThreads, 0, st>>>(n_items, d_items, d_cand, seq, ld, ncol, tab,
                                                                                     d_slot_of, d_keep, d_unresolved);
        S_CUDA(cudaGetLastError());
        unsigned long long unresolved = 0;
        S_CUDA(cudaMemcpyAsync(&unresolved, d_unresolved, sizeof(unresolved), cudaMemcpyDeviceToHost, st));
        S_CUDA(cudaStreamSynchronize(st));
        ++passes;
        if (unresolved == 0) break;
        if (passes > 64) return fail(CYP_E_CUDA, "cyp_session_round: de-duplication did not converge");
        rc = compact_indices(n_cand, PredFlagEq{d_keep, 2}, sc, &d_items, &n_items, st);
        if (rc != CYP_OK) return rc;
    }
    *d_keep_out = d_keep;
    *passes_out = passes;
    return CYP_OK;
}

int ensure_records(cyp_session *s, int64_t need) {
    if (need <= s->records_cap) return CYP_OK;
    const int64_t cap = std::max<int64_t>(need, s->records_cap + s->records_cap / 2);
    uint64_t *n_rec = nullptr;
    S_CUDA(dev_alloc(&n_rec, sizeof(uint64_t) * 3 * cap));
    if (s->n_records) S_CUDA(cudaMemcpy(n_rec, s->d_records, sizeof(uint64_t) * 3 * s->n_records, cudaMemcpyDeviceToDevice));
    dev_free(s->d_records);
    s->d_records = n_rec;
    s->records_cap = cap;
    return CYP_OK;
}

// Keeps the rows of the last round's store segment whose flag is 1 (order preserved).
int compact_last_round(cyp_session *s, const uint8_t *d_keep, Scratch &sc, cudaStream_t st, int64_t *n_keep_out) {
    const int64_t m = s->n_records;
    int64_t *d_src = nullptr, n_keep = 0;
    int rc = compact_indices(m, PredFlagEq{d_keep, 1}, sc, &d_src, &n_keep, st);
    if (rc != CYP_OK) return rc;
    if (n_keep < m) {
        // move through a temporary so overlapping source/destination rows cannot race
        int32_t *t_store = nullptr, *t_round = nullptr, *t_slot = nullptr;
        int64_t *t_walker = nullptr;
        uint64_t *t_rec = nullptr;
        const int32_t ncol = s->max_steps;
        S_CUDA(sc.alloc(&t_store, static_cast<size_t>(n_keep) * ncol));
        S_CUDA(sc.alloc(&t_round, n_keep));
        S_CUDA(sc.alloc(&t_walker, n_keep));
        S_CUDA(sc.alloc(&t_slot, n_keep));
        S_CUDA(sc.alloc(&t_rec, 3 * n_keep));
        if (n_keep) {
            const int64_t row0 = s->last_row0;
            move_rows_kernel<<<blocks_for(n_keep * 32, kThreads), kThreads, 0, st>>>(
                n_keep, d_src, row0, s->d_store + row0 * ncol, ncol, s->d_st_round, s->d_st_walker, s->d_st_slot, s->d_records,
                t_store, t_round, t_walker, t_slot, t_rec);
            S_CUDA(cudaGetLastError());
            S_CUDA(cudaMemcpyAsync(s->d_store + row0 * ncol, t_store, sizeof(int32_t) * static_cast<size_t>(n_keep) * ncol, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_st_round + row0, t_round, sizeof(int32_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_st_walker + row0, t_walker, sizeof(int64_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_st_slot + row0, t_slot, sizeof(int32_t) * n_keep, cudaMemcpyDeviceToDevice, st));
            S_CUDA(cudaMemcpyAsync(s->d_records, t_rec, sizeof(uint64_t) * 3 * n_keep, cudaMemcpyDeviceToDevice, st));
        }
        S_CUDA(cudaStreamSynchronize(st));
        s->store_rows = s->last_row0 + n_keep;
        s->n_records = n_keep;
    }
    *n_keep_out = n_keep;
    return CYP_OK;
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

// One chunk of a round: walkers [wb, we).  K2 in filter mode, K3 within the chunk, survivors appended to
// the store and to the round's record list.  Statistics are accumulated into *stats.
int run_chunk(cyp_session *s, uint32_t round, int64_t sims_per_root, int64_t wb, int64_t we, const double *h_uniforms,
              cyp_round_stats *stats) {
    cudaStream_t st = nullptr;
    const int64_t W = we - wb;
    const int32_t ncol = s->max_steps, nt = s->max_steps - 1;
    Tracer tr;
    Scratch sc;
    int32_t *d_seq = nullptr, *d_slot = nullptr;
    uint64_t *d_hash = nullptr;
    double *d_u = nullptr;
    unsigned long long *d_counters = nullptr;     // [0] nocrossing [1] pass [2] terminal [3] unresolved
    S_CUDA(sc.alloc(&d_seq, static_cast<size_t>(W) * ncol));
    S_CUDA(sc.alloc(&d_slot, W));
    S_CUDA(sc.alloc(&d_hash, 2 * W));
    S_CUDA(sc.alloc(&d_counters, 4));
    S_CUDA(cudaMemsetAsync(d_counters, 0, 4 * sizeof(unsigned long long), st));
    if (h_uniforms && nt > 0) {
        S_CUDA(sc.alloc(&d_u, static_cast<size_t>(W) * nt));
        S_CUDA(cudaMemcpyAsync(d_u, h_uniforms, sizeof(double) * static_cast<size_t>(W) * nt, cudaMemcpyHostToDevice, st));
    }
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
    p.uniforms = d_u; p.out_seq = d_seq; p.ld = ncol; p.nocross = d_counters;
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
    if (s->min_clusters > 1 && s->d_code32) {
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

    tr.mark("compact candidates");
    int64_t *d_kept = nullptr, n_kept = n_cand;      // candidate ranks that survive; nullptr = all
    int passes = 0;
    if (s->unique && n_cand > 1) {
        uint8_t *d_keep = nullptr;
        rc = dedup_rows(s, sc, n_cand, d_cand, d_seq, ncol, ncol, d_hash, 2, W, d_counters + 3, st, &d_keep, &passes);
        if (rc != CYP_OK) return rc;
        rc = compact_indices(n_cand, PredFlagEq{d_keep, 1}, sc, &d_kept, &n_kept, st);
        if (rc != CYP_OK) return rc;
    }

    tr.mark("dedup");
    // ---- append survivors to the store
    rc = grow_store(s, s->store_rows + n_kept);
    if (rc != CYP_OK) return rc;
    tr.mark("grow_store");
    rc = ensure_records(s, s->n_records + n_kept);
    if (rc != CYP_OK) return rc;
    if (n_kept) {
        gather_rows_kernel<<<blocks_for(n_kept * 32, kThreads), kThreads, 0, st>>>(
            n_kept, d_kept, d_cand, d_seq, ncol, ncol, d_slot, d_hash, wb, static_cast<int32_t>(round), s->d_store,
            s->store_rows, s->d_st_round, s->d_st_walker, s->d_st_slot, s->d_records + 3 * s->n_records);
        S_CUDA(cudaGetLastError());
    }
    S_CUDA(cudaEventRecord(ev[2], st));
    unsigned long long counters[4] = {0, 0, 0, 0};
    S_CUDA(cudaMemcpyAsync(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
    S_CUDA(cudaStreamSynchronize(st));
    tr.mark("gather + sync");
    float ms_walk = 0.f, ms_filter = 0.f;
    S_CUDA(cudaEventElapsedTime(&ms_walk, ev[0], ev[1]));
    S_CUDA(cudaEventElapsedTime(&ms_filter, ev[1], ev[2]));
    s->store_rows += n_kept;
    s->n_records += n_kept;
    stats->ms_walk += ms_walk;
    stats->ms_filter += ms_filter;
    stats->n_nocrossing += static_cast<int64_t>(counters[0]);
    stats->n_pass_clusters += static_cast<int64_t>(counters[1]);
    stats->n_terminal += static_cast<int64_t>(counters[2]);
    stats->dedup_passes = std::max(stats->dedup_passes, passes);
    return CYP_OK;
}

}  // namespace

extern "C" int cyp_session_set_chunk_walkers(cyp_session *s, int64_t walkers) {
    CYP_REQUIRE(s != nullptr && walkers >= 0, "cyp_session_set_chunk_walkers: bad argument");
    s->chunk_walkers = walkers;
    return CYP_OK;
}

extern "C" int cyp_session_round(cyp_session *s, uint32_t round, int64_t sims_per_root, int64_t walker_begin,
                                 int64_t walker_end, const double *h_uniforms, cyp_round_stats *stats) {
    CYP_REQUIRE(s != nullptr && stats != nullptr, "cyp_session_round: NULL argument");
    CYP_REQUIRE(sims_per_root > 0 && walker_begin >= 0 && walker_end >= walker_begin &&
                    walker_end <= s->n_roots * sims_per_root, "cyp_session_round: bad walker range");
    DeviceGuard guard(s->device);
    cudaStream_t st = nullptr;
    *stats = cyp_round_stats{};
    const int64_t W = walker_end - walker_begin;
    const int32_t ncol = s->max_steps, nt = s->max_steps - 1;
    stats->n_walkers = W;
    s->n_records = 0;
    s->last_row0 = s->store_rows;
    stats->store_rows = s->store_rows;
    if (W == 0) return CYP_OK;

    // A round is walked in chunks that fit the free HBM: per walker the trajectory, hash and slot
    // (+ injected uniforms), the K3 scratch if every walker were a candidate, and room for the store to
    // take the chunk's survivors (grow_store copies, hence the factor).
    const double per_walker = 4.0 * ncol + 20.0 + (h_uniforms ? 8.0 * nt : 0.0) + 120.0 + 2.5 * 4.0 * ncol;
    int64_t chunk = s->chunk_walkers;
    Tracer tr;
    if (chunk <= 0) {
        const double free_b = static_cast<double>(free_device_bytes(s->device));
        tr.mark("free_device_bytes");
        chunk = static_cast<int64_t>(0.8 * free_b / per_walker);
        if (chunk < 1024)
            return fail(CYP_E_NOMEM, "cyp_session_round: " + std::to_string(static_cast<long long>(free_b / (1 << 20))) +
                                         " MiB free is not enough for 1024 walkers x " + std::to_string(ncol) + " steps");
    }
    int n_chunks = 0;
    for (int64_t wb = walker_begin; wb < walker_end; wb += chunk, ++n_chunks) {
        const int64_t we = std::min(walker_end, wb + chunk);
        const double *u = h_uniforms ? h_uniforms + (wb - walker_begin) * static_cast<int64_t>(nt) : nullptr;
        const int rc = run_chunk(s, round, sims_per_root, wb, we, u, stats);
        if (rc != CYP_OK) return rc;
    }
    // first occurrence across chunks: de-duplicate the round's store segment (chunks are in walker order)
    if (n_chunks > 1 && s->unique && s->n_records > 1) {
        Scratch sc;
        unsigned long long *d_unres = nullptr;
        S_CUDA(sc.alloc(&d_unres, 1));
        uint8_t *d_keep = nullptr;
        int passes = 0;
        int rc = dedup_rows(s, sc, s->n_records, nullptr, s->d_store + s->last_row0 * ncol, ncol, ncol, s->d_records, 3,
                            s->n_records, d_unres, st, &d_keep, &passes);
        if (rc != CYP_OK) return rc;
        int64_t n_keep = 0;
        rc = compact_last_round(s, d_keep, sc, st, &n_keep);
        if (rc != CYP_OK) return rc;
        stats->dedup_passes = std::max(stats->dedup_passes, passes);
    }
    stats->n_kept = s->n_records;
    stats->store_rows = s->store_rows;
    stats->n_chunks = n_chunks;
    if (stats->n_nocrossing)
        return fail(CYP_E_NOCROSSING, "cyp_session_round: a uniform exceeded its row's last CDF value (reference: IndexError at markov_chain_sampler.py:49)");
    return CYP_OK;
}

extern "C" int cyp_session_rows(const cyp_session *s, int64_t *n_rows) {
    CYP_REQUIRE(s != nullptr && n_rows != nullptr, "cyp_session_rows: NULL argument");
    *n_rows = s->store_rows;
    return CYP_OK;
}

extern "C" int cyp_session_fetch_index(const cyp_session *s, int64_t row_begin, int64_t n_rows, int32_t *h_round,
                                       int64_t *h_walker, int32_t *h_end_slot) {
    CYP_REQUIRE(s != nullptr, "cyp_session_fetch_index: NULL session");
    CYP_REQUIRE(row_begin >= 0 && n_rows >= 0 && row_begin + n_rows <= s->store_rows, "cyp_session_fetch_index: bad row range");
    DeviceGuard guard(s->device);
    if (n_rows == 0) return CYP_OK;
    if (h_round) S_CUDA(cudaMemcpy(h_round, s->d_st_round + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_walker) S_CUDA(cudaMemcpy(h_walker, s->d_st_walker + row_begin, sizeof(int64_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_end_slot) S_CUDA(cudaMemcpy(h_end_slot, s->d_st_slot + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    return CYP_OK;
}

extern "C" int cyp_session_fetch_rows(const cyp_session *s, const int64_t *h_rows, int64_t n_rows, int32_t *h_seq) {
    CYP_REQUIRE(s != nullptr && (n_rows == 0 || (h_rows && h_seq)), "cyp_session_fetch_rows: NULL argument");
    for (int64_t i = 0; i < n_rows; ++i)
        CYP_REQUIRE(h_rows[i] >= 0 && h_rows[i] < s->store_rows, "cyp_session_fetch_rows: row out of range");
    if (n_rows == 0) return CYP_OK;
    DeviceGuard guard(s->device);
    Scratch sc;
    int64_t *d_rows = nullptr;
    int32_t *d_out = nullptr;
    S_CUDA(sc.alloc(&d_rows, n_rows));
    S_CUDA(sc.alloc(&d_out, static_cast<size_t>(n_rows) * s->max_steps));
    S_CUDA(cudaMemcpy(d_rows, h_rows, sizeof(int64_t) * n_rows, cudaMemcpyHostToDevice));
    fetch_rows_kernel<<<blocks_for(n_rows * 32, kThreads), kThreads>>>(n_rows, d_rows, s->d_store, s->max_steps, d_out);
    S_CUDA(cudaGetLastError());
    S_CUDA(cudaMemcpy(h_seq, d_out, sizeof(int32_t) * static_cast<size_t>(n_rows) * s->max_steps, cudaMemcpyDeviceToHost));
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
    DeviceGuard guard(s->device);
    cudaStream_t st = nullptr;
    Scratch sc;
    DedupTable tab{};
    int rc = alloc_table(n_all, sc, &tab, st);
    if (rc != CYP_OK) return rc;
    records_insert_kernel<<<blocks_for(n_all, kThreads), kThreads, 0, st>>>(n_all, d_all_records, tab);
    uint8_t *d_keep = nullptr;
    const int64_t m = s->n_records;
    S_CUDA(sc.alloc(&d_keep, m));
    records_lookup_kernel<<<blocks_for(m, kThreads), kThreads, 0, st>>>(m, s->d_records, tab, d_keep);
    S_CUDA(cudaGetLastError());
    int64_t n_keep = 0;
    rc = compact_last_round(s, d_keep, sc, st, &n_keep);
    if (rc != CYP_OK) return rc;
    S_CUDA(cudaStreamSynchronize(st));
    *n_dropped = m - n_keep;
    return CYP_OK;
}
