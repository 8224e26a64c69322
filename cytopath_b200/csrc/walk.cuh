// a2: markov_sim (reference cytopath/markov_chain_sampler.py:35-52) — the step-loop kernel K2.
#pragma once

#include "common.cuh"

namespace cyp {

// A walker named explicitly (instead of by its position in a contiguous range of one round): what the replay
// launches of session.cu pass.  Because the uniform stream is keyed by (walker, step, round), walking the same
// item again reproduces the same trajectory: kept sequences are stored as these 16 bytes, not as rows.
struct ReplayItem {
    int64_t walker;     // global walker id of its round
    uint32_t round;
    int32_t root;       // root cell (= root_cells[walker / sims_per_root] of that round)
};

struct WalkParams {
    const RowInfo *rows;            // [n_cells] row descriptor (read for the root cell only)
    const void *cdf;                // float / double, sector-aligned rows
    const EdgeRec *edge;            // same offsets as cdf: target cell + target's row descriptor
    const int32_t *roots;           // [n_roots]
    int64_t sims_per_root;
    int64_t walker_begin;           // global id of local walker 0
    int64_t n_walkers;
    int32_t nt;                     // transitions per walker (max_steps - 1)
    uint32_t round;
    uint64_t seed;
    const double *uniforms;         // [n_walkers, nt] or nullptr -> Philox stream
    const ReplayItem *replay = nullptr;   // [n_walkers] or nullptr: walker w is walker_begin + w of round `round`, starting at roots[.. / sims_per_root]
    int32_t *out_seq;               // [n_walkers, ld]; nullptr: no trajectory is written (filter modes: hash + slot only)
    int64_t ld;                     // row stride of out_seq in elements (>= nt + 1)
    unsigned long long *nocross;    // may be nullptr
    // filter outputs, only read/written by the FILTER modes
    const uint8_t *code8;           // [n_cells] code of the truncated cluster label (n_codes <= 256); the step
                                    // loop reads the same code from EdgeRec.code, this array serves the roots
    int32_t min_clusters;
    const int32_t *end_slot;        // [n_cells] end-point slot or -1
    int32_t *out_slot;              // [n_walkers]: -2 fails min_clusters, -1 not terminal, else end slot
    uint64_t *out_hash;             // [n_walkers, 2] 128-bit sequence hash
    // fused rows (walk_v2.cu; filled in by launch_walk_v2 from the graph)
    const unsigned char *rows2 = nullptr;
    const uint32_t *edge4 = nullptr;
    const uint32_t *word = nullptr;
    int l1_rows = 0;                // row copies allocate in L1 (small graphs: many walkers per cell)
};

// Kernel modes: what is fused into the step loop besides the walk itself.
enum WalkMode : int {
    kModePlain = 0,      // trajectories only
    kModeFilter0 = 1,    // + hash + terminal slot, no cluster bitmap (min_clusters <= 1)
    kModeFilter1 = 2,    // + 64-bit cluster bitmap  (n_codes <= 64)
    kModeFilter4 = 3,    // + 256-bit cluster bitmap (n_codes <= 256)
};

struct WalkVariant {
    int lanes;    // G: lanes cooperating on one walker (1 = thread per walker ... 32 = warp per walker)
    int unroll;   // independent 16-byte loads per lane per chunk
};

#ifdef __CUDACC__
// 16-byte edge-record load that does not allocate an L1 line: the record is used once, and the
// L1 left over by the shared-memory carve-out is needed for in-flight misses (profiles/r01).
__device__ __forceinline__ EdgeRec load_edge(const EdgeRec *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return EdgeRec{v.x, static_cast<uint32_t>(v.y), static_cast<uint32_t>(v.z), static_cast<uint32_t>(v.w)};
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {      // murmur3 fmix64
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// What the filter modes accumulate along a trajectory: a 128-bit hash of the cell sequence
// (de-duplication key, :329-333) and the bitmap of truncated cluster labels visited (:322).
template <int MODE>
struct WalkerDigest {
    static constexpr int MASKW = (MODE == kModeFilter1) ? 1 : (MODE == kModeFilter4) ? 4 : 0;
    static constexpr bool FILTER = MODE != kModePlain;
    uint64_t h1 = 0x9E3779B97F4A7C15ull, h2 = 0xD6E8FEB86659FD93ull;
    uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;          // label bitmap words (scalars: no local-memory array)

    // `code` = label code of `cell`: from the edge record for steps, from code8 for the root
    __device__ __forceinline__ void visit(int32_t cell, unsigned code) {
        if constexpr (FILTER) {
            h1 = mix64(h1 ^ static_cast<uint32_t>(cell));
            h2 = (h2 + static_cast<uint32_t>(cell)) * 0x9FB21C651E98DF25ull;
            h2 ^= h2 >> 29;
        }
        if constexpr (MASKW > 0) {
            const uint64_t bit = 1ull << (code & 63);
            if constexpr (MASKW == 1) {
                m0 |= bit;
            } else {
                const unsigned word = code >> 6;
                m0 |= (word == 0) ? bit : 0ull;
                m1 |= (word == 1) ? bit : 0ull;
                m2 |= (word == 2) ? bit : 0ull;
                m3 |= (word == 3) ? bit : 0ull;
            }
        }
    }
    // one thread per walker calls this after the last step
    __device__ __forceinline__ void finish(const WalkParams &p, int64_t w, int32_t last_cell) const {
        if constexpr (FILTER) {
            const int distinct = __popcll(m0) + __popcll(m1) + __popcll(m2) + __popcll(m3);
            const bool pass = (MASKW == 0) || distinct >= p.min_clusters;
            p.out_slot[w] = pass ? p.end_slot[last_cell] : -2;
            p.out_hash[2 * w] = h1;
            p.out_hash[2 * w + 1] = mix64(h2);
        }
    }
};
#endif

// Chooses lanes/unroll from the mean row length so that one chunk covers a typical row.
WalkVariant pick_walk_variant(const cyp_graph *g, int variant);

// placement.cu: times a short walk of the caller's walkers on the current placement of the fused rows and on three fresh
// allocations, keeps the fastest (once per graph, before its first large walk)
int calibrate_placement(cyp_graph *g, const WalkParams &like, float *ms_before, float *ms_after, int *moves);

// Enqueues K2 on `stream`.  Returns a CYP_* code.
int launch_walk(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream);

// K2, TMA flavour (walk_tma.cu): one walker per thread, rows staged in shared memory by
// cp.async.bulk.  `variant` = 9000 + threads per CTA / 32 (e.g. 9016 = 512 threads); 0 = auto.
int launch_walk_tma(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream);
const char *walk_tma_name(const cyp_graph *g, int variant);

constexpr int kNotApplicable = 1;

// K2, fused-row flavour (walk_v2.cu, the default where cyp_graph::v2_ok): `variant` = 9500 + threads per
// CTA / 32; 0 = auto.  Returns kNotApplicable (nothing launched) when the graph has no fused rows.
int launch_walk_v2(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream);
const char *walk_v2_name(const cyp_graph *g, int variant);
// true when launch_walk(variant 0) runs the fused-row kernel on this graph
bool walk_auto_is_v2(const cyp_graph *g);

}  // namespace cyp
