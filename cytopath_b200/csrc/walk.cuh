// a2: markov_sim (reference cytopath/markov_chain_sampler.py:35-52) — the step-loop kernel K2.
#pragma once

#include "common.cuh"

namespace cyp {

struct WalkParams {
    const RowInfo *rows;            // [n_cells]
    const void *cdf;                // float / double, sector-aligned rows
    const int32_t *col;             // same offsets as cdf
    const int32_t *roots;           // [n_roots]
    int64_t sims_per_root;
    int64_t walker_begin;           // global id of local walker 0
    int64_t n_walkers;
    int32_t nt;                     // transitions per walker (max_steps - 1)
    uint32_t round;
    uint64_t seed;
    const double *uniforms;         // [n_walkers, nt] or nullptr -> Philox stream
    int32_t *out_seq;               // [n_walkers, ld]
    int64_t ld;                     // row stride of out_seq in elements (>= nt + 1)
    unsigned long long *nocross;    // may be nullptr
    // filter outputs, only read/written by the FILTER modes
    const uint8_t *code8;           // [n_cells] code of the truncated cluster label (n_codes <= 256)
    int32_t min_clusters;
    const int32_t *end_slot;        // [n_cells] end-point slot or -1
    int32_t *out_slot;              // [n_walkers]: -2 fails min_clusters, -1 not terminal, else end slot
    uint64_t *out_hash;             // [n_walkers, 2] 128-bit sequence hash
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

// Chooses lanes/unroll from the mean row length so that one chunk covers a typical row.
WalkVariant pick_walk_variant(const cyp_graph *g, int variant);

// Enqueues K2 on `stream`.  Returns a CYP_* code.
int launch_walk(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream);

}  // namespace cyp
