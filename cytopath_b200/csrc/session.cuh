// Per-device state of one sampling() run (session.cu) and what the multi-device layer (group.cu) shares with it.
#pragma once

#include <atomic>
#include <vector>

#include "walk.cuh"

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
    // index of the kept sequences (accumulation order of :335-340): 16 bytes per sequence, the rows are replayed on demand
    int32_t *d_k_round = nullptr;
    int64_t *d_k_walker = nullptr;
    int32_t *d_k_slot = nullptr;
    int64_t kept_rows = 0, kept_cap = 0;
    // (hash_lo, hash_hi, walker) records of the sequences kept by the last round
    uint64_t *d_records = nullptr;
    int64_t n_records = 0, records_cap = 0;
    int64_t last_row0 = 0;
    uint32_t last_round = 0;
    int64_t chunk_walkers = 0;   // 0 = as many walkers per chunk as the free HBM allows
    // sims_per_root of every round walked so far: a replayed walker starts at roots[walker / sims_per_root of its round]
    std::vector<int64_t> sims_by_round;
    int64_t *d_sims_by_round = nullptr;
    size_t sims_uploaded = 0;
};

namespace cyp {

// Rows of n replay items into d_rows [n, max_steps] on `st` (plain mode of the graph's default step kernel).
int session_replay(const cyp_session *s, int64_t n, const ReplayItem *d_items, int32_t *d_rows, cudaStream_t st);
// Rows one replay batch may hold (bounded by 1 GB and by a quarter of the free HBM).
int64_t session_replay_batch_rows(const cyp_session *s);
// h_counts[slot] = rows of the index segment that end on that end point.
int slot_histogram(const int32_t *d_slot, int64_t n, int32_t n_slots, int64_t *h_counts, cudaStream_t st);
// items[k] = replay item of walker[stride * ids[k]] (ids may be nullptr), round round_of[ids[k]] or `round`
void launch_make_items(int64_t n, const int64_t *d_ids, const uint64_t *d_walker, int walker_stride, const int32_t *d_round_of,
                       uint32_t round, const cyp_session *s, ReplayItem *d_items, cudaStream_t st);

}  // namespace cyp
