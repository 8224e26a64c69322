/* cytopath_b200 — C ABI of the B200-native Markov-chain trajectory sampler.
 *
 * This is the drop-in boundary for ONE path of aron0093/cytopath: `cytopath.sampling`
 * (reference cytopath/markov_chain_sampler.py:89-414).  The reference has no FFI layer
 * (it is pure Python); these entry points are what a ctypes binding inside the
 * reference's `markov_chain_sampler.py` would call instead of its numpy loops.  Each
 * declaration names the reference lines it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative CYP_E_* code;
 *     `cyp_last_error()` returns a thread-local message for the last failure.
 *   - handles (`cyp_graph`, `cyp_session`) own device memory; callers own host buffers.
 *   - pointers named `h_*` are host memory, `d_*` device memory on the handle's device.
 *   - calls are synchronous on return unless they take a `stream` (a cudaStream_t passed
 *     as void*; NULL = the legacy default stream), in which case work is only enqueued.
 *   - there is no CPU implementation behind any of these: without a CUDA device they fail.
 *
 * Uniform stream (frozen; restated in oracle/philox.py and oracle/walk.c)
 *   key = (seed lo, seed hi); ctr = (walker lo, walker hi, t >> 1, round);
 *   x = Philox4x32-10(ctr, key); (a, b) = t even ? (x0, x1) : (x2, x3);
 *   u = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53.
 *   `walker` is the row of the round's sample matrix (root-major: walker / sims_per_root
 *   is the index into root_cells, like markov_chain_sampler.py:314-317), `t` the 0-based
 *   transition, `round` the 0-based sampling round.  Results are therefore independent of
 *   how walkers are sharded over launches, chunks or GPUs.
 */
#ifndef CYTOPATH_B200_H
#define CYTOPATH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYP_VERSION 100              /* 0.1.0 */

#define CYP_OK 0
#define CYP_E_INVALID (-1)           /* bad argument */
#define CYP_E_CUDA (-2)              /* CUDA runtime error (message has the detail) */
#define CYP_E_NOMEM (-3)             /* device memory would be exceeded */
#define CYP_E_NOCROSSING (-4)        /* some u exceeded its row's last CDF value: the reference
                                        raises IndexError at markov_chain_sampler.py:49 */

/* CDF recipes (markov_chain_sampler.py:18-19,:23,:28,:31) */
#define CYP_CDF_REFERENCE 0          /* float32(cumsum_f64) rounded to 3 decimals: bit-exact with the reference */
#define CYP_CDF_EXACT 1              /* float64 sequential cumsum (north-star mode) */

typedef struct cyp_graph cyp_graph;
typedef struct cyp_session cyp_session;

int cyp_version(void);
const char *cyp_last_error(void);
int cyp_device_count(int *n);
/* Device memory is recycled across calls (CUDA stream-ordered pool + an exact-size cache of large blocks): this
 * returns all of it that is not in use to the driver.  device < 0: every device. */
int cyp_release_cached_memory(int device);

/* ---- a1: matrix_prep (markov_chain_sampler.py:11-31) -------------------------------
 * Uploads a canonical CSR (columns ascending, no duplicates, no explicit zeros: what
 * `sparse.csr_matrix(dense)` at :14 holds) and builds the per-row CDF on the device with
 * a segmented prefix-sum kernel that adds strictly left to right in float64, like
 * np.cumsum (:19).  The zero-padded ELL arrays of :22-28 are not reproduced: padding
 * never changes the selection. */
int cyp_graph_create(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                     const int32_t *h_col, const double *h_val, int cdf_mode, cyp_graph **out);
/* Same with float32 values (what scvelo stores): widened exactly on the device, so the result is
 * identical to passing np.float64(values) (:168) while the upload is a third smaller. */
int cyp_graph_create_f32(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                         const int32_t *h_col, const float *h_val, int cdf_mode, cyp_graph **out);
/* Smallest and largest row sum (float64, entries added left to right): the caller applies the
 * reference's normalisation check `1 - 1e-3 <= sum <= 1 + 1e-3` (:182-187) to them. */
int cyp_graph_row_sum_range(const cyp_graph *g, double *min_sum, double *max_sum);
/* What K1 saw in the CSR it was given: are every row's columns strictly increasing (sorted, no duplicates), and is there
 * no explicit zero?  The reference samples from the canonical form (it densifies and re-sparsifies, :168-170, :14); a
 * caller that holds a CSR of unknown form builds the graph from it as it is and only canonicalises (and rebuilds) when
 * one of the two comes back 0 -- the check costs the host nothing. */
int cyp_graph_canonical(const cyp_graph *g, int *columns_strictly_increasing, int *no_explicit_zeros);
int cyp_graph_destroy(cyp_graph *g);
int cyp_graph_info(const cyp_graph *g, int64_t *n_cells, int64_t *nnz, int32_t *max_row_nnz,
                   int *cdf_mode, int *device, int64_t *device_bytes);
/* CDF values in CSR order: float[nnz] (reference mode) or double[nnz] (exact mode). */
int cyp_graph_fetch_cdf(const cyp_graph *g, void *h_cdf);
/* Fused-row layout of the default step kernel: *fused = 1 when the graph has it (else the
 * other outputs are 0), bytes of the row blocks, bytes of the 4-byte target words, largest block in bytes. */
int cyp_graph_layout(const cyp_graph *g, int *fused, int64_t *row_block_bytes, int64_t *target_word_bytes,
                     int32_t *max_block_bytes);
/* Test / inspection hook: copies the fused rows to the host.  h_rows [row_block_bytes], h_words [n_cells] (row word
 * of every cell: sectors << 27 | first sector), h_targets [target_word_bytes / 4] (row words of all targets, rows
 * padded to 4 words).  Fails with CYP_E_INVALID when the graph has no fused rows. */
int cyp_graph_fetch_fused(const cyp_graph *g, uint8_t *h_rows, uint32_t *h_words, uint32_t *h_targets);
/* Placement calibration (graphs whose fused rows exceed half the L2): where the rows land in physical memory decides which
 * L2 slices serve the hot rows, and the same kernel runs 20 % apart on identical arrays; before the first large walk on such
 * a graph the library times a short walk of the caller's own walkers on the placement it got and on three fresh allocations
 * and keeps the fastest.  Probe times before / after (0: not calibrated yet) and how often the arrays were moved. */
int cyp_graph_placement(const cyp_graph *g, float *probe_ms_before, float *probe_ms_after, int *moves);
/* Milliseconds the K1 prefix-sum kernel took when the graph was created (CUDA events). */
int cyp_graph_build_ms(const cyp_graph *g, float *ms);

/* ---- a5: iterate_state_probability (markov_chain_sampler.py:62-87) --------------------
 * max_iter iterations of p <- p @ T from h_init, with T given as CSR exactly as stored in
 * adata.uns[matrix_key] (any column order; it is transposed on the device);
 * h_eps_history[i] = cosine distance between p after iteration i+1 and h_stationary (:73-76).
 * The caller derives max_steps with check_convergence_criteria (:54-60).
 * h_final_state (may be NULL) receives p after the last iteration; h_history (may be NULL) is double [max_iter, n_cells]
 * and receives p BEFORE every iteration (the `state_history` the reference returns, :65,:71). */
int cyp_power_iteration(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                        const int32_t *h_col, const double *h_val, const double *h_init,
                        const double *h_stationary, int32_t max_iter, double *h_eps_history,
                        double *h_final_state, double *h_history);

/* Same on the matrix a graph already holds in HBM (no second upload): valid when adata.uns[matrix_key] is the canonical
 * matrix the graph was built from. */
int cyp_graph_power_iteration(const cyp_graph *g, const double *h_init, const double *h_stationary, int32_t max_iter,
                              double *h_eps_history, double *h_final_state);

/* ---- a2: markov_sim (markov_chain_sampler.py:35-52) --------------------------------
 * Walkers [walker_begin, walker_begin + n_walkers) of one round, each starting at
 * root_cells[walker / sims_per_root] and taking n_transitions = max_steps - 1 steps with
 * next = col[first j with u <= cdf[j]] (:48-49, ties go to the bin).  `uniforms`, when not
 * NULL, is double [n_walkers, n_transitions] and replaces the Philox stream (the hook the
 * parity tests use to drive the reference's numpy rule with identical numbers).
 * out_seq is int32 [n_walkers, n_transitions + 1], column 0 = root (:41).
 * *n_nocrossing counts walker-steps with no bin (the reference raises IndexError there);
 * the host variant returns CYP_E_NOCROSSING if it is non-zero but still fills out_seq. */
int cyp_walk(const cyp_graph *g, const int32_t *h_root_cells, int64_t n_roots, int64_t sims_per_root,
             int64_t walker_begin, int64_t n_walkers, int32_t n_transitions, uint64_t seed, uint32_t round,
             const double *h_uniforms, int32_t *h_out_seq, int64_t *n_nocrossing);
/* Same, everything resident in HBM, enqueued on `stream`.  d_nocrossing (may be NULL) is
 * incremented atomically.  `variant` selects the step kernel: 0 = auto (fused rows -- 12-bit CDF
 * codes with the targets of the most probable entries inlined -- when the graph has them and 97 %
 * of its rows fit 256 bytes, else staged 16-bit CDF codes + edge records); 9500 + T/32 = fused rows
 * with T threads per CTA (9600 + T/32: two CTAs per SM; 9700 + T/32: slot sized for the longest
 * row, bulk copies above 256 bytes); 9400 + T/32 = two walkers per thread on 16-bit codes;
 * 9300 + T/32 / 9200 + T/32 = staged 16-bit codes, cooperative cp.async / per-thread bulk copies;
 * 9000 + T/32 = the CDF rows as stored, bulk copies; L*100 + U = register flavour with L lanes per
 * walker and U 16-byte loads per lane per chunk.  All flavours return identical sequences. */
int cyp_walk_device(const cyp_graph *g, const int32_t *d_root_cells, int64_t n_roots, int64_t sims_per_root,
                    int64_t walker_begin, int64_t n_walkers, int32_t n_transitions, uint64_t seed,
                    uint32_t round, const double *d_uniforms, int32_t *d_out_seq,
                    unsigned long long *d_nocrossing, int variant, void *stream);
/* Name of the kernel variant `cyp_walk_device` would pick / picked for this graph. */
const char *cyp_walk_variant_name(const cyp_graph *g, int variant);

/* score = sum_t -log(float32(T[c_t, c_t+1])) per row, float32 (markov_chain_sampler.py:50,:52),
 * evaluated on the device for the rows given (numpy's pairwise summation order). */
int cyp_transition_scores(const cyp_graph *g, const int32_t *h_seq, int64_t n_rows, int32_t n_steps,
                          float *h_score);

/* ---- a3: one sampling round + filters (markov_chain_sampler.py:299-354) -------------
 * A session fixes the per-run inputs of the round loop: root cells, the per-cell code of
 * the 8-character-truncated cluster label (:39,:51) with min_clusters (:322), the
 * end-point slot of every cell (-1 = not an end point; :343-349), max_steps, `unique`
 * (:329-333) and the seed.  Each round walks its share of the round's walkers, keeps those
 * that (1) visit >= min_clusters distinct labels and (2) end on an end point, removes
 * duplicates keeping the first occurrence (lowest walker id, = np.unique(..., return_index)
 * + sorted at :330-331), and appends the survivors, in walker order, to a device-resident
 * store.  Sequences that do not end on an end point can never be selected by :385-403 and
 * are dropped on the device. */
int cyp_session_create(const cyp_graph *g, const int32_t *h_root_cells, int64_t n_roots,
                       const int32_t *h_cell_code, int32_t n_codes, int32_t min_clusters,
                       const int32_t *h_end_slot_of_cell, int32_t n_end_slots,
                       int32_t max_steps, int unique, uint64_t seed, cyp_session **out);
int cyp_session_destroy(cyp_session *s);

typedef struct cyp_round_stats {
    int64_t n_walkers;        /* walkers simulated by this call */
    int64_t n_pass_clusters;  /* of those, visiting >= min_clusters labels */
    int64_t n_terminal;       /* of those, also ending on an end point */
    int64_t n_kept;           /* after de-duplication: rows appended to the store */
    int64_t n_nocrossing;     /* walker-steps with no CDF crossing */
    int64_t store_rows;       /* total rows in the store after this call */
    int32_t dedup_passes;     /* hash passes used (1 unless a 128-bit hash collided) */
    float ms_walk;            /* step kernel, CUDA events */
    float ms_filter;          /* filter + dedup + compaction kernels */
    int32_t n_chunks;         /* pieces the round was walked in (1 unless it did not fit the free HBM) */
    int64_t n_replayed;       /* walkers walked a second time to compare sequences with equal hashes */
} cyp_round_stats;

/* Walkers [walker_begin, walker_end) of round `round` with sims_per_root walkers per root. */
int cyp_session_round(cyp_session *s, uint32_t round, int64_t sims_per_root, int64_t walker_begin,
                      int64_t walker_end, cyp_round_stats *stats);
/* Rows kept so far, then per kept row: round, walker id, end-point slot (store order =
 * round-major, walker order within a round, i.e. the accumulation order of :335-340). */
int cyp_session_rows(const cyp_session *s, int64_t *n_rows);
int cyp_session_fetch_index(const cyp_session *s, int64_t row_begin, int64_t n_rows, int32_t *h_round,
                            int64_t *h_walker, int32_t *h_end_slot);
/* h_counts [n_end_slots]: how many of the kept rows [row_begin, row_begin + n_rows) end on each end point (:341-354). */
int cyp_session_slot_counts(const cyp_session *s, int64_t row_begin, int64_t n_rows, int64_t *h_counts);
/* Kept rows (by row index) as int32 [n_rows, max_steps] on the host.  The session stores 16 bytes per kept sequence
 * (round, walker, end-point slot), not the sequence: the rows are produced by walking those walkers again -- the
 * uniform stream is keyed by (walker, step, round), so the replay is bit-identical. */
int cyp_session_fetch_rows(const cyp_session *s, const int64_t *h_rows, int64_t n_rows, int32_t *h_seq);

/* Multi-GPU de-duplication support (walkers sharded by contiguous id blocks).  After a
 * round, each rank exports the (hash_lo, hash_hi, walker) records of the rows it just
 * appended; the ranks all-gather them (NCCL) and each rank drops the rows whose sequence
 * was also produced by a lower walker id on another rank. */
int cyp_session_round_records(const cyp_session *s, uint64_t **d_records, int64_t *n_records);
int cyp_session_round_prune(cyp_session *s, const uint64_t *d_all_records, int64_t n_all, int64_t *n_dropped);

/* ---- a6 / multi-GPU: the joblib fan-out over root cells (markov_chain_sampler.py:306-319) as a group of devices ----
 * A group shares one sampling() run: the walkers of every round are cut into contiguous id blocks, one per rank, the
 * graph is replicated, and the ranks exchange 24-byte records of kept sequences over NCCL (all-gather), never rows.
 * Two ways to form one: in ONE process that drives n devices (cyp_group_create_local: ncclCommInitAll, one host thread
 * per device inside each call -- the boundary is a notebook function call), or one process per GPU
 * (cyp_group_unique_id on rank 0, the 128 bytes sent to the others by any means, cyp_group_create_rank everywhere).
 * A group of one device needs no NCCL and does not load it.  All ranks make the same calls with the same arguments. */
typedef struct cyp_group cyp_group;
int cyp_group_create_local(int n_devices, const int *devices, cyp_group **out);
int cyp_group_unique_id(uint8_t *id128);
int cyp_group_create_rank(const uint8_t *id128, int rank, int world, int device, cyp_group **out);
int cyp_group_destroy(cyp_group *g);
/* Frees the group's graphs, sessions and index (device memory) but keeps its communicators for the next run. */
int cyp_group_release(cyp_group *g);
int cyp_group_info(const cyp_group *g, int *world, int *n_local, int *first_rank, int *nccl_version);
/* The graph / session of the i-th local device (owned by the group). */
cyp_graph *cyp_group_graph(const cyp_group *g, int local_index);
cyp_session *cyp_group_session(const cyp_group *g, int local_index);
/* a1 for a group: the rank `root_rank` builds the graph from its host CSR (cyp_graph_create / _f32; the other ranks
 * may pass NULL arrays) and its device arrays are broadcast to every other device (ncclBroadcast over NVLink) instead
 * of one host upload per GPU.  ms_build / ms_broadcast (may be NULL): host wall time of the two phases. */
int cyp_group_graph_create(cyp_group *g, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, const int32_t *h_col,
                           const void *h_val, int val_is_f32, int cdf_mode, int root_rank, float *ms_build, float *ms_broadcast);
/* a1 for a group when EVERY rank holds the host CSR (each torchrun process has the AnnData; a single process with N
 * devices has the one host copy): every rank uploads 1/world of the column and value arrays over its own PCIe link, an
 * in-place ncclAllGather over NVLink completes them on every device, and each device runs K1 / K1b itself -- the host
 * upload, the longest part of a graph build, shrinks with the number of GPUs.  The ranks first compare (n_cells, nnz,
 * dtype, cdf_mode, a sampled checksum of the three arrays): CYP_E_INVALID on every rank if they differ.
 * ms_upload / ms_gather: device time of the first local device's own upload / of the all-gather; ms_total: host wall. */
int cyp_group_graph_create_sharded(cyp_group *g, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, const int32_t *h_col,
                                   const void *h_val, int val_is_f32, int cdf_mode, float *ms_upload, float *ms_gather,
                                   float *ms_total);
/* cyp_session_create on every local device (same arguments on all ranks). */
int cyp_group_session_create(cyp_group *g, const int32_t *h_root_cells, int64_t n_roots, const int32_t *h_cell_code,
                             int32_t n_codes, int32_t min_clusters, const int32_t *h_end_slot_of_cell, int32_t n_end_slots,
                             int32_t max_steps, int unique, uint64_t seed);

typedef struct cyp_group_round_stats {
    int64_t n_walkers, n_pass_clusters, n_terminal;   /* sums over all ranks */
    int64_t n_kept;            /* sequences the round added to the global index (after the cross-rank first-occurrence prune) */
    int64_t n_nocrossing, n_replayed;
    int64_t total_rows;        /* global index size after this round */
    int64_t exchange_bytes;    /* bytes every rank received through the round's all-gathers */
    float ms_walk, ms_filter;  /* max over ranks of the device times of K2 / K3 */
    float ms_local;            /* host wall time of the local rounds */
    float ms_exchange;         /* host wall time of the cross-rank part: all-gathers, prune, index update */
    int32_t n_chunks, dedup_passes;
} cyp_group_round_stats;

/* One sampling round (:299-354) over the whole group: rank r walks walkers [total r / world, total (r+1) / world) of
 * the round's n_roots * sims_per_root, the first occurrence of every sequence survives group-wide, and every rank ends
 * up with the global index of kept sequences (round, walker, end-point slot) in accumulation order (:335-340).
 * h_slot_counts (may be NULL): int64 [n_end_slots], how many of the sequences this round added end on each end point. */
int cyp_group_round(cyp_group *g, uint32_t round, int64_t sims_per_root, int64_t *h_slot_counts, cyp_group_round_stats *stats);
int cyp_group_rows(const cyp_group *g, int64_t *n_rows);
int cyp_group_fetch_index(const cyp_group *g, int64_t row_begin, int64_t n_rows, int32_t *h_round, int64_t *h_walker, int32_t *h_slot);
/* Order-sensitive 128-bit digest of the global index: equal on every rank and for every way of cutting the rounds. */
int cyp_group_digest(const cyp_group *g, uint64_t *digest2);

typedef struct cyp_group_select_stats {
    int64_t gather_bytes;      /* bytes of the all-gather of the selected sequences */
    float ms_locate_replay, ms_gather, ms_fetch;
} cyp_group_select_stats;

/* The final draw's rows (:395-403).  Pick k is the h_offset[k]-th kept sequence, in accumulation order, that ends on
 * end point h_slot[k] (:386-394).  The picks are split over the ranks, each rank replays its share, the int32
 * sequences are all-gathered, and h_seq [n_pick, max_steps] (+ h_score [n_pick], the transition scores of :50,:52, may
 * be NULL) are returned on every process. */
int cyp_group_select(cyp_group *g, int64_t n_pick, const int32_t *h_slot, const int64_t *h_offset, int32_t *h_seq, float *h_score,
                     cyp_group_select_stats *stats);
int cyp_group_set_chunk_walkers(cyp_group *g, int64_t walkers);
/* Host bytes of rank `root_rank` to every process of the group (e.g. the indices rank 0 drew with np.random, :395). */
int cyp_group_broadcast_bytes(cyp_group *g, void *h_buf, int64_t n_bytes, int root_rank);

/* Walkers per chunk of a round (0 = as many as the free HBM allows).  A round larger than a chunk is
 * walked piecewise; the first-occurrence de-duplication is then completed across the chunks, so the
 * result does not depend on the chunk size. */
int cyp_session_set_chunk_walkers(cyp_session *s, int64_t walkers);

/* ---- f2: distance_matrix (trajectory_estimation/sample_clustering.py:57-71) -------------
 * Symmetric [n_chains, n_chains] matrix of Hausdorff distances between chains of chain_len points in
 * `dim` dimensions (h_coords is double [n_chains, chain_len, dim], what coordinate_assigner :38-55
 * returns).  metric: 0 euclidean (the reference's default, :135), 1 manhattan, 2 chebyshev, 3 cosine,
 * as in the `hausdorff` package the reference calls.  kernel_ms (may be NULL): device time. */
int cyp_hausdorff_matrix(int device, const double *h_coords, int64_t n_chains, int32_t chain_len, int32_t dim,
                         int metric, double *h_out, float *kernel_ms);

/* Measurement aid (bench.py roofline.gather_ceiling_frac): gathers/s the device sustains for independent random reads
 * of gather_bytes (32 / 64 / 128 / 256) contiguous bytes out of a working set of working_set_bytes -- the step kernel's
 * access pattern without the dependency between a walker's steps. */
int cyp_gather_ceiling(int device, int64_t working_set_bytes, int32_t gather_bytes, double *gathers_per_s);

/* test hook: keep only the low `bits` bits of the 128-bit sequence hash (forces collisions) */
int cyp_session_set_hash_bits(cyp_session *s, int bits);

/* f4 -- directionality_score (reference cytopath/trajectory_estimation/cyto_path.py:155-243).
 * Masked velocity graph as CSR over n_cells (h_row_ptr int64 [n_cells+1], h_col, h_w = expm1(weight) as float64);
 * h_pos [n_cells, dims] cell coordinates; h_dirs [n_dirs, dims] trajectory directions; one instance per (trajectory,
 * step, cell): h_cell, h_fwd / h_bwd = index of the forward / backward direction or -1 (first step: forward only,
 * last step: backward only, else max of both).  h_score [n_inst]: mean over the cell's entries of
 * cos(direction, x_target - x_cell) * w; nan for a cell without entries.  kernel_ms may be NULL. */
int cyp_directionality_scores(int device, int64_t n_cells, int32_t dims, const double *h_pos, const int64_t *h_row_ptr,
                              const int32_t *h_col, const double *h_w, int64_t n_dirs, const double *h_dirs,
                              int64_t n_inst, const int32_t *h_cell, const int32_t *h_fwd, const int32_t *h_bwd,
                              double *h_score, float *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
