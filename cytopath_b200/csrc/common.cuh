// Shared declarations for the cytopath_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/cytopath_b200.h"

namespace cyp {

// ---- error plumbing ---------------------------------------------------------------
int fail(int code, const std::string &msg);

#define CYP_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::cyp::fail(CYP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define CYP_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) return ::cyp::fail(CYP_E_INVALID, (msg));   \
    } while (0)

// Device memory comes from the CUDA stream-ordered pool of the current device with its release
// threshold lifted, so create/destroy cycles (one graph + session per sampling() call) reuse
// memory instead of paying cudaMalloc/cudaFree each time.  All calls use the legacy default stream.
cudaError_t dev_alloc(void **p, size_t bytes);
void dev_free(void *p);
size_t dev_cached_bytes(int dev);     // bytes of freed >= 64 MB blocks kept for exact-size reuse (graph.cu)
void dev_release_cached(int dev);     // dev < 0: all devices
template <typename T>
inline cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc(reinterpret_cast<void **>(p), bytes); }

// SMs of the current device (cached per host thread and device; the library may drive several devices from one process)
int current_sm_count();

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != prev && prev >= 0) cudaSetDevice(prev);
    }
};

// ---- Philox4x32-10 and the frozen walker uniform stream (include/cytopath_b200.h) ----
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += kPhiloxW0;
        k.y += kPhiloxW1;
    }
    return c;
}

__device__ __forceinline__ double bits_to_unit_double(uint32_t a, uint32_t b) {
    // numpy's 53-bit construction: ((a >> 5) * 2^26 + (b >> 6)) / 2^53, exact in double
    uint64_t mant = (static_cast<uint64_t>(a >> 5) << 26) + static_cast<uint64_t>(b >> 6);
    return static_cast<double>(mant) * (1.0 / 9007199254740992.0);
}

// Both uniforms of Philox call `pair` (transitions 2*pair and 2*pair+1) of one walker.
__device__ __forceinline__ void walker_uniform_pair(uint64_t seed, uint32_t round, uint64_t walker, uint32_t pair,
                                                    double &u_even, double &u_odd) {
    uint4 x = philox4x32_10(make_uint4(static_cast<uint32_t>(walker), static_cast<uint32_t>(walker >> 32), pair, round),
                            make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    u_even = bits_to_unit_double(x.x, x.y);
    u_odd = bits_to_unit_double(x.z, x.w);
}

// ---- device-resident graph ----------------------------------------------------------
// Rows are stored aligned: every row starts at a multiple of kRowAlign = 16 elements in every
// per-edge array (CDF float64/float32, 16-bit CDF codes, edge records, col, val), i.e. on a
// 32-byte sector boundary even for the 2-byte codes.  A row of k entries touches
// ceil(k*size/32) sectors, 16-byte vector loads are always legal, and a row is a legal
// cp.async.bulk source.  One offset (in units of 16 elements) addresses all arrays.
constexpr int kRowAlign = 16;
struct RowInfo {          // read once per walker, for its root cell
    uint32_t off_units;   // row start / kRowAlign
    uint32_t len;         // entries in the row
};
// One 16-byte record per edge, at the same offset as the edge's CDF entry: the target cell AND
// the target's row descriptor.  The walker therefore needs no separate RowInfo lookup between
// steps: a step is two dependent memory phases (CDF row -> edge record) instead of three.
struct EdgeRec {
    int32_t col;          // target cell (-1 in padding)
    uint32_t off_units;   // target's row start / kRowAlign
    uint32_t len;         // target's row length
    uint32_t code;        // target's cluster-label code for the session that last labelled the graph
};

// ---- fused rows ("v2" layout, what the default step kernel walk_v2.cu reads) --------------------
// One block of 32-byte sectors per cell holding everything a step needs except, sometimes, the target:
//   +0  int32  cell            the cell id (trajectory column; the walker learns it when the row arrives)
//   +4  uint16 len             entries in the row
//   +6  uint8  label           code of the cell's truncated cluster label (written per session)
//   +7  uint8  m               number of inlined targets
//   +8  uint64 inline_mask     bit j set: the target of entry j is inlined (only entries j < 64 qualify)
//   +16 uint32 edge_base       the row's targets start at edge4[edge_base * 4]
//   +20 uint8  hi[len]         upper 8 bits of the 12-bit CDF codes
//   ..  uint8  lo[(len+1)/2]   lower 4 bits, two per byte (entry j in nibble j & 1 of byte j >> 1)
//   ..  (pad to 4 bytes) uint32 inl[m]   targets of the m most probable entries, in entry order
// A target is a ROW WORD: (sectors of the target's block) << kWordOffBits | first sector of the block --
// exactly what a warp needs to start copying the next row, so for the inlined entries (the most
// probable ones: the slack of the block's last sector is filled with them) a step is ONE gather.
// 12-bit codes: exact mode q = min(4095, floor(cdf * 4096)) (undecided compares are settled against the
// float64 CDF, as for the 16-bit codes); reference mode q = 4 m with cdf = fl32(m / 1000), lossless.
constexpr int kWordOffBits = 27;
constexpr uint32_t kWordOffMask = (1u << kWordOffBits) - 1u;
constexpr int kV2Header = 20;
constexpr int kV2MaxSectors = 31;          // 5 bits of a row word (saturating)
__host__ __device__ inline int v2_used_bytes(int len) { return (kV2Header + len + (len + 1) / 2 + 3) & ~3; }
// sectors of a row's block: what it uses, plus `extra` sectors of inlined targets while the block stays below
// kV2MaxSectors (longer blocks exist -- the row word's sector field saturates there -- but get no extra)
__host__ __device__ inline int v2_block_sectors(int len, int extra) {
    const int n = (v2_used_bytes(len) + 31) / 32;
    return (len > 0 && n + extra <= kV2MaxSectors) ? n + extra : n;
}

}  // namespace cyp

struct cyp_graph {
    int device = 0;
    int cdf_mode = 0;
    int64_t n = 0;
    int64_t nnz = 0;
    int64_t padded = 0;        // elements in the padded arrays
    int32_t max_len = 0;
    double mean_len = 0.0;
    int align_elems = cyp::kRowAlign;
    int q16_ok = 0;            // the 16-bit CDF codes are usable (all CDF values in range, CDF non-decreasing)
    int monotone = 1;          // every probability >= 0, so each row's CDF is non-decreasing (binary search is exact)
    int noncanonical = 0;      // what K1 saw in the CSR it was given: bit 0 a row's columns are not strictly increasing, bit 1 an explicit zero
    cyp::RowInfo *d_row = nullptr;   // [n]
    void *d_cdf = nullptr;           // float or double [padded]
    uint16_t *d_q16 = nullptr;       // [padded] 16-bit CDF codes (see walk_tma.cu)
    int32_t *d_col = nullptr;        // [padded] column indices (transition scores look edges up here)
    cyp::EdgeRec *d_edge = nullptr;  // [padded] what the step kernel reads after the first-crossing search
    double *d_val = nullptr;         // [padded] raw probabilities (transition scores)
    int64_t *d_row_ptr = nullptr;    // [n+1] original CSR offsets (fetch_cdf un-pads with it)
    int64_t device_bytes = 0;
    float build_ms = 0.f;
    double row_sum_min = 0.0, row_sum_max = 0.0;     // range of the float64 row sums (normalisation check)
    mutable uint64_t labels_owner = 0;   // id of the session whose label codes are in d_edge[].code (0 = none)
    // fused rows (see above); v2_ok = 0 when a row needs more than kV2MaxSectors sectors, the codes are unusable
    // (q16_ok / monotone) or the blocks exceed 2^27 sectors
    int v2_ok = 0;
    int v2_max_sectors = 0;              // largest block, in sectors
    int v2_uniform = 0;                  // at least 90 % of the blocks fill the slot: the step kernel copies whole slots without per-chunk predicates
    int v2_slot_sectors = 0;             // shared-memory slot of the step kernel: covers >= 97 % of the rows (power of two, <= 8
                                         // sectors) or, if no such size does, the largest block; longer rows are searched in HBM
    int64_t v2_sectors = 0;              // total sectors of d_rows2
    unsigned char *d_rows2 = nullptr;    // [v2_sectors * 32]
    unsigned char *d_rows2_base = nullptr;   // the allocation d_rows2 lies in (d_rows2 = base + placement offset)
    int64_t rows2_bytes = 0;                 // bytes of d_rows2 (blocks + slack for whole-slot copies past the last block)
    uint32_t *d_edge4 = nullptr;         // [v2_edge_units * 4] row words of all targets, rows padded to 4 words
    int64_t v2_edge_units = 0;
    uint32_t *d_word = nullptr;          // [n] row word of every cell (the roots start from it)
    mutable uint64_t labels2_owner = 0;  // id of the session whose label codes are in the fused rows' headers
    float placement_ms_before = 0.f, placement_ms_after = 0.f;     // probe times of calibrate_placement (0: not calibrated)
    int placement_moves = 0;
    mutable int placement_state = 0;     // large walks seen so far (saturating at 2): the second one calibrates the placement of the fused rows (placement.cu)
};

namespace cyp {
// K0 (layout.cu): the row layout on the device -- padded row offsets, fused-row words, target bases -- from the uploaded
// CSR row offsets; the host only reads this summary.  Two candidates of "extra inline sectors per row" are measured at once.
struct RowLayoutSummary {
    int64_t units16;                             // sum of ceil(len / kRowAlign): the padded arrays hold units16 * kRowAlign elements
    int64_t sectors[2];                          // total sectors of the fused rows, per candidate
    int64_t units4;                              // sum of ceil(len / 4): 4-word units of the target array
    int64_t hist[2][kV2MaxSectors + 1];          // rows per block size (saturating), per candidate
    int32_t max_len;
    int32_t max_sectors[2];
    int32_t bad;                                 // a row length below 0 or above 2^30
};
struct RowLayoutScratch {
    int64_t n = 0;
    int extra[2] = {0, 0};
    int64_t *d_scan = nullptr;                   // 4 x (n + 1): exclusive sums of units16, sectors[0], sectors[1], units4
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    RowLayoutSummary *d_sum = nullptr;
};
cudaError_t layout_measure(int64_t n, int align_elems, const int64_t *d_row_ptr, int extra0, int extra1, RowLayoutScratch *s,
                           cudaStream_t stream, RowLayoutSummary *h_out);
cudaError_t layout_write(const RowLayoutScratch *s, const int64_t *d_row_ptr, int which, RowInfo *d_row, uint32_t *d_word, uint32_t *d_ebase,
                         cudaStream_t stream);
void layout_release(RowLayoutScratch *s);
// K1b (rows_v2.cu): the fused rows, built from K1's output chunk by chunk; v2_finish sets g->v2_ok.
struct V2Build {
    bool planned = false;
    int which = 0;                           // the candidate of RowLayoutSummary the plan took
    int64_t sectors = 0, units = 0;
    int max_sectors = 0, slot_sectors = 0, extra = 0, uniform = 0;
    size_t rows_bytes = 0;
    uint32_t *d_ebase = nullptr;             // per cell: start of its targets in edge4 (units of 4 words)
    int *d_bad = nullptr;
};
// the two candidates of extra inline sectors layout_measure should look at for this graph (CYP_V2_EXTRA pins both)
void v2_candidates(int *extra0, int *extra1);
bool v2_plan(const cyp_graph *g, const RowLayoutSummary &rows, const int extra[2], V2Build *b);
cudaError_t v2_alloc(cyp_graph *g, V2Build *b);
cudaError_t v2_launch(cyp_graph *g, V2Build *b, int64_t row_begin, int64_t row_end, cudaStream_t stream);
cudaError_t v2_finish(cyp_graph *g, V2Build *b, bool codes_ok);
void v2_release(cyp_graph *g, V2Build *b);
// graph.cu, for the multi-device layer (group.cu): a graph built on one device is replicated by broadcasting its
// device arrays over NVLink instead of uploading the CSR from the host once per GPU.
struct GraphWire {                 // the host-side fields of a cyp_graph, as they travel to the other ranks
    int32_t cdf_mode, max_len, q16_ok, monotone, v2_ok, v2_max_sectors, v2_slot_sectors, v2_uniform, has_edge_records, noncanonical;
    int64_t n, nnz, padded, v2_sectors, v2_edge_units, rows2_bytes;
    double mean_len, row_sum_min, row_sum_max;
    float build_ms, pad2_;
};
GraphWire graph_wire(const cyp_graph *g);
// allocates (does not fill) the device arrays of a graph described by `w` on `device`
int graph_alloc_like(const GraphWire &w, int device, cyp_graph **out);
// (pointer, bytes) of every device array of the graph, in an order that only depends on the wire fields
void graph_device_arrays(const cyp_graph *g, std::vector<std::pair<void *, size_t>> *out);
// The raw column / value arrays of a CSR that are already on the device (or on their way there: `ready` fires on the
// stream that fills them).  What the group's sharded upload hands to the graph build, which takes ownership of both arrays.
struct RawCsr {
    int32_t *d_col = nullptr;
    void *d_val = nullptr;          // float64 or float32
    cudaEvent_t ready = nullptr;
};
// cyp_graph_create with the column / value arrays taken from `raw` instead of the host; h_row_ptr is still a host array
int graph_create_raw(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, bool val_f32, int cdf_mode, const RawCsr &raw, cyp_graph **out);

// graph.cu: builds d_q16 and d_edge (the arrays of the step kernels other than walk_v2.cu) if they do not exist yet.
// Synchronous, default stream; call before launching one of those kernels.
int ensure_edge_records(const cyp_graph *g);
}
