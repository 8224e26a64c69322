// K2, staged-row flavour (default): the Markov-chain step loop (reference
// cytopath/markov_chain_sampler.py:45-51) with one walker per THREAD, one persistent CTA per SM, and
// the current row of every walker staged in a private shared-memory slot.
//
// Why: the register flavour (walk.cu) is latency bound — a walker-step is a chain of dependent loads
// and the row bytes in flight are limited by registers (ncu: 85 % long-scoreboard stalls, 35 % DRAM
// utilisation at 40 warps/SM).  Here a row in flight costs shared memory instead: a 512..1024-thread
// CTA keeps 512..1024 rows in flight per SM.  Per step a thread
//   1. gets its row into the slot: either its own cp.async.bulk (UBLKCP, completion on the warp's
//      mbarrier with the byte count) or, for small rows, the warp's cooperative 16-byte cp.async copies;
//   2. finds the first crossing `u <= cdf[j]` (ties to the bin, :49) in shared memory: binary search
//      when the row is non-decreasing (all probabilities >= 0), else the literal left-to-right scan;
//      the row is the CDF as stored or its 16-bit codes (exact by construction, see "row formats");
//   3. loads ONE 16-byte edge record: next cell, its row descriptor and its label code (:48,:51), so
//      there is no separate descriptor or label lookup;
//   4. computes Philox once per two steps and stores the trajectory as one int4 per four steps.
// Rows longer than a slot are streamed through it in chunks.
#include <cstdio>
#include <type_traits>

#include "walk.cuh"

namespace cyp {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map); completion is reported to `bar` in bytes
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace

// ---- row formats ------------------------------------------------------------------------------
// What is copied into the slot and how the first crossing `min j : u <= cdf[j]` (:49) is found.
//   RowF64 / RowF32   the CDF itself (exact / reference mode): 8 or 4 bytes per entry.
//   RowQ16Exact       16-bit codes q = floor(cdf * 65536) of the float64 CDF: 2 bytes per entry.
//                     u <= q/65536 proves u <= cdf; u > (q+1)/65536 proves u > cdf; only entries
//                     whose code equals floor(u*65536)-ish (probability ~ k / 65536 per step) are
//                     ambiguous, and those are compared against the float64 value in HBM, so the
//                     selected index is ALWAYS the one the float64 comparison selects.
//   RowQ16Ref         reference mode: the stored float32 value is exactly fl32(m / 1000) with
//                     integer m (np.round(decimals=3), :31), so the code m is lossless; u is mapped
//                     to the smallest m with u <= fl32(m/1000) and the search is an integer compare.
struct RowF64 {
    using elem = double;
    static constexpr const char *name = "f64";
};
struct RowF32 {
    using elem = float;
    static constexpr const char *name = "f32";
};
struct RowQ16Exact {
    using elem = uint16_t;
    static constexpr const char *name = "q16/f64";
};
struct RowQ16Ref {
    using elem = uint16_t;
    static constexpr const char *name = "q16/f32";
};

// first j in [0, m) with slot[j] >= key (slot non-decreasing); m if none
__device__ __forceinline__ int lower_bound_u16(const uint16_t *slot, int m, int key) {
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (static_cast<int>(slot[mid]) >= key) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Returns the first index j in [0, m) of the chunk with u <= cdf[row_off + base + j], or -1.
template <typename Row>
__device__ __forceinline__ int first_crossing(const typename Row::elem *slot, int m, double u, int monotone,
                                              const void *cdf_full, int64_t row_elem0) {
    if constexpr (std::is_same<Row, RowQ16Exact>::value) {
        const int t_def = static_cast<int>(ceil(u * 65536.0));       // q >= t_def      <=>  u <= q / 65536
        int j = lower_bound_u16(slot, m, t_def - 1);                   // q >= t_def - 1  <=>  u <= (q + 1) / 65536
        const double *exact = static_cast<const double *>(cdf_full) + row_elem0;
        while (j < m && static_cast<int>(slot[j]) < t_def) {          // ambiguous entries: decide in float64
            if (u <= __ldg(exact + j)) return j;
            ++j;
        }
        return j < m ? j : -1;
    } else if constexpr (std::is_same<Row, RowQ16Ref>::value) {
        // smallest integer key with u <= (double)fl32(key / 1000)
        int key = static_cast<int>(ceil(u * 1000.0));
        while (key > 0 && u <= static_cast<double>(__fdiv_rn(static_cast<float>(key - 1), 1000.0f))) --key;
        while (u > static_cast<double>(__fdiv_rn(static_cast<float>(key), 1000.0f))) ++key;
        const int j = lower_bound_u16(slot, m, key);
        return j < m ? j : -1;
    } else {
        if (monotone) {
            int lo = 0, hi = m;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (u <= static_cast<double>(slot[mid])) hi = mid; else lo = mid + 1;
            }
            return lo < m ? lo : -1;
        }
        for (int j = 0; j < m; ++j)
            if (u <= static_cast<double>(slot[j])) return j;
        return -1;
    }
}

// COPY selects how a warp's 32 rows reach the slots:
//   kCopyBulk     every lane issues its own cp.async.bulk (UBLKCP).  ptxas serialises the 32 lanes of a
//                 warp (the instruction takes uniform-register operands): ~10 issue slots per row,
//                 independent of the row size.  Right for rows of several hundred bytes.
//   kCopyCpAsync  the warp copies cooperatively with 16-byte cp.async (LDGSTS): slot_bytes/16 lanes per
//                 row, so a 128-byte row of 16-bit codes costs one instruction per 4 rows.  Row
//                 addresses travel by shuffle.  Right for small rows (<= 512 bytes per slot).
constexpr int kCopyBulk = 0, kCopyCpAsync = 1;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

template <typename Row, int MODE, int COPY>
__global__ void __launch_bounds__(1024, 1) walk_tma_kernel(const WalkParams p, const void *__restrict__ row_data,
                                                           const int slot_elems, const int monotone) {
    using Elem = typename Row::elem;
    constexpr int COPY_ALIGN = 32 / sizeof(Elem);        // copies are whole 32-byte sectors
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nthreads = blockDim.x;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    Elem *slot = reinterpret_cast<Elem *>(smem_raw + (((nthreads >> 5) * 8 + 127) & ~127)) + static_cast<size_t>(threadIdx.x) * slot_elems;
    uint64_t *bar = bars + warp;
    if constexpr (COPY == kCopyBulk) {
        if (lane == 0) mbar_init(bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    // cooperative copy geometry: `cpr` 16-byte chunks per slot (a power of two <= 32)
    constexpr int EPC = 16 / sizeof(Elem);               // elements per 16-byte chunk
    const int cpr = slot_elems / EPC;
    const int cpr_shift = 31 - __clz(cpr);
    Elem *const warp_slots = slot - static_cast<size_t>(lane) * slot_elems;

    const Elem *__restrict__ rowsrc = static_cast<const Elem *>(row_data);
    const EdgeRec *__restrict__ edge = p.edge;
    const RowInfo *__restrict__ rows = p.rows;
    const int nt = p.nt;
    const bool vec_store = (p.ld & 3) == 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * nthreads;

    for (int64_t wbase = static_cast<int64_t>(blockIdx.x) * nthreads + warp * 32; wbase < p.n_walkers; wbase += stride) {
        const int64_t w = wbase + lane;
        const bool active = w < p.n_walkers;
        uint64_t gw = static_cast<uint64_t>(p.walker_begin + w);
        uint32_t round = p.round;
        int32_t cur = 0;
        if (active) {
            if (p.replay) {
                const ReplayItem it = p.replay[w];
                gw = static_cast<uint64_t>(it.walker);
                round = it.round;
                cur = it.root;
            } else {
                cur = p.roots[gw / static_cast<uint64_t>(p.sims_per_root)];
            }
        }
        RowInfo ri{0u, 0u};
        if (active) ri = rows[cur];
        int32_t *__restrict__ out_row = (p.out_seq && active) ? p.out_seq + w * p.ld : nullptr;
        const double *__restrict__ urow = (p.uniforms && active) ? p.uniforms + w * static_cast<int64_t>(nt) : nullptr;
        int32_t b0 = cur, b1 = 0, b2 = 0, b3 = 0;         // columns 4q .. 4q+3 of the trajectory
        if (out_row && (nt == 0 || !vec_store)) out_row[0] = cur;
        WalkerDigest<MODE> dg;
        if (active) dg.visit(cur, WalkerDigest<MODE>::MASKW > 0 ? p.code8[cur] : 0u);
        unsigned misses = 0;
        double u_even = 0.0, u_odd = 0.0;

        for (int t = 0; t < nt; ++t) {
            double u;
            if (urow) {
                u = urow[t];
            } else {
                if ((t & 1) == 0) walker_uniform_pair(p.seed, round, gw, static_cast<uint32_t>(t >> 1), u_even, u_odd);
                u = (t & 1) ? u_odd : u_even;
            }
            const int len = static_cast<int>(ri.len);
            const int len_pad = (len + COPY_ALIGN - 1) & ~(COPY_ALIGN - 1);
            const int64_t off = static_cast<int64_t>(ri.off_units) * kRowAlign;
            int found = -1;
            for (int base = 0;; base += slot_elems) {    // warp-uniform: all lanes take every barrier phase
                const bool need = found < 0 && base < len;
                if (!__any_sync(0xffffffffu, need)) break;
                const int n = need ? min(len_pad - base, slot_elems) : 0;
                if constexpr (COPY == kCopyBulk) {
                    if (n > 0) {
                        const uint32_t bytes = static_cast<uint32_t>(n) * sizeof(Elem);
                        mbar_arrive_expect_tx(bar, bytes);
                        bulk_copy_g2s(slot, rowsrc + off + base, bytes, bar);
                    } else {
                        mbar_arrive(bar);
                    }
                    mbar_wait(bar, parity);
                    parity ^= 1u;
                } else {
                    const int my_chunk = lane & (cpr - 1);
                    const unsigned my_units = ri.off_units;
                    for (int i = 0; i < cpr; ++i) {          // instruction i serves walkers i*(32/cpr) .. of the warp
                        const int src_lane = ((i << 5) >> cpr_shift) + (lane >> cpr_shift);
                        const unsigned units = __shfl_sync(0xffffffffu, my_units, src_lane);
                        const int n_w = __shfl_sync(0xffffffffu, n, src_lane);
                        if (my_chunk * EPC < n_w) {
                            Elem *dst = warp_slots + static_cast<size_t>(src_lane) * slot_elems + my_chunk * EPC;
                            const Elem *src = rowsrc + static_cast<int64_t>(units) * kRowAlign + base + my_chunk * EPC;
                            cp_async16(dst, src);
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                }
                if (n > 0) {
                    const int j = first_crossing<Row>(slot, min(n, len - base), u, monotone, p.cdf, off + base);
                    if (j >= 0) found = base + j;
                }
                if constexpr (COPY == kCopyCpAsync) __syncwarp();   // other lanes write this slot in the next copy
            }
            if (active) {
                unsigned code = 0;
                if (found >= 0) {
                    const EdgeRec e = load_edge(edge + off + found);
                    cur = e.col;
                    ri = RowInfo{e.off_units, e.len};
                    code = e.code;
                } else {
                    ++misses;
                    if constexpr (WalkerDigest<MODE>::MASKW > 0) code = p.code8[cur];
                }
                const int c = t + 1;
                if (!out_row) {
                    // no trajectory buffer: digest only
                } else if (vec_store) {
                    const int q = c & 3;
                    if (q == 0) b0 = cur; else if (q == 1) b1 = cur; else if (q == 2) b2 = cur; else b3 = cur;
                    if (q == 3) {
                        *reinterpret_cast<int4 *>(out_row + (c & ~3)) = make_int4(b0, b1, b2, b3);
                    } else if (c == nt) {
                        out_row[c & ~3] = b0;
                        if (q >= 1) out_row[(c & ~3) + 1] = b1;
                        if (q >= 2) out_row[(c & ~3) + 2] = b2;
                    }
                } else {
                    out_row[c] = cur;
                }
                dg.visit(cur, code);
            }
        }
        if (active) {
            if (misses && p.nocross) atomicAdd(p.nocross, static_cast<unsigned long long>(misses));
            dg.finish(p, w, cur);
        }
    }
}

// Tried and dropped: (1) L2 eviction-priority hints (rows evict_last, edge records and trajectory
// stores evict_first through createpolicy descriptors): identical timings on every workload;
// (2) two CTAs of 768 threads per SM for the cp.async flavour (needs <= 42 registers): slower than
// one CTA of 1024 threads (29.9 vs 32.9 G walker-steps/s at C4);
// (3) rotating the 16-byte chunks of every slot per lane (cp.async flavour) and padding bulk slots to an
// odd number of 16-byte units, to spread the binary-search probes over the shared-memory banks: no gain
// (36.8 vs 38.0 G walker-steps/s at C4), so the shared-memory pipe is not the limit either;
// (4) (profiles/r01_sweep4_pingpong_*.log) a ping-pong flavour with two walkers per
// thread sharing one slot, so that one walker's row copy overlaps the other's edge-record load.
// Same throughput as this kernel at equal thread counts (19.1 vs 19.0 G walker-steps/s at C4): the
// memory system, not the per-thread dependency chain, is the limit at this point.

struct TmaConfig {
    int threads;
    int slot_elems;
    size_t smem;
    bool q16;
    bool cpasync;         // cooperative 16-byte cp.async copies instead of per-lane bulk copies
};

// Launch shape.  Measured on B200 (profiles/r01_sweep*.log): throughput rises with the row bytes
// in flight per SM; ~210 KB of slots (512 threads for 416-byte rows) is the best point for
// HBM-bound float64 rows and more than 768 threads brings nothing there; 2-byte rows and graphs
// that fit the L2 are latency bound and take every thread that fits.
// variant: 0 = auto; 9000 + T/32 = CDF rows as stored, bulk copies, T threads; 9200 + T/32 = 16-bit
// codes, bulk copies; 9300 + T/32 = 16-bit codes, cooperative cp.async copies.
static TmaConfig tma_config(const cyp_graph *g, int variant) {
    const bool want_q16 = (variant == 0 || variant >= 9200) && g->q16_ok && g->monotone;
    bool cpasync = want_q16 && (variant == 0 || variant >= 9300);
    if (variant >= 9300) variant -= 100;
    const int elem = want_q16 ? 2 : (g->cdf_mode == CYP_CDF_EXACT) ? 8 : 4;
    const int align = 32 / elem;
    const int max_pad = (g->max_len + align - 1) / align * align;
    const int mean_pad = (static_cast<int>(g->mean_len + 0.999) + align - 1) / align * align;
    const size_t hard_budget = 226 * 1024;
    int threads = (variant > 9200) ? (variant - 9200) * 32 : (variant > 9000 && variant < 9200) ? (variant - 9000) * 32 : 0;
    int slot = (max_pad <= 2 * mean_pad) ? max_pad : 2 * mean_pad;   // ragged graphs: long rows stream in chunks
    if (threads == 0) {
        const bool l2_resident = static_cast<double>(g->padded) * (elem + 16) < 64.0 * 1024 * 1024;
        const size_t budget = (l2_resident || want_q16) ? hard_budget : 216 * 1024;
        const int cap = (l2_resident || want_q16) ? 1024 : 768;
        threads = static_cast<int>(budget / (static_cast<size_t>(slot) * elem)) / 32 * 32;
        if (threads > cap) threads = cap;
        if (threads < 128) threads = 128;
    }
    if (threads > 1024) threads = 1024;
    if (threads < 32) threads = 32;
    size_t cap_elems = (hard_budget - 1024) / (static_cast<size_t>(threads) * elem);
    cap_elems = cap_elems / align * align;
    if (static_cast<size_t>(slot) > cap_elems) slot = static_cast<int>(cap_elems);
    if (slot < align) slot = align;
    if (cpasync) {                                     // slot = power-of-two number of 16-byte chunks, at most 32
        int bytes = 16;
        while (bytes < slot * elem && bytes < 512) bytes <<= 1;
        const bool fits = static_cast<size_t>(bytes) * threads + 1024 <= hard_budget && bytes >= slot * elem;
        const bool small = bytes <= 256 || variant != 0;             // auto: bulk copies win for slots above 256 bytes (C5)
        if (fits && small) slot = bytes / elem;
        else cpasync = false;
    }
    const size_t smem = ((threads / 32 * 8 + 127) & ~127) + static_cast<size_t>(threads) * slot * elem;
    return TmaConfig{threads, slot, smem, want_q16, cpasync};
}

template <typename Row, int MODE, int COPY>
static int launch_tma_one(const cyp_graph *g, const WalkParams &p, const TmaConfig &c, cudaStream_t stream) {
    auto kernel = walk_tma_kernel<Row, MODE, COPY>;
    const int sm_count = current_sm_count();
    if (sm_count <= 0) return fail(CYP_E_CUDA, "launch_walk_tma: no CUDA device");
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute of the function
    CYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c.smem)));
    int ctas_per_sm = 1;
    CYP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, c.threads, c.smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int64_t need = (p.n_walkers + c.threads - 1) / c.threads;
    const int64_t resident = static_cast<int64_t>(sm_count) * ctas_per_sm;                // persistent CTAs, grid-stride over walkers
    const unsigned grid = static_cast<unsigned>(need < resident ? need : resident);
    if (grid == 0) return CYP_OK;
    const void *row_data = std::is_same<typename Row::elem, uint16_t>::value ? static_cast<const void *>(g->d_q16) : g->d_cdf;
    kernel<<<grid, c.threads, c.smem, stream>>>(p, row_data, c.slot_elems, g->monotone);
    CYP_CUDA(cudaGetLastError());
    return CYP_OK;
}

template <typename Row, int COPY>
static int launch_tma_mode(const cyp_graph *g, const WalkParams &p, WalkMode mode, const TmaConfig &c, cudaStream_t stream) {
    switch (mode) {
        case kModePlain: return launch_tma_one<Row, kModePlain, COPY>(g, p, c, stream);
        case kModeFilter0: return launch_tma_one<Row, kModeFilter0, COPY>(g, p, c, stream);
        case kModeFilter1: return launch_tma_one<Row, kModeFilter1, COPY>(g, p, c, stream);
        case kModeFilter4: return launch_tma_one<Row, kModeFilter4, COPY>(g, p, c, stream);
    }
    return fail(CYP_E_INVALID, "launch_walk_tma: bad mode");
}

int launch_walk_tma(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream) {
    const TmaConfig c = tma_config(g, variant);
    const bool exact = g->cdf_mode == CYP_CDF_EXACT;
    if (c.q16 && c.cpasync)
        return exact ? launch_tma_mode<RowQ16Exact, kCopyCpAsync>(g, p, mode, c, stream)
                     : launch_tma_mode<RowQ16Ref, kCopyCpAsync>(g, p, mode, c, stream);
    if (c.q16)
        return exact ? launch_tma_mode<RowQ16Exact, kCopyBulk>(g, p, mode, c, stream)
                     : launch_tma_mode<RowQ16Ref, kCopyBulk>(g, p, mode, c, stream);
    return exact ? launch_tma_mode<RowF64, kCopyBulk>(g, p, mode, c, stream) : launch_tma_mode<RowF32, kCopyBulk>(g, p, mode, c, stream);
}

const char *walk_tma_name(const cyp_graph *g, int variant) {
    static thread_local char buf[128];
    const TmaConfig c = tma_config(g, variant);
    const bool exact = g->cdf_mode == CYP_CDF_EXACT;
    const char *row = c.q16 ? (exact ? RowQ16Exact::name : RowQ16Ref::name) : (exact ? RowF64::name : RowF32::name);
    snprintf(buf, sizeof(buf), "walk_tma_kernel<%s,%s,threads=%d,slot=%d,%s>", row, c.cpasync ? "cp.async" : "bulk", c.threads,
             c.slot_elems, g->monotone ? "bsearch" : "scan");
    return buf;
}

}  // namespace cyp
