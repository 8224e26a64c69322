// K2, fused-row flavour (default where the layout applies): the Markov-chain step loop (reference
// cytopath/markov_chain_sampler.py:45-51) over the "v2" rows of common.cuh / rows_v2.cu.
//
// One walker per thread, one persistent CTA per SM, one shared-memory slot per thread.  Per step:
//   1. the warp copies its 32 row blocks into the slots with cooperative 16-byte cp.async (LDGSTS): a walker
//      is represented by its ROW WORD (sectors << 27 | first sector), which is all a copy needs, so one
//      width-limited shuffle per copy instruction moves it between lanes;
//   2. the header tells the cell id (-> trajectory column, hash, label bitmap), the row length and which
//      entries have their target inlined;
//   3. first crossing `min j : u <= cdf[j]` (ties to the bin, :49): fixed-depth branchless lower bound over
//      the upper-8-bit plane, then the 12-bit codes decide; a compare the code cannot decide (code ==
//      floor(u*4096)-ish, ~k/4096 of the steps) is settled against the float64 CDF, so the selected entry
//      is ALWAYS the one the reference's comparison selects;
//   4. the next row word: from the slot if the entry is inlined (the row's most probable entries: most
//      steps), else one 4-byte gather from edge4.
// The uniform stream, the comparisons and therefore the trajectories are bit-identical to walk.cu /
// walk_tma.cu / walk_pp.cu and to the oracle.
#include <cstdio>
#include <cstdlib>

#include "walk.cuh"

namespace cyp {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async16_l1(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ uint32_t ldg_word(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

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
        "V2_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra V2_WAIT_DONE;\n"
        "bra V2_WAIT_LOOP;\n"
        "V2_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map); completion is reported to `bar` in bytes
__device__ __forceinline__ void bulk_copy_g2s(const void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kModeExact = 0, kModeRef = 1;

}  // namespace

// CPR = 16-byte chunks per slot; CDFM = kModeExact / kModeRef; L1ROWS: the row copies allocate in L1 (cp.async.ca
// instead of .cg).  On small graphs many walkers sit on the same few cells, every SM asks the same L2 lines at once
// and the L2 slices serialise them: 13.7 G walker-steps/s at 3,696 cells against 85 G at 20k+ cells.  With L1
// allocation the hot rows are served by the SM's own L1: 80 G at 3,696 .. 10,000 cells (profiles/r01_small_graphs.log);
// above ~16k cells the L1-bypassing copies are the faster ones.
// CPR = 0: long rows.  Every thread issues ONE cp.async.bulk (UBLKCP) for its whole block and the warp's 32 copies
// complete on one mbarrier; the slot is exactly the largest block (`bulk_stride` bytes, any multiple of 32).
// OVERSIZE: the graph has rows whose block is longer than the slot; they are searched in global memory (a template
// parameter because the test costs 2-3 % on graphs that have none).
template <int CPR, int CDFM, int MODE, bool L1ROWS, bool OVERSIZE>
__global__ void __launch_bounds__(1024, 1) walk_v2_kernel(const WalkParams p, const int log_n /* search depth: 2^log_n >= longest row */,
                                                          const int bulk_stride) {
    constexpr bool BULK = CPR == 0;
    const int SLOT_STRIDE = BULK ? bulk_stride : 16 * CPR + 32;   // +32 spreads same-position probes over 4 bank groups and keeps slots sector aligned
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int nthreads = blockDim.x;
    const int bar_bytes = BULK ? (((nthreads >> 5) * 8 + 127) & ~127) : 0;
    const unsigned char *const slot = smem_raw + bar_bytes + static_cast<size_t>(threadIdx.x) * SLOT_STRIDE;
    uint64_t *const bar = reinterpret_cast<uint64_t *>(smem_raw) + (threadIdx.x >> 5);
    uint32_t parity = 0;
    if constexpr (BULK) {
        if (lane == 0) mbar_init(bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    // cooperative copy geometry: lane = (group, chunk); instruction i moves chunk `chunk` of the block of lane (lane & ~(CPR-1)) + i
    const int chunk = BULK ? 0 : lane & ((BULK ? 1 : CPR) - 1);
    const uint32_t dst0 = smem_u32(slot) - static_cast<uint32_t>(chunk) * SLOT_STRIDE + static_cast<uint32_t>(chunk) * 16u;
    const uint32_t src_c = static_cast<uint32_t>(chunk) * 16u;     // byte offsets into rows2 fit 32 bits (< 2^27 sectors)
    const uint32_t need_thr = (static_cast<uint32_t>(chunk >> 1) + 1u) << kWordOffBits;     // word >= need_thr <=> sectors > chunk / 2

    const int nt = p.nt;
    const bool vec_store = (p.ld & 3) == 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * nthreads;
    auto issue_copy = [&](uint32_t word) {
        if constexpr (BULK) {
            const uint32_t bytes = (word >> kWordOffBits) * 32u;
            if (bytes) {
                mbar_arrive_expect_tx(bar, bytes);
                bulk_copy_g2s(slot, p.rows2 + (static_cast<size_t>(word & kWordOffMask) << 5), bytes, bar);
            } else {
                mbar_arrive(bar);
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < (BULK ? 1 : CPR); ++i) {
            const uint32_t wd = __shfl_sync(0xffffffffu, word, i, BULK ? 32 : CPR);
            const uint32_t off = (wd << 5) + src_c;                  // the shift drops the sector count
            if (wd >= need_thr) {
                if constexpr (L1ROWS) cp_async16_l1(dst0 + static_cast<uint32_t>(i) * SLOT_STRIDE, p.rows2 + off);
                else cp_async16(dst0 + static_cast<uint32_t>(i) * SLOT_STRIDE, p.rows2 + off);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int64_t wbase = static_cast<int64_t>(blockIdx.x) * nthreads + (threadIdx.x & ~31); wbase < p.n_walkers; wbase += stride) {
        const int64_t w = wbase + lane;
        const bool active = w < p.n_walkers;
        const uint64_t gw = static_cast<uint64_t>(p.walker_begin + w);
        uint32_t word = 0u;                               // inactive lanes copy nothing
        if (active) word = p.word[p.roots[gw / static_cast<uint64_t>(p.sims_per_root)]];
        int32_t *__restrict__ out_row = p.out_seq + (active ? w : 0) * p.ld;
        const double *__restrict__ urow = (p.uniforms && active) ? p.uniforms + w * static_cast<int64_t>(nt) : nullptr;
        int32_t b0 = 0, b1 = 0, b2 = 0;
        int32_t cell = 0;
        WalkerDigest<MODE> dg;
        double u_even = 0.0, u_odd = 0.0;
        issue_copy(word);
        if (!p.uniforms && nt > 0) walker_uniform_pair(p.seed, p.round, gw, 0u, u_even, u_odd);

        for (int t = 0; t <= nt; ++t) {
            if constexpr (BULK) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            uint32_t next = word;
            if (active) {
                const uint4 h = *reinterpret_cast<const uint4 *>(slot);          // cell | len, label, m | inline mask
                cell = static_cast<int32_t>(h.x);
                const int len = static_cast<int>(h.y & 0xffffu);
                // ---- trajectory column t, hash, label bitmap
                if (vec_store) {
                    const int q = t & 3;
                    if (q == 3) {
                        *reinterpret_cast<int4 *>(out_row + (t & ~3)) = make_int4(b0, b1, b2, cell);
                    } else {
                        if (q == 0) b0 = cell; else if (q == 1) b1 = cell; else b2 = cell;
                        if (t == nt) {
                            out_row[t & ~3] = b0;
                            if (q >= 1) out_row[(t & ~3) + 1] = b1;
                            if (q >= 2) out_row[(t & ~3) + 2] = b2;
                        }
                    }
                } else {
                    out_row[t] = cell;
                }
                dg.visit(cell, (h.y >> 16) & 0xffu);
                if (t < nt) {
                    // ---- first crossing
                    const double u = urow ? urow[t] : ((t & 1) ? u_odd : u_even);
                    int key, t_def = 0;
                    if constexpr (CDFM == kModeExact) {
                        t_def = static_cast<int>(ceil(u * 4096.0));            // q >= t_def      <=>  u <= q / 4096
                        key = t_def - 1;                                         // q >= t_def - 1  <=>  u <= (q + 1) / 4096
                    } else {
                        key = static_cast<int>(ceil(u * 1000.0));               // smallest integer m with u <= (double)fl32(m / 1000)
                        while (key > 0 && u <= static_cast<double>(__fdiv_rn(static_cast<float>(key - 1), 1000.0f))) --key;
                        while (u > static_cast<double>(__fdiv_rn(static_cast<float>(key), 1000.0f))) ++key;
                        key <<= 2;                                               // the codes are 4 m (rows_v2.cu): 8 significant bits in the upper plane
                    }
                    // a block longer than the slot (a few long rows of a ragged graph: the slot is sized for the typical
                    // row) is searched where it lies, in global memory; its header did fit
                    const bool oversize = OVERSIZE && !BULK && (word >> kWordOffBits) * 32u > static_cast<uint32_t>(16 * CPR);
                    auto follow = [&](const unsigned char *blk) -> int {        // returns the entry, sets `next`
                        const unsigned char *hi = blk + kV2Header;
                        const unsigned char *lo = hi + len;
                        const int key_hi = key >> 4;
                        int j = 0;
                        for (int step = 1 << (log_n - 1); step >= 1; step >>= 1) {
                            const int probe = j + step - 1;
                            if (probe < len && static_cast<int>(hi[probe]) < key_hi) j += step;
                        }
                        auto code_at = [&](int jj) { return (static_cast<int>(hi[jj]) << 4) | ((static_cast<int>(lo[jj >> 1]) >> ((jj & 1) * 4)) & 15); };
                        if constexpr (CDFM == kModeExact) {
                            while (j < len) {
                                const int q = code_at(j);
                                if (q >= t_def) break;                           // certain crossing
                                if (q == key) {                                  // undecided by the code: float64 compare (:49)
                                    const RowInfo ri = p.rows[cell];
                                    if (u <= __ldg(static_cast<const double *>(p.cdf) + static_cast<int64_t>(ri.off_units) * kRowAlign + j)) break;
                                }
                                ++j;
                            }
                        } else {
                            while (j < len && code_at(j) < key) ++j;
                        }
                        // ---- next row word
                        if (j < len) {
                            const unsigned long long mask = (static_cast<unsigned long long>(h.w) << 32) | h.z;
                            if (j < 64 && ((mask >> j) & 1ull)) {
                                const int rank = __popcll(mask & ((1ull << j) - 1ull));
                                next = *reinterpret_cast<const uint32_t *>(blk + v2_used_bytes(len) + 4 * rank);
                            } else {
                                const uint32_t ebase = *reinterpret_cast<const uint32_t *>(blk + 16);
                                next = ldg_word(p.edge4 + static_cast<size_t>(ebase) * 4 + j);
                            }
                        }
                        return j;
                    };
                    const int j = oversize ? follow(p.rows2 + (static_cast<size_t>(word & kWordOffMask) << 5)) : follow(slot);
                    if (j >= len && p.nocross) atomicAdd(p.nocross, 1ull);       // stays where it is (reference: IndexError at :49)
                }
            }
            if (t == nt) break;
            word = next;
            __syncwarp();                                                         // every lane is done with its slot
            issue_copy(word);
            if (!p.uniforms && (t & 1)) walker_uniform_pair(p.seed, p.round, gw, static_cast<uint32_t>((t + 1) >> 1), u_even, u_odd);
        }
        if (active) dg.finish(p, w, cell);
    }
}

// Tried on C4 and dropped (profiles/r01_v2_experiments.md, r01_v2_flags.log): L2 cache hints (streaming
// trajectory stores, evict-first target gathers, evict-last row copies: 52.2-52.3 G walker-steps/s with any
// combination, i.e. no effect), a persisting-L2 access-policy window on the fused rows (hitRatio 0.3..1.0: 24-34 G,
// much worse), two CTAs of 512..640 threads per SM (57.7-58.6 vs 57.0 G: no gain from more walkers in flight), a
// fully unrolled 7-level search with a separate refine loop (3-6 % slower: two more dependent shared-memory round
// trips per step), one or two extra sectors of inlined targets per row (51 / 48 vs 57 G at C4; +10 % at C3, where it
// is now the default), a lower bound on the full 12-bit code unrolled per row length (no gain), row copies through
// registers (LDG.128 + STS.128: 35 vs 58 G).  ncu (profiles/r01_ncu_walk_v2_fused_c4_exact.csv): L2 hit rate 70 %, issue
// slots 61 % busy, DRAM 2 TB/s.  A dependency-free kernel gathers 96 G rows/s of 128 bytes out of a 128 MB array on the
// same GPU (profiles/r01_gather_ceilings_tuned.log); this kernel, where every gather waits for the search on the row
// before, moves 58 G/s plus the target words and the trajectory stores.  What moves it: bytes per step, the L2 hit rate
// (locality of the graph), the share of inlined targets (skew of the probabilities).
constexpr int64_t kV2L1Cells = 16384;      // below: L1-allocating row copies (see walk_v2_kernel)

struct V2Config {
    int cpr;            // 16-byte chunks per slot (cooperative cp.async copies); 0 = one bulk copy per row
    int ctas_per_sm;
    int threads;
    int log_n;
    int stride;         // bytes between slots
    bool oversize;      // some blocks are longer than the slot
    size_t smem;
};

// variant: 0 = auto; 9500 + T/32 = T threads per CTA; 9600 + T/32 = two CTAs of T threads per SM (experiment);
// 9700 + T/32 = slot sized for the longest row (bulk copies above 256 bytes).
static bool v2_config(const cyp_graph *g, int variant, int64_t n_walkers, int sm_count, V2Config *out) {
    if (!g->v2_ok) return false;
    // slot = the block size that covers (almost) every row (cyp_graph::v2_slot_sectors, rows_v2.cu); variants 97xx
    // force the largest block instead (then rows above 256 bytes take the bulk flavour)
    bool force_max = false;
    if (variant > 9700) { force_max = true; variant -= 200; }
    if (force_max && g->v2_max_sectors > kV2MaxSectors) return false;      // a bulk copy needs the block's exact size in the row word
    const int block = (force_max ? g->v2_max_sectors : g->v2_slot_sectors) * 32;
    int cpr = 0, stride = 0;
    if (block <= 256) {                                  // small rows: cooperative 16-byte copies, power-of-two slots
        cpr = 2;
        while (cpr * 16 < block) cpr <<= 1;
        stride = 16 * cpr + 32;
    } else {                                             // long rows: one bulk copy per row, slot = the largest block
        stride = block + ((block % 128 == 0) ? 32 : 0);  // never a multiple of 128: same-position probes spread over the banks
    }
    int ctas = 1;
    if (variant > 9600) { ctas = 2; variant -= 100; }
    const size_t budget = 226 * 1024 / ctas - 1024;
    int threads = static_cast<int>(budget / stride) / 32 * 32;
    if (cpr == 0) threads = static_cast<int>((budget - (((threads >> 5) * 8 + 127) & ~127)) / stride) / 32 * 32;
    if (threads > 1024) threads = 1024;
    if (variant > 9500) {
        threads = (variant - 9500) * 32 < threads ? (variant - 9500) * 32 : threads;
    } else if (n_walkers > 0 && sm_count > 0) {
        // wave balancing: every warp runs as many 32-walker batches as with the largest CTA, with no more warps than that needs
        const int64_t batches = (n_walkers + 31) / 32;
        const int64_t per_round = static_cast<int64_t>(sm_count) * (threads / 32);
        const int64_t rounds = (batches + per_round - 1) / per_round;
        const int64_t warps = (batches + static_cast<int64_t>(sm_count) * rounds - 1) / (static_cast<int64_t>(sm_count) * rounds);
        if (warps * 32 < threads) threads = static_cast<int>(warps) * 32;
    }
    if (threads < 32) threads = 32;
    int log_n = 1;
    while ((1 << log_n) < g->max_len) ++log_n;
    const size_t bars = cpr == 0 ? static_cast<size_t>(((threads >> 5) * 8 + 127) & ~127) : 0;
    *out = V2Config{cpr, ctas, threads, log_n, stride, cpr != 0 && g->v2_max_sectors * 32 > 16 * cpr, bars + static_cast<size_t>(threads) * stride};
    return true;
}

static int v2_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_dev = dev;
    }
    return cached;
}

template <int CPR, int CDFM, int MODE, bool L1ROWS, bool OVERSIZE>
static int launch_v2_l1(const WalkParams &p, const V2Config &c, int sm_count, cudaStream_t stream) {
    auto kernel = walk_v2_kernel<CPR, CDFM, MODE, L1ROWS, OVERSIZE>;
    static size_t smem_set = 0;
    if (c.smem > smem_set) {
        CYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c.smem)));
        smem_set = c.smem;
    }
    const int64_t need = (p.n_walkers + c.threads - 1) / c.threads;
    const int64_t resident = static_cast<int64_t>(sm_count) * c.ctas_per_sm;
    const unsigned grid = static_cast<unsigned>(need < resident ? need : resident);       // persistent CTAs
    if (grid == 0) return CYP_OK;
    static const bool trace = getenv("CYP_TRACE") != nullptr;
    if (trace) {
        int dev = -1;
        cudaGetDevice(&dev);
        fprintf(stderr, "[cyp trace] walk_v2: device %d grid %u threads %d smem %zu walkers %lld begin %lld sims/root %lld nt %d rows2 %p edge4 %p out %p\n", dev, grid,
                c.threads, c.smem, static_cast<long long>(p.n_walkers), static_cast<long long>(p.walker_begin),
                static_cast<long long>(p.sims_per_root), p.nt, static_cast<const void *>(p.rows2), static_cast<const void *>(p.edge4), static_cast<const void *>(p.out_seq));
    }
    kernel<<<grid, c.threads, c.smem, stream>>>(p, c.log_n, c.stride);
    CYP_CUDA(cudaGetLastError());
    return CYP_OK;
}

template <int CPR, int CDFM, int MODE>
static int launch_v2_one(const WalkParams &p, const V2Config &c, int sm_count, cudaStream_t stream) {
    if constexpr (CPR == 0) return launch_v2_l1<CPR, CDFM, MODE, false, false>(p, c, sm_count, stream);     // bulk copies: whole rows, L2 only
    if (c.oversize)
        return p.l1_rows ? launch_v2_l1<CPR, CDFM, MODE, true, true>(p, c, sm_count, stream) : launch_v2_l1<CPR, CDFM, MODE, false, true>(p, c, sm_count, stream);
    return p.l1_rows ? launch_v2_l1<CPR, CDFM, MODE, true, false>(p, c, sm_count, stream) : launch_v2_l1<CPR, CDFM, MODE, false, false>(p, c, sm_count, stream);
}

template <int CPR, int CDFM>
static int launch_v2_mode(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    switch (mode) {
        case kModePlain: return launch_v2_one<CPR, CDFM, kModePlain>(p, c, sm_count, stream);
        case kModeFilter0: return launch_v2_one<CPR, CDFM, kModeFilter0>(p, c, sm_count, stream);
        case kModeFilter1: return launch_v2_one<CPR, CDFM, kModeFilter1>(p, c, sm_count, stream);
        case kModeFilter4: return launch_v2_one<CPR, CDFM, kModeFilter4>(p, c, sm_count, stream);
    }
    return fail(CYP_E_INVALID, "launch_walk_v2: bad mode");
}

template <int CPR>
static int launch_v2_cdf(const cyp_graph *g, const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    return g->cdf_mode == CYP_CDF_EXACT ? launch_v2_mode<CPR, kModeExact>(p, mode, c, sm_count, stream)
                                        : launch_v2_mode<CPR, kModeRef>(p, mode, c, sm_count, stream);
}

int launch_walk_v2(const cyp_graph *g, const WalkParams &p_in, WalkMode mode, int variant, cudaStream_t stream) {
    const int sm_count = v2_sm_count();
    V2Config c;
    if (sm_count <= 0 || !v2_config(g, variant, p_in.n_walkers, sm_count, &c)) return kNotApplicable;
    WalkParams p = p_in;
    p.rows2 = g->d_rows2;
    p.edge4 = g->d_edge4;
    p.word = g->d_word;
    p.l1_rows = g->n < kV2L1Cells ? 1 : 0;
    if (const char *e = getenv("CYP_V2_L1")) p.l1_rows = atoi(e);
    switch (c.cpr) {
        case 0: return launch_v2_cdf<0>(g, p, mode, c, sm_count, stream);
        case 2: return launch_v2_cdf<2>(g, p, mode, c, sm_count, stream);
        case 4: return launch_v2_cdf<4>(g, p, mode, c, sm_count, stream);
        case 8: return launch_v2_cdf<8>(g, p, mode, c, sm_count, stream);
        case 16: return launch_v2_cdf<16>(g, p, mode, c, sm_count, stream);
    }
    return kNotApplicable;
}

const char *walk_v2_name(const cyp_graph *g, int variant) {
    static thread_local char buf[128];
    V2Config c;
    if (!v2_config(g, variant, 0, 0, &c)) return nullptr;
    snprintf(buf, sizeof(buf), "walk_v2_kernel<%s,fused rows,threads<=%d,slot=%dB%s>", g->cdf_mode == CYP_CDF_EXACT ? "q12/f64" : "q12/f32",
             c.threads, c.cpr ? 16 * c.cpr : c.stride, c.cpr == 0 ? ",bulk" : (g->n < kV2L1Cells ? ",L1 rows" : ""));
    return buf;
}

}  // namespace cyp
