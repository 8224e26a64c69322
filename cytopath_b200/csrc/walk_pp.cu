// K2, paired flavour: the Markov-chain step loop (reference cytopath/markov_chain_sampler.py:45-51)
// with TWO walkers per thread sharing ONE shared-memory row slot.
//
// A step of a walker is two dependent memory phases: (1) its current row of 16-bit CDF codes
// (global -> shared, cooperative 16-byte cp.async), then, after the first-crossing search, (2) one
// 16-byte edge record (next cell + its row descriptor + its label code).  In walk_tma.cu a thread
// sits idle through both.  Here a thread alternates between walkers A and B:
//
//     wait A.row | search A | issue A.edge | use B.edge (issued half a step ago) | issue B.row
//     wait B.row | search B | issue B.edge | use A.edge                          | issue A.row ...
//
// so an edge-record load always overlaps the other walker's row copy, a slot is never needed by both
// walkers at once (A's search is over before B's row lands), and a 1024-thread CTA keeps 2048 walkers
// in flight on 128 KB (+ padding) of slots.  ncu of walk_tma.cu at C4 (profiles/r01_ncu_final_*): 41 %
// issue utilisation, 350 instructions per warp-step of which 158 are the copy loop; this kernel also
// (a) unrolls the copy loop with the slot size as a template parameter and packs (row offset, sectors)
//     into ONE shuffled word, width-limited shuffles (no lane arithmetic): 5 instructions per 16-byte
//     chunk instruction instead of ~20;
// (b) searches with a fixed-depth branchless lower bound (one LDS + compare + predicated add per level);
// (c) pads slots by 32 bytes so that same-position probes of the 32 lanes fall into 4 bank groups.
// Results are bit-identical to every other K2 flavour (same uniform stream, same comparisons).
//
// Applicability: 16-bit codes usable (cyp_graph::q16_ok, monotone rows), every padded row fits one
// slot (<= 256 codes), padded array < 2^32 elements.  Otherwise launch_walk_pp returns kPpNotApplicable
// and the caller uses walk_tma.cu.
#include <cstdio>
#include <type_traits>

#include "walk.cuh"

namespace cyp {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

// A row travels between lanes as ONE word: (32-byte sectors of the padded row) << kOffBits | row offset / 16.
constexpr int kOffBits = 27;

constexpr int kModeExact = 0, kModeRef = 1;

// One walker's registers.
template <int MODE>
struct Walker {
    int32_t cur = 0;                 // current cell; after an edge-record load was issued: the next cell
    uint32_t off_units = 0, len = 0; // row descriptor of `cur`
    uint32_t code = 0;               // label code of `cur`
    int32_t b0 = 0, b1 = 0, b2 = 0;  // trajectory columns 4q .. 4q+2 waiting for the int4 store
    double u_even = 0.0, u_odd = 0.0;
    WalkerDigest<MODE> dg;
};

}  // namespace

// CPR = 16-byte chunks per slot (slot holds 8*CPR codes); CDFM = kModeExact / kModeRef.
template <int CPR, int CDFM, int MODE>
__global__ void __launch_bounds__(1024, 1) walk_pp_kernel(const WalkParams p, const uint16_t *__restrict__ q16) {
    constexpr int SLOT_ELEMS = 8 * CPR;
    constexpr int SLOT_STRIDE = 16 * CPR + 32;           // bytes.  +32: same-position probes of the 32 lanes spread over 4 bank groups.
    // (+16 would give 8 groups, but then the two 16-byte halves of a global sector land in different 32-byte shared-memory
    // sectors for every other slot and LDGSTS fetches the sector twice: ncu showed 6 instead of 4 sectors per 128-byte row.)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int nthreads = blockDim.x;
    unsigned char *const my_slot_b = smem_raw + static_cast<size_t>(threadIdx.x) * SLOT_STRIDE;
    const uint16_t *const slot = reinterpret_cast<const uint16_t *>(my_slot_b);
    // cooperative copy geometry: lane = (group, chunk); copy instruction i moves chunk `chunk` of the row of the
    // group's i-th walker, i.e. of lane (lane & ~(CPR-1)) + i
    const int chunk = lane & (CPR - 1);
    const uint32_t dst0 = smem_u32(my_slot_b) - static_cast<uint32_t>(chunk) * SLOT_STRIDE + static_cast<uint32_t>(chunk) * 16u;
    const unsigned char *const src_c = reinterpret_cast<const unsigned char *>(q16) + chunk * 16;
    const uint32_t need_thr = (static_cast<uint32_t>(chunk >> 1) + 1u) << kOffBits;   // packed >= need_thr <=> sectors > chunk/2

    const EdgeRec *__restrict__ edge = p.edge;
    const int nt = p.nt;
    const bool vec_store = (p.ld & 3) == 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * nthreads * 2;

    for (int64_t wbase = (static_cast<int64_t>(blockIdx.x) * nthreads + (threadIdx.x & ~31)) * 2; wbase < p.n_walkers; wbase += stride) {
        Walker<MODE> A, B;
        const int64_t wA = wbase + lane, wB = wbase + 32 + lane;
        const bool actA = wA < p.n_walkers, actB = wB < p.n_walkers;
        int32_t *__restrict__ outA = p.out_seq + (actA ? wA : 0) * p.ld;
        int32_t *__restrict__ outB = p.out_seq + (actB ? wB : 0) * p.ld;

        auto init = [&](Walker<MODE> &X, bool act, int64_t w, int32_t *out) {
            if (act) {
                const uint64_t gw = static_cast<uint64_t>(p.walker_begin + w);
                X.cur = p.roots[gw / static_cast<uint64_t>(p.sims_per_root)];
                const RowInfo ri = p.rows[X.cur];
                X.off_units = ri.off_units;
                X.len = ri.len;
                if constexpr (WalkerDigest<MODE>::MASKW > 0) X.code = p.code8[X.cur];
                X.b0 = X.cur;
                if (nt == 0 || !vec_store) out[0] = X.cur;
                X.dg.visit(X.cur, X.code);
            }
        };
        // uniform of transition t for walker w (Philox once per two transitions)
        auto draw = [&](Walker<MODE> &X, bool act, int64_t w, int t) {
            if (p.uniforms) {
                if (act) ((t & 1) ? X.u_odd : X.u_even) = p.uniforms[w * static_cast<int64_t>(nt) + t];
            } else if ((t & 1) == 0) {
                walker_uniform_pair(p.seed, p.round, static_cast<uint64_t>(p.walker_begin + w), static_cast<uint32_t>(t >> 1), X.u_even, X.u_odd);
            }
        };
        // the warp's 32 rows of walker set X -> the 32 slots
        auto issue_copy = [&](const Walker<MODE> &X, bool act) {
            const uint32_t len_pad = (X.len + 15u) & ~15u;
            const uint32_t packed = act ? (((len_pad >> 4) << kOffBits) | X.off_units) : 0u;   // (32-byte sectors, row offset)
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                const uint32_t pk = __shfl_sync(0xffffffffu, packed, i, CPR);
                const unsigned char *src = src_c + static_cast<uint64_t>(pk & ((1u << kOffBits) - 1u)) * (kRowAlign * 2);
                if (pk >= need_thr) cp_async16(dst0 + static_cast<uint32_t>(i) * SLOT_STRIDE, src);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // first crossing in the staged row of X, then the edge-record load (its result lands in X's registers and
        // is first touched half a step later)
        auto search_and_follow = [&](Walker<MODE> &X, bool act, int t) {
            if (!act) return;
            const double u = (t & 1) ? X.u_odd : X.u_even;
            const int len = static_cast<int>(X.len);
            int key, t_def = 0;
            if constexpr (CDFM == kModeExact) {
                t_def = static_cast<int>(ceil(u * 65536.0));           // q >= t_def      <=>  u <= q / 65536
                key = t_def - 1;                                         // q >= t_def - 1  <=>  u <= (q + 1) / 65536
            } else {
                key = static_cast<int>(ceil(u * 1000.0));               // smallest integer m with u <= (double)fl32(m / 1000)
                while (key > 0 && u <= static_cast<double>(__fdiv_rn(static_cast<float>(key - 1), 1000.0f))) --key;
                while (u > static_cast<double>(__fdiv_rn(static_cast<float>(key), 1000.0f))) ++key;
            }
            int j = 0;
#pragma unroll
            for (int step = SLOT_ELEMS / 2; step >= 1; step >>= 1) {
                const int probe = j + step - 1;
                if (probe < len && static_cast<int>(slot[probe]) < key) j += step;
            }
            const int64_t off = static_cast<int64_t>(X.off_units) * kRowAlign;
            if constexpr (CDFM == kModeExact) {
                const double *exact = static_cast<const double *>(p.cdf) + off;
                while (j < len) {
                    const int q = static_cast<int>(slot[j]);
                    if (q >= t_def) break;                              // certain crossing
                    if (q >= key && u <= __ldg(exact + j)) break;       // undecided by the code: float64 compare (:49)
                    ++j;
                }
            } else {
                while (j < len && static_cast<int>(slot[j]) < key) ++j;
            }
            if (j < len) {
                const EdgeRec e = load_edge(edge + off + j);
                X.cur = e.col;
                X.off_units = e.off_units;
                X.len = e.len;
                X.code = e.code;
            } else if (p.nocross) {
                atomicAdd(p.nocross, 1ull);                             // stays where it is (reference: IndexError at :49)
            }
        };
        // X.cur is the cell of trajectory column c
        auto record = [&](Walker<MODE> &X, bool act, int c, int32_t *out) {
            if (!act) return;
            if (vec_store) {
                const int q = c & 3;
                if (q == 3) {
                    *reinterpret_cast<int4 *>(out + (c & ~3)) = make_int4(X.b0, X.b1, X.b2, X.cur);
                } else {
                    if (q == 0) X.b0 = X.cur; else if (q == 1) X.b1 = X.cur; else X.b2 = X.cur;
                    if (c == nt) {
                        out[c & ~3] = X.b0;
                        if (q >= 1) out[(c & ~3) + 1] = X.b1;
                        if (q >= 2) out[(c & ~3) + 2] = X.b2;
                    }
                }
            } else {
                out[c] = X.cur;
            }
            X.dg.visit(X.cur, X.code);
        };

        init(A, actA, wA, outA);
        init(B, actB, wB, outB);
        if (nt > 0) {
            draw(A, actA, wA, 0);
            draw(B, actB, wB, 0);
            issue_copy(A, actA);
        }
        for (int t = 0; t < nt; ++t) {
            // ---- A steps (transition t); B's row goes out
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            search_and_follow(A, actA, t);
            __syncwarp();                                   // every lane is done with its slot
            if (t > 0) record(B, actB, t, outB);            // B's edge record of transition t-1 is needed from here on
            issue_copy(B, actB);
            if (t + 1 < nt) draw(A, actA, wA, t + 1);
            // ---- B steps (transition t); A's row of transition t+1 goes out
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            search_and_follow(B, actB, t);
            __syncwarp();
            record(A, actA, t + 1, outA);
            if (t + 1 < nt) {
                issue_copy(A, actA);
                draw(B, actB, wB, t + 1);
            }
        }
        if (nt > 0) record(B, actB, nt, outB);
        if (actA) A.dg.finish(p, wA, A.cur);
        if (actB) B.dg.finish(p, wB, B.cur);
    }
}

struct PpConfig {
    int cpr;        // 16-byte chunks per slot
    int threads;
    size_t smem;
};

// variant: 0 = auto; 9400 + T/32 = T threads per CTA.
static bool pp_config(const cyp_graph *g, int variant, int64_t n_walkers, int sm_count, PpConfig *out) {
    if (!(g->q16_ok && g->monotone)) return false;
    if (g->padded >= (static_cast<int64_t>(kRowAlign) << kOffBits)) return false;
    const int max_pad = (g->max_len + 15) / 16 * 16;
    int cpr = 2;
    while (cpr * 8 < max_pad) cpr <<= 1;
    if (cpr > 32) return false;
    const size_t stride = 16 * cpr + 32;
    int threads = static_cast<int>((226 * 1024) / stride) / 32 * 32;
    if (threads > 1024) threads = 1024;
    if (variant > 9400) {
        threads = (variant - 9400) * 32 < threads ? (variant - 9400) * 32 : threads;
    } else if (n_walkers > 0 && sm_count > 0) {
        // wave balancing: every warp runs the same number of 64-walker batches as with the largest CTA, but
        // with no more warps than that needs (fewer walkers in flight = lower memory latency, same batch count)
        const int64_t batches = (n_walkers + 63) / 64;
        const int64_t per_round = static_cast<int64_t>(sm_count) * (threads / 32);
        const int64_t rounds = (batches + per_round - 1) / per_round;
        const int64_t warps = (batches + static_cast<int64_t>(sm_count) * rounds - 1) / (static_cast<int64_t>(sm_count) * rounds);
        if (warps * 32 < threads) threads = static_cast<int>(warps) * 32;
    }
    if (threads < 32) threads = 32;
    *out = PpConfig{cpr, threads, static_cast<size_t>(threads) * stride};
    return true;
}

template <int CPR, int CDFM, int MODE>
static int launch_pp_one(const cyp_graph *g, const WalkParams &p, const PpConfig &c, int sm_count, cudaStream_t stream) {
    auto kernel = walk_pp_kernel<CPR, CDFM, MODE>;
    static size_t smem_set = 0;
    if (c.smem > smem_set) {
        CYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c.smem)));
        smem_set = c.smem;
    }
    const int64_t need = (p.n_walkers + 2 * c.threads - 1) / (2 * c.threads);
    const unsigned grid = static_cast<unsigned>(need < sm_count ? need : sm_count);     // persistent: one CTA per SM
    if (grid == 0) return CYP_OK;
    kernel<<<grid, c.threads, c.smem, stream>>>(p, g->d_q16);
    CYP_CUDA(cudaGetLastError());
    return CYP_OK;
}

template <int CPR, int CDFM>
static int launch_pp_mode(const cyp_graph *g, const WalkParams &p, WalkMode mode, const PpConfig &c, int sm_count, cudaStream_t stream) {
    switch (mode) {
        case kModePlain: return launch_pp_one<CPR, CDFM, kModePlain>(g, p, c, sm_count, stream);
        case kModeFilter0: return launch_pp_one<CPR, CDFM, kModeFilter0>(g, p, c, sm_count, stream);
        case kModeFilter1: return launch_pp_one<CPR, CDFM, kModeFilter1>(g, p, c, sm_count, stream);
        case kModeFilter4: return launch_pp_one<CPR, CDFM, kModeFilter4>(g, p, c, sm_count, stream);
    }
    return fail(CYP_E_INVALID, "launch_walk_pp: bad mode");
}

template <int CPR>
static int launch_pp_cdf(const cyp_graph *g, const WalkParams &p, WalkMode mode, const PpConfig &c, int sm_count, cudaStream_t stream) {
    return g->cdf_mode == CYP_CDF_EXACT ? launch_pp_mode<CPR, kModeExact>(g, p, mode, c, sm_count, stream)
                                        : launch_pp_mode<CPR, kModeRef>(g, p, mode, c, sm_count, stream);
}

static int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_dev = dev;
    }
    return cached;
}

int launch_walk_pp(const cyp_graph *g, const WalkParams &p, WalkMode mode, int variant, cudaStream_t stream) {
    const int sm_count = device_sm_count();
    PpConfig c;
    if (sm_count <= 0 || !pp_config(g, variant, p.n_walkers, sm_count, &c)) return kPpNotApplicable;
    switch (c.cpr) {
        case 2: return launch_pp_cdf<2>(g, p, mode, c, sm_count, stream);
        case 4: return launch_pp_cdf<4>(g, p, mode, c, sm_count, stream);
        case 8: return launch_pp_cdf<8>(g, p, mode, c, sm_count, stream);
        case 16: return launch_pp_cdf<16>(g, p, mode, c, sm_count, stream);
        case 32: return launch_pp_cdf<32>(g, p, mode, c, sm_count, stream);
    }
    return kPpNotApplicable;
}

const char *walk_pp_name(const cyp_graph *g, int variant) {
    static thread_local char buf[128];
    PpConfig c;
    if (!pp_config(g, variant, 0, 0, &c)) return nullptr;
    snprintf(buf, sizeof(buf), "walk_pp_kernel<%s,cp.async,2 walkers/thread,threads<=%d,slot=%d>",
             g->cdf_mode == CYP_CDF_EXACT ? "q16/f64" : "q16/f32", c.threads, 8 * c.cpr);
    return buf;
}

}  // namespace cyp
