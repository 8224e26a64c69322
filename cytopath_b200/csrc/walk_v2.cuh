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
//
// Shared-memory layout (round 2; profiles/r02_k2_*.md).  Slots are exactly 16*CPR bytes (32 .. 256) with no
// padding, and every slot is ROTATED: the 16-byte chunk c of a row sits at chunk (c + rot) % CPR of its slot,
// rot = lane % CPR (for slots below 128 bytes: the lanes that share a 128-byte bank row rotate together).
//   * every cp.async of a quarter warp writes one aligned 128-byte line: 4 shared-memory wavefronts per copy
//     instruction instead of 7 (the padded 160-byte slots of round 1 split six of every eight 128-byte writes
//     over two lines: 56 wavefronts per warp-step, ncu r01);
//   * lanes reading the same row position (the header, the first search probes) hit 8 different bank groups:
//     4-way conflicts instead of 8-way.
// Instruction diet (377 -> ~200 per warp-step): the step is unrolled four times so that the trajectory
// buffer, the Philox refresh and the parity of the uniform are static; with the native stream the 12-bit key
// comes from the Philox words by integer arithmetic (no I2F.F64 / DMUL / F2I.CEIL on the path: the double is
// only built for an undecided compare); the search is unrolled; a graph whose rows (nearly) all fill their
// slot copies without per-chunk predicates.
// The uniform stream, the comparisons and therefore the trajectories are bit-identical to walk.cu /
// walk_tma.cu and to the oracle.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "walk.cuh"

namespace cyp {

namespace v2 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <bool L1>
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gmem_src) {
    if constexpr (L1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ uint32_t ldg_word(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return static_cast<int>(v);
}

constexpr int kModeExact = 0, kModeRef = 1;
// what the graph's rows look like against the slot: every block (nearly) fills it / shorter blocks exist (copies
// are predicated per chunk) / blocks longer than the slot exist as well (searched where they lie, in HBM)
constexpr int kUniform = 0, kRagged = 1, kOversize = 2;

template <int N>
using IC = std::integral_constant<int, N>;

}  // namespace v2

// CPR = 16-byte chunks per slot; CDFM = kModeExact / kModeRef; FLAVOR: see above; L1ROWS: the row copies allocate
// in L1 (cp.async.ca instead of .cg).  On small graphs many walkers sit on the same few cells, every SM asks the
// same L2 lines at once and the L2 slices serialise them: 13.7 G walker-steps/s at 3,696 cells against 85 G at
// 20k+ cells.  With L1 allocation the hot rows are served by the SM's own L1: 80 G at 3,696 .. 10,000 cells
// (profiles/r01_small_graphs.log); above ~16k cells the L1-bypassing copies are the faster ones.
// CTAS = CTAs per SM: 1 x 1024 threads (up to 64 registers) or 2 x 640 threads (48 registers, a few spills in the
// filter modes): 25 % more walkers in flight.  Measured (profiles/r02_k2_experiments.md): +7.6 % at C4 exact (rows out
// of HBM: latency to hide), -4.6 % at C3 (rows in L2: the register diet costs more than the extra warps give).
// Tried in round 2 and dropped: two walkers per thread with a slot each, the warp alternating between its two sets
// while the other set's copies are in flight (53.7 G with 2 x 896 walkers per SM against 60.3 G: a warp's serial
// dependent chain, not the exposed memory latency, is what limits a step -- ncu: 2,330 of 4,240 cycles per warp-step).
template <int MODE>
struct V2Walker {
    uint32_t word = 0u;          // row word of the row being copied / searched
    bool active = false;
    int32_t *out_row = nullptr;
    const double *urow = nullptr;
    int32_t b0 = 0, b1 = 0, b2 = 0, cell = 0;
    uint32_t gw_lo = 0u, gw_hi = 0u, round = 0u;
    int64_t w = 0;
    uint4 x = make_uint4(0u, 0u, 0u, 0u);     // Philox output for transitions 2k (x, y) and 2k + 1 (z, w)
    WalkerDigest<MODE> dg;
};

template <int CPR, int CDFM, int MODE, int FLAVOR, bool L1ROWS, int CTAS>
__global__ void __launch_bounds__(CTAS == 2 ? 640 : 1024, CTAS) walk_v2_kernel(const WalkParams p, const int log_n /* 2^log_n >= longest row */) {
    using namespace v2;
    constexpr int SLOT = 16 * CPR;
    constexpr uint32_t SMASK = SLOT - 1;
    constexpr int LPB = CPR >= 8 ? 1 : 8 / CPR;            // lanes whose slots share one 128-byte bank row
    constexpr int DEPTH = CPR == 2 ? 3 : CPR == 4 ? 5 : CPR == 8 ? 7 : 8;     // 2^DEPTH >= longest row a slot can hold
    extern __shared__ __align__(256) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int nthreads = blockDim.x;
    const uint32_t slot_sm = smem_u32(smem_raw) + static_cast<uint32_t>(threadIdx.x) * SLOT;     // SLOT-aligned: OR-able
    const uint32_t rot16 = static_cast<uint32_t>((lane / LPB) & (CPR - 1)) * 16u;
    const uint32_t r20 = kV2Header + rot16;

    // cooperative copy geometry: lane = (group, c); instruction i serves the row of lane gb + i: this lane fetches
    // logical chunk c of that row (source = row + 16 c) and writes it to the chunk (c + rot_i) % CPR of that row's
    // slot (CPR precomputed shared-memory addresses)
    const int c = lane & (CPR - 1);
    const int gb = lane & ~(CPR - 1);
    const unsigned char *const src_c = p.rows2 + c * 16;
    uint32_t dst[CPR];
#pragma unroll
    for (int i = 0; i < CPR; ++i)
        dst[i] = slot_sm + static_cast<uint32_t>(i - c) * SLOT + (static_cast<uint32_t>(c + (((gb + i) / LPB) & (CPR - 1))) * 16u & SMASK);

    const int nt = p.nt;
    const bool vec_store = (p.ld & 3) == 0;
    const bool native = p.uniforms == nullptr;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * nthreads;
    const int depth = log_n < DEPTH ? log_n : DEPTH;
    const uint2 pkey = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));

    auto issue_copy = [&](uint32_t word) {
        if constexpr (FLAVOR == kUniform) {
            const uint32_t woff = word << 5;                           // byte offset of the block (the shift drops the sector count)
#pragma unroll
            for (int i = 0; i < CPR; ++i) cp_async16<L1ROWS>(dst[i], src_c + __shfl_sync(0xffffffffu, woff, i, CPR));
        } else {
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                const uint32_t wd = __shfl_sync(0xffffffffu, word, i, CPR);
                if ((wd >> kWordOffBits) > static_cast<uint32_t>(c >> 1))          // the block has this sector
                    cp_async16<L1ROWS>(dst[i], src_c + (wd << 5));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // One transition of the walker whose row is in the slot: writes trajectory column t, returns the next row word.
    auto process = [&](V2Walker<MODE> &s, auto Q, const int t) -> uint32_t {
        constexpr int q = decltype(Q)::value;
        const uint32_t slot_k = slot_sm;
        auto at = [&](uint32_t x) { return slot_k | ((x + rot16) & SMASK); };        // address of row byte x of this slot
        uint32_t next = s.word;
        if (!s.active) return next;
        const uint4 h = lds128(at(0));                                // cell | len, label, m | inline mask
        s.cell = static_cast<int32_t>(h.x);
        const int len = static_cast<int>(h.y & 0xffffu);
        // ---- trajectory column t, hash, label bitmap
        if (s.out_row == nullptr) {
            // filter modes without a trajectory buffer: the digest is all that is kept
        } else if (vec_store) {
            if constexpr (q == 3) {
                *reinterpret_cast<int4 *>(s.out_row + (t - 3)) = make_int4(s.b0, s.b1, s.b2, s.cell);
            } else {
                if constexpr (q == 0) s.b0 = s.cell; else if constexpr (q == 1) s.b1 = s.cell; else s.b2 = s.cell;
                if (t == nt) {
                    s.out_row[t - q] = s.b0;
                    if constexpr (q >= 1) s.out_row[t - q + 1] = s.b1;
                    if constexpr (q >= 2) s.out_row[t - q + 2] = s.b2;
                }
            }
        } else {
            s.out_row[t] = s.cell;
        }
        s.dg.visit(s.cell, (h.y >> 16) & 0xffu);
        if (t >= nt) return next;
        // ---- the key of this transition's uniform
        const uint32_t ua = (q & 1) ? s.x.z : s.x.x, ub = (q & 1) ? s.x.w : s.x.y;
        double u = 0.0;
        int key, t_def = 0;
        if constexpr (CDFM == kModeExact) {
            // t_def = ceil(u * 4096): code >= t_def <=> u <= code / 4096;  key = t_def - 1: code >= key <=> u <= (code + 1) / 4096
            if (native) {
                // u = mant * 2^-53, mant = (ua >> 5) * 2^26 + (ub >> 6): ceil(mant / 2^41) in integers
                t_def = static_cast<int>(ua >> 20) + ((((ua & 0xFFFE0u) | (ub >> 6)) != 0u) ? 1 : 0);
            } else {
                u = s.urow[t];
                t_def = static_cast<int>(ceil(u * 4096.0));
            }
            key = t_def - 1;
        } else {
            u = native ? bits_to_unit_double(ua, ub) : s.urow[t];
            key = static_cast<int>(ceil(u * 1000.0));               // smallest integer m with u <= (double)fl32(m / 1000)
            while (key > 0 && u <= static_cast<double>(__fdiv_rn(static_cast<float>(key - 1), 1000.0f))) --key;
            while (u > static_cast<double>(__fdiv_rn(static_cast<float>(key), 1000.0f))) ++key;
            key <<= 2;                                               // the codes are 4 m (rows_v2.cu): 8 significant bits in the upper plane
        }
        const int key_hi = key >> 4;
        auto settle = [&](int j) {                                   // undecided by the code: the float64 compare of :49
            if (native) u = bits_to_unit_double(ua, ub);
            const RowInfo ri = p.rows[s.cell];
            return u <= __ldg(static_cast<const double *>(p.cdf) + static_cast<int64_t>(ri.off_units) * kRowAlign + j);
        };
        int j = 0;
        bool in_slot = true;
        if constexpr (FLAVOR == kOversize) in_slot = (s.word >> kWordOffBits) * 32u <= static_cast<uint32_t>(SLOT);
        if (in_slot) {
            // ---- first crossing in the slot: lower bound on the upper plane, then the 12-bit codes
#pragma unroll
            for (int lv = DEPTH - 1; lv >= 0; --lv) {
                if (lv >= DEPTH - 2 && lv >= depth) continue;        // uniform: the top levels only exist for the longest rows a slot holds
                const int probe = j + (1 << lv) - 1;
                if (probe < len && lds8(slot_k | ((static_cast<uint32_t>(probe) + r20) & SMASK)) < key_hi) j += 1 << lv;
            }
            const uint32_t lo0 = r20 + static_cast<uint32_t>(len);
            auto code_at = [&](int jj) {
                const int hi = lds8(slot_k | ((static_cast<uint32_t>(jj) + r20) & SMASK));
                const int lo = lds8(slot_k | ((static_cast<uint32_t>(jj >> 1) + lo0) & SMASK));
                return (hi << 4) | ((lo >> ((jj & 1) * 4)) & 15);
            };
            // (reading both candidates of the short run of entries whose upper byte equals the key's at once, to keep the
            // warp converged, is slower: 59.1 vs 60.7 G at C4, profiles/r02_k2_experiments.md)
            if constexpr (CDFM == kModeExact) {
                while (j < len) {
                    const int code = code_at(j);
                    if (code >= t_def) break;                        // certain crossing
                    if (code == key && settle(j)) break;
                    ++j;
                }
            } else {
                while (j < len && code_at(j) < key) ++j;
            }
            // ---- next row word
            if (j < len) {
                const uint32_t mword = j < 32 ? h.z : h.w;          // inline mask, the half that holds bit j
                const uint32_t below = j < 32 ? 0u : static_cast<uint32_t>(__popc(h.z));
                if (j < 64 && ((mword >> (j & 31)) & 1u)) {
                    const uint32_t rank = below + static_cast<uint32_t>(__popc(mword & ((1u << (j & 31)) - 1u)));
                    next = lds32(at(static_cast<uint32_t>(v2_used_bytes(len)) + 4u * rank));
                } else {
                    next = ldg_word(p.edge4 + static_cast<size_t>(lds32(at(16))) * 4 + j);
                }
            }
        } else {
            // a block longer than the slot (a few long rows of a ragged graph: the slot is sized for the typical
            // row) is searched where it lies, in global memory; its header did fit
            const unsigned char *blk = p.rows2 + (static_cast<size_t>(s.word & kWordOffMask) << 5);
            const unsigned char *hi = blk + kV2Header;
            const unsigned char *lo = hi + len;
            for (int st = 1 << (log_n - 1); st >= 1; st >>= 1) {
                const int probe = j + st - 1;
                if (probe < len && static_cast<int>(hi[probe]) < key_hi) j += st;
            }
            auto code_at = [&](int jj) { return (static_cast<int>(hi[jj]) << 4) | ((static_cast<int>(lo[jj >> 1]) >> ((jj & 1) * 4)) & 15); };
            if constexpr (CDFM == kModeExact) {
                while (j < len) {
                    const int code = code_at(j);
                    if (code >= t_def) break;
                    if (code == key && settle(j)) break;
                    ++j;
                }
            } else {
                while (j < len && code_at(j) < key) ++j;
            }
            if (j < len) {
                const unsigned long long mask = (static_cast<unsigned long long>(h.w) << 32) | h.z;
                if (j < 64 && ((mask >> j) & 1ull)) {
                    const int rank = __popcll(mask & ((1ull << j) - 1ull));
                    next = *reinterpret_cast<const uint32_t *>(blk + v2_used_bytes(len) + 4 * rank);
                } else {
                    next = ldg_word(p.edge4 + static_cast<size_t>(*reinterpret_cast<const uint32_t *>(blk + 16)) * 4 + j);
                }
            }
        }
        if (j >= len && p.nocross) atomicAdd(p.nocross, 1ull);       // stays where it is (reference: IndexError at :49)
        return next;
    };

    for (int64_t wbase = static_cast<int64_t>(blockIdx.x) * nthreads + (threadIdx.x & ~31); wbase < p.n_walkers; wbase += stride) {
        V2Walker<MODE> s;
        s.w = wbase + lane;
        s.active = s.w < p.n_walkers;
        uint64_t gw = static_cast<uint64_t>(p.walker_begin + s.w);
        s.round = p.round;
        if (s.active) {                                               // else word = 0: the first block / nothing is copied
            int32_t root;
            if (p.replay) {
                const ReplayItem it = p.replay[s.w];
                gw = static_cast<uint64_t>(it.walker);
                s.round = it.round;
                root = it.root;
            } else {
                root = p.roots[gw / static_cast<uint64_t>(p.sims_per_root)];
            }
            s.word = p.word[root];
        }
        s.gw_lo = static_cast<uint32_t>(gw);
        s.gw_hi = static_cast<uint32_t>(gw >> 32);
        s.out_row = (p.out_seq && s.active) ? p.out_seq + s.w * p.ld : nullptr;
        s.urow = (!native && s.active) ? p.uniforms + s.w * static_cast<int64_t>(nt) : nullptr;
        issue_copy(s.word);
        if (native && nt > 0) s.x = philox4x32_10(make_uint4(s.gw_lo, s.gw_hi, 0u, s.round), pkey);
        int t = 0;

        // one step; Q = t % 4 is static.  Returns true after the last column has been written.
        auto step = [&](auto Q) -> bool {
            constexpr int q = decltype(Q)::value;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const uint32_t next = process(s, Q, t);
            if (t == nt) return true;
            s.word = next;
            ++t;
            __syncwarp();                                             // every lane is done with its slot
            issue_copy(next);
            if constexpr ((q & 1) != 0)                               // t is even now: the uniforms of transitions t, t + 1
                if (native) s.x = philox4x32_10(make_uint4(s.gw_lo, s.gw_hi, static_cast<uint32_t>(t >> 1), s.round), pkey);
            return false;
        };
        while (true) {
            if (step(IC<0>{})) break;
            if (step(IC<1>{})) break;
            if (step(IC<2>{})) break;
            if (step(IC<3>{})) break;
        }
        if (s.active) s.dg.finish(p, s.w, s.cell);
    }
}

// Tried on C4 and dropped in round 1 (profiles/r01_v2_experiments.md, r01_v2_flags.log): L2 cache hints (streaming
// trajectory stores, evict-first target gathers, evict-last row copies: no effect), a persisting-L2 access-policy
// window on the fused rows (24-34 G, much worse), two CTAs of 512..640 threads per SM (no gain from more walkers in
// flight), one or two extra sectors of inlined targets per row (51 / 48 vs 57 G at C4; +10 % at C3, where it is the
// default), row copies through registers (LDG.128 + STS.128: 35 vs 58 G), a bulk-copy (UBLKCP) flavour for rows
// above 256 bytes (23.1 G at C5 against 24.4 G for the 16-bit rows of walk_tma.cu: few entries of a 200-entry row
// fit the slack, so most steps still need the second gather).
constexpr int64_t kV2L1Cells = 16384;      // below: L1-allocating row copies (see walk_v2_kernel)

struct V2Config {
    int cpr;            // 16-byte chunks per slot
    int ctas;           // CTAs per SM
    int threads;
    int log_n;
    int flavor;         // v2::kUniform / kRagged / kOversize
    int l1_rows;        // row copies allocate in L1 (small graphs)
    size_t smem;
};

// variant: 0 = auto; 9500 + T/32 = one CTA of T threads per SM; 9700 + T/32 = two CTAs of T threads per SM
inline bool v2_config(const cyp_graph *g, int variant, int64_t n_walkers, int sm_count, V2Config *out) {
    if (!g->v2_ok) return false;
    // auto: two CTAs per SM where the rows come out of HBM (fused rows above half the L2) and the key is an integer (exact mode)
    int ctas = (g->cdf_mode == CYP_CDF_EXACT && g->rows2_bytes > (64ll << 20)) ? 2 : 1;
    if (variant > 9700) { ctas = 2; variant -= 200; }
    else if (variant > 9500) ctas = 1;
    const int block = g->v2_slot_sectors * 32;          // covers (almost) every row (cyp_graph::v2_slot_sectors, rows_v2.cu)
    if (block > 256) return false;                       // long rows: walk_tma.cu
    int cpr = 2;
    while (cpr * 16 < block) cpr <<= 1;
    const int stride = 16 * cpr;
    const int flavor = g->v2_max_sectors * 32 > stride ? v2::kOversize : (g->v2_uniform ? v2::kUniform : v2::kRagged);
    int l1_rows = g->n < kV2L1Cells ? 1 : 0;
    if (const char *e = getenv("CYP_V2_L1")) l1_rows = atoi(e) ? 1 : 0;
    if (!(cpr == 8 && flavor == v2::kUniform && !l1_rows)) ctas = 1;       // the only shape instantiated with two CTAs per SM
    const size_t budget = (226 * 1024 - 1024) / ctas - (ctas > 1 ? 1024 : 0);
    int threads = static_cast<int>(budget / stride) / 32 * 32;
    const int tmax = ctas == 2 ? 640 : 1024;
    if (threads > tmax) threads = tmax;
    if (variant > 9500) {
        threads = (variant - 9500) * 32 < threads ? (variant - 9500) * 32 : threads;
    } else if (n_walkers > 0 && sm_count > 0) {
        // wave balancing: every warp runs as many 32-walker batches as with the largest CTA, with no more warps than that needs
        const int64_t batches = (n_walkers + 31) / 32;
        const int64_t per_round = static_cast<int64_t>(sm_count) * ctas * (threads / 32);
        const int64_t rounds = (batches + per_round - 1) / per_round;
        const int64_t warps = (batches + static_cast<int64_t>(sm_count) * ctas * rounds - 1) / (static_cast<int64_t>(sm_count) * ctas * rounds);
        if (warps * 32 < threads) threads = static_cast<int>(warps) * 32;
    }
    if (threads < 32) threads = 32;
    int log_n = 1;
    while ((1 << log_n) < g->max_len) ++log_n;
    *out = V2Config{cpr, ctas, threads, log_n, flavor, l1_rows, static_cast<size_t>(threads) * stride};
    return true;
}

template <int CPR, int CDFM, int MODE, int FLAVOR, bool L1ROWS, int CTAS = 1>
static int launch_v2_kernel(const WalkParams &p, const V2Config &c, int sm_count, cudaStream_t stream) {
    auto kernel = walk_v2_kernel<CPR, CDFM, MODE, FLAVOR, L1ROWS, CTAS>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute of the function
    CYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c.smem)));
    const int64_t need = (p.n_walkers + c.threads - 1) / c.threads;
    const unsigned grid = static_cast<unsigned>(need < sm_count * CTAS ? need : sm_count * CTAS);       // persistent CTAs
    if (grid == 0) return CYP_OK;
    static const bool trace = getenv("CYP_TRACE") != nullptr;
    if (trace) {
        int dev = -1;
        cudaGetDevice(&dev);
        fprintf(stderr, "[cyp trace] walk_v2: device %d grid %u threads %d smem %zu flavor %d walkers %lld begin %lld sims/root %lld nt %d\n", dev, grid,
                c.threads, c.smem, FLAVOR, static_cast<long long>(p.n_walkers), static_cast<long long>(p.walker_begin),
                static_cast<long long>(p.sims_per_root), p.nt);
    }
    kernel<<<grid, c.threads, c.smem, stream>>>(p, c.log_n);
    CYP_CUDA(cudaGetLastError());
    return CYP_OK;
}

template <int CPR, int CDFM, int MODE>
static int launch_v2_flavor(const WalkParams &p, const V2Config &c, int sm_count, cudaStream_t stream) {
    // small graphs (L1-allocating copies) always take the predicated copies: 5 instantiations, not 6
    if (c.l1_rows) {
        return c.flavor == v2::kOversize ? launch_v2_kernel<CPR, CDFM, MODE, v2::kOversize, true>(p, c, sm_count, stream)
                                         : launch_v2_kernel<CPR, CDFM, MODE, v2::kRagged, true>(p, c, sm_count, stream);
    }
    switch (c.flavor) {
        case v2::kUniform:
            if constexpr (CPR == 8) {
                if (c.ctas == 2) return launch_v2_kernel<CPR, CDFM, MODE, v2::kUniform, false, 2>(p, c, sm_count, stream);
            }
            return launch_v2_kernel<CPR, CDFM, MODE, v2::kUniform, false>(p, c, sm_count, stream);
        case v2::kRagged: return launch_v2_kernel<CPR, CDFM, MODE, v2::kRagged, false>(p, c, sm_count, stream);
        default: return launch_v2_kernel<CPR, CDFM, MODE, v2::kOversize, false>(p, c, sm_count, stream);
    }
}

template <int CPR, int CDFM>
static int launch_v2_mode(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    switch (mode) {
        case kModePlain: return launch_v2_flavor<CPR, CDFM, kModePlain>(p, c, sm_count, stream);
        case kModeFilter0: return launch_v2_flavor<CPR, CDFM, kModeFilter0>(p, c, sm_count, stream);
        case kModeFilter1: return launch_v2_flavor<CPR, CDFM, kModeFilter1>(p, c, sm_count, stream);
        case kModeFilter4: return launch_v2_flavor<CPR, CDFM, kModeFilter4>(p, c, sm_count, stream);
    }
    return fail(CYP_E_INVALID, "launch_walk_v2: bad mode");
}

// one translation unit per CDF mode (walk_v2_exact.cu, walk_v2_ref.cu): halves the longest compile of the build
template <int CDFM>
static int launch_v2_cpr(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    switch (c.cpr) {
        case 2: return launch_v2_mode<2, CDFM>(p, mode, c, sm_count, stream);
        case 4: return launch_v2_mode<4, CDFM>(p, mode, c, sm_count, stream);
        case 8: return launch_v2_mode<8, CDFM>(p, mode, c, sm_count, stream);
        case 16: return launch_v2_mode<16, CDFM>(p, mode, c, sm_count, stream);
    }
    return kNotApplicable;
}

int launch_walk_v2_exact(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream);
int launch_walk_v2_ref(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream);

}  // namespace cyp
