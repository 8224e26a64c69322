// K2: the Markov-chain step loop (reference cytopath/markov_chain_sampler.py:45-51).
//
// Work decomposition.  G lanes of a warp (G = 1, 2, 4, 8, 16 or 32, chosen from the mean row
// length) cooperate on one walker and keep it for its whole trajectory:
//   1. the row descriptor {offset, len} of the current cell is already in registers (it came
//      with the previous step's edge record; the root's is loaded once),
//   2. the row's CDF segment is fetched with UNROLL independent 16-byte loads per lane
//      (G * UNROLL * 16 bytes per chunk; a typical row is one chunk, so every sector of the
//      row is requested once and all requests are in flight together),
//   3. each lane finds its first element with u <= cdf (ties go to the bin, `<=` at :49),
//      the group takes the minimum index with redux.sync (first crossing, :49),
//   4. one 16-byte load of the edge record at that position gives the next cell (:48) and its
//      row descriptor.
// Uniforms come from Philox4x32-10 keyed by (seed; walker, t/2, round): lane i of the group
// computes the pair for transitions 2i, 2i+1 of the current batch of 2G transitions and the
// group reads them with shuffles, so no lane repeats another lane's Philox work.
// Trajectories are written in groups of G consecutive int32 (one 32-byte sector when G = 8).
//
// Roofline: HBM.  Algorithmic bytes per walker-step = sizeof(cdf) * row_len + 8 + 4 + 4
// (SURVEY.md 8d).  The kernel is latency/occupancy bound below that; see DESIGN.md.
#include <cstdio>
#include <cstdlib>

#include "walk.cuh"

namespace cyp {

constexpr int kWalkThreads = 256;

template <typename CdfT> struct Vec16;
template <> struct Vec16<double> {
    using type = double2;
    static constexpr int n = 2;
    static __device__ __forceinline__ void unpack(const double2 &v, double (&e)[2]) { e[0] = v.x; e[1] = v.y; }
    static __device__ __forceinline__ double2 never() { return make_double2(-1.0, -1.0); }
};
template <> struct Vec16<float> {
    using type = float4;
    static constexpr int n = 4;
    static __device__ __forceinline__ void unpack(const float4 &v, double (&e)[4]) {
        e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;      // float -> double is exact (:49 compares in float64)
    }
    static __device__ __forceinline__ float4 never() { return make_float4(-1.f, -1.f, -1.f, -1.f); }
};

template <int G>
__device__ __forceinline__ unsigned group_min(unsigned mask, unsigned v) {
    if constexpr (G == 1) return v;
    else return __reduce_min_sync(mask, v);
}

template <typename CdfT, int G, int UNROLL, int MODE>
__global__ void __launch_bounds__(kWalkThreads) walk_kernel(const WalkParams p) {
    using V = Vec16<CdfT>;
    constexpr int VEC = V::n;
    constexpr int ALIGN = kRowAlign;
    constexpr int CHUNK = G * VEC * UNROLL;

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int64_t n_groups = static_cast<int64_t>(gridDim.x) * (kWalkThreads / G);
    const CdfT *__restrict__ cdf = static_cast<const CdfT *>(p.cdf);
    const EdgeRec *__restrict__ edge = p.edge;
    const RowInfo *__restrict__ rows = p.rows;
    const int nt = p.nt;

    for (int64_t w = (static_cast<int64_t>(blockIdx.x) * kWalkThreads + threadIdx.x) / G; w < p.n_walkers; w += n_groups) {
        uint64_t gw = static_cast<uint64_t>(p.walker_begin + w);
        uint32_t round = p.round;
        int32_t cur;
        if (p.replay) {
            const ReplayItem it = p.replay[w];
            gw = static_cast<uint64_t>(it.walker);
            round = it.round;
            cur = it.root;
        } else {
            cur = p.roots[gw / static_cast<uint64_t>(p.sims_per_root)];
        }
        RowInfo ri = rows[cur];
        int32_t *__restrict__ out_row = p.out_seq ? p.out_seq + w * p.ld : nullptr;
        const double *__restrict__ urow = p.uniforms ? p.uniforms + w * static_cast<int64_t>(nt) : nullptr;

        int32_t mine = cur;                               // lane (c % G) holds column c until the group flushes
        if (G == 1 && out_row) out_row[0] = cur;
        WalkerDigest<MODE> dg;
        dg.visit(cur, WalkerDigest<MODE>::MASKW > 0 ? p.code8[cur] : 0u);
        unsigned misses = 0;
        double u_even = 0.0, u_odd = 0.0;

        for (int t = 0; t < nt; ++t) {
            // ---- uniform for transition t
            double u;
            if (urow) {
                u = urow[t];
            } else {
                const int r = t & (2 * G - 1);
                if (r == 0) walker_uniform_pair(p.seed, round, gw, static_cast<uint32_t>((t >> 1) + gl), u_even, u_odd);
                const double mine_u = (r & 1) ? u_odd : u_even;
                u = (G == 1) ? mine_u : __shfl_sync(gmask, mine_u, r >> 1, G);
            }
            // ---- row descriptor and first crossing
            const int len = static_cast<int>(ri.len);
            const int64_t off = static_cast<int64_t>(ri.off_units) * ALIGN;
            const CdfT *__restrict__ rowp = cdf + off;
            unsigned found = 0xffffffffu;
            for (int base = 0; base < len; base += CHUNK) {
                typename V::type v[UNROLL];
#pragma unroll
                for (int i = 0; i < UNROLL; ++i) {
                    const int idx = base + (i * G + gl) * VEC;
                    v[i] = (idx < len) ? __ldg(reinterpret_cast<const typename V::type *>(rowp + idx)) : V::never();
                }
#pragma unroll
                for (int i = UNROLL - 1; i >= 0; --i) {   // descending so the smallest index wins
                    double e[VEC];
                    V::unpack(v[i], e);
                    const int idx = base + (i * G + gl) * VEC;
#pragma unroll
                    for (int k = VEC - 1; k >= 0; --k)
                        if (u <= e[k]) found = static_cast<unsigned>(idx + k);
                }
                found = group_min<G>(gmask, found);
                if (found != 0xffffffffu) break;
            }
            unsigned code = 0;
            if (found != 0xffffffffu) {
                const EdgeRec e = load_edge(edge + off + found);
                cur = e.col;
                ri = RowInfo{e.off_units, e.len};
                code = e.code;
            } else {
                ++misses;
                if constexpr (WalkerDigest<MODE>::MASKW > 0) code = p.code8[cur];
            }
            // ---- record column c = t + 1
            const int c = t + 1;
            if ((c & (G - 1)) == gl) mine = cur;
            if ((c & (G - 1)) == G - 1 || c == nt) {
                const int cbase = c & ~(G - 1);
                if (out_row && cbase + gl <= c) out_row[cbase + gl] = mine;
            }
            dg.visit(cur, code);
        }
        if (nt == 0 && G > 1 && gl == 0 && out_row) out_row[0] = mine;
        if (gl == 0) {
            if (misses && p.nocross) atomicAdd(p.nocross, static_cast<unsigned long long>(misses));
            dg.finish(p, w, cur);
        }
    }
}

// ---- dispatch -----------------------------------------------------------------------
template <typename CdfT, int G, int UNROLL, int MODE>
static int launch_one(const WalkParams &p, cudaStream_t stream) {
    auto kernel = walk_kernel<CdfT, G, UNROLL, MODE>;
    const int sm_count = current_sm_count();
    if (sm_count <= 0) return fail(CYP_E_CUDA, "launch_walk: no CUDA device");
    int blocks_per_sm = 0;
    CYP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kWalkThreads, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t groups_per_block = kWalkThreads / G;
    const int64_t need = (p.n_walkers + groups_per_block - 1) / groups_per_block;
    const int64_t resident = static_cast<int64_t>(sm_count) * blocks_per_sm;   // grid-stride over walkers
    const unsigned grid = static_cast<unsigned>(need < resident ? need : resident);
    if (grid == 0) return CYP_OK;
    kernel<<<grid, kWalkThreads, 0, stream>>>(p);
    CYP_CUDA(cudaGetLastError());
    return CYP_OK;
}

template <typename CdfT, int G, int UNROLL>
static int launch_mode(const WalkParams &p, WalkMode mode, cudaStream_t stream) {
    switch (mode) {
        case kModePlain: return launch_one<CdfT, G, UNROLL, kModePlain>(p, stream);
        case kModeFilter0: return launch_one<CdfT, G, UNROLL, kModeFilter0>(p, stream);
        case kModeFilter1: return launch_one<CdfT, G, UNROLL, kModeFilter1>(p, stream);
        case kModeFilter4: return launch_one<CdfT, G, UNROLL, kModeFilter4>(p, stream);
    }
    return fail(CYP_E_INVALID, "launch_walk: bad mode");
}

template <typename CdfT>
static int launch_variant(const WalkParams &p, WalkVariant v, WalkMode mode, cudaStream_t stream) {
#define CYP_V(G_, U_) \
    if (v.lanes == G_ && v.unroll == U_) return launch_mode<CdfT, G_, U_>(p, mode, stream);
    CYP_V(1, 4) CYP_V(2, 4) CYP_V(4, 4) CYP_V(8, 4) CYP_V(16, 4) CYP_V(32, 4)
#undef CYP_V
    return fail(CYP_E_INVALID, "launch_walk: unknown kernel variant (lanes*100+unroll)");
}

// long rows (C5, k = 200, 352-byte blocks): 16-bit rows + edge records through walk_tma.cu's bulk copies do 24.4 G
// walker-steps/s, the fused rows 21.2 G with cooperative copies and 23.1 G with a bulk flavour of their own (round 1,
// since removed: few entries of a 200-entry row fit the slack, so most steps still need the second gather)
bool walk_auto_is_v2(const cyp_graph *g) { return g->v2_ok && g->v2_slot_sectors * 32 <= 256; }

WalkVariant pick_walk_variant(const cyp_graph *g, int variant) {
    if (variant > 0) return WalkVariant{variant / 100, variant % 100};
    // one chunk (lanes * unroll * 16 bytes) should cover a typical row
    const int vec = (g->cdf_mode == CYP_CDF_EXACT) ? 2 : 4;
    const double want = g->mean_len;
    static const WalkVariant table[] = {{1, 4}, {2, 4}, {4, 4}, {8, 4}, {16, 4}, {32, 4}};
    for (const WalkVariant &v : table)
        if (v.lanes * v.unroll * vec >= want) return v;
    return WalkVariant{32, 4};
}

int launch_walk(const cyp_graph *g, const WalkParams &p_in, WalkMode mode, int variant, cudaStream_t stream) {
    if (variant == 0 && g->placement_state < 2 && !p_in.replay && !p_in.uniforms && p_in.nt >= 16 && p_in.n_walkers >= 65536) {
        // the SECOND large walk on this graph picks the placement of the fused rows with this very workload (placement.cu):
        // the ~1.5 ms of probing pay back over the further rounds of a sampling() run, not on a graph that is walked once
        if (++g->placement_state == 2) {
            cyp_graph *gm = const_cast<cyp_graph *>(g);
            const int rc = calibrate_placement(gm, p_in, &gm->placement_ms_before, &gm->placement_ms_after, &gm->placement_moves);
            if (rc != CYP_OK) return rc;
        }
    }
    if (variant >= 9500 || (variant == 0 && walk_auto_is_v2(g))) {
        const int rc = launch_walk_v2(g, p_in, mode, variant, stream);
        if (rc != kNotApplicable) return rc;
        return fail(CYP_E_INVALID, "launch_walk: the fused-row kernel (variant 95xx) does not apply to this graph");
    }
    {                              // every other flavour reads the 16-byte edge records (and the 16-bit codes)
        const int rc = ensure_edge_records(g);
        if (rc != CYP_OK) return rc;
    }
    WalkParams p = p_in;
    p.edge = g->d_edge;
    if (variant == 0 || variant >= 9000) return launch_walk_tma(g, p, mode, variant, stream);
    const WalkVariant v = pick_walk_variant(g, variant);
    if (g->cdf_mode == CYP_CDF_EXACT) return launch_variant<double>(p, v, mode, stream);
    return launch_variant<float>(p, v, mode, stream);
}

}  // namespace cyp

using namespace cyp;

extern "C" const char *cyp_walk_variant_name(const cyp_graph *g, int variant) {
    static thread_local char buf[96];
    if (!g) return "";
    if (variant >= 9500 || (variant == 0 && walk_auto_is_v2(g))) {
        const char *name = walk_v2_name(g, variant);
        return name ? name : "not applicable";
    }
    if (variant >= 9000 || variant == 0) return walk_tma_name(g, variant);
    const WalkVariant v = pick_walk_variant(g, variant);
    snprintf(buf, sizeof(buf), "walk_kernel<%s,lanes=%d,unroll=%d>", g->cdf_mode == CYP_CDF_EXACT ? "f64" : "f32", v.lanes,
             v.unroll);
    return buf;
}

extern "C" int cyp_walk_device(const cyp_graph *g, const int32_t *d_root_cells, int64_t n_roots, int64_t sims_per_root,
                               int64_t walker_begin, int64_t n_walkers, int32_t n_transitions, uint64_t seed,
                               uint32_t round, const double *d_uniforms, int32_t *d_out_seq,
                               unsigned long long *d_nocrossing, int variant, void *stream) {
    CYP_REQUIRE(g != nullptr, "cyp_walk_device: NULL graph");
    CYP_REQUIRE(d_root_cells && d_out_seq, "cyp_walk_device: NULL buffer");
    CYP_REQUIRE(n_roots > 0 && sims_per_root > 0 && n_walkers >= 0 && n_transitions >= 0 && walker_begin >= 0,
                "cyp_walk_device: bad sizes");
    CYP_REQUIRE(walker_begin + n_walkers <= n_roots * sims_per_root, "cyp_walk_device: walker range exceeds n_roots*sims_per_root");
    DeviceGuard guard(g->device);
    WalkParams p{};
    p.rows = g->d_row;
    p.cdf = g->d_cdf;
    p.edge = g->d_edge;
    p.roots = d_root_cells;
    p.sims_per_root = sims_per_root;
    p.walker_begin = walker_begin;
    p.n_walkers = n_walkers;
    p.nt = n_transitions;
    p.round = round;
    p.seed = seed;
    p.uniforms = d_uniforms;
    p.out_seq = d_out_seq;
    p.ld = static_cast<int64_t>(n_transitions) + 1;
    p.nocross = d_nocrossing;
    return launch_walk(g, p, kModePlain, variant, static_cast<cudaStream_t>(stream));
}

extern "C" int cyp_walk(const cyp_graph *g, const int32_t *h_root_cells, int64_t n_roots, int64_t sims_per_root,
                        int64_t walker_begin, int64_t n_walkers, int32_t n_transitions, uint64_t seed, uint32_t round,
                        const double *h_uniforms, int32_t *h_out_seq, int64_t *n_nocrossing) {
    CYP_REQUIRE(g != nullptr, "cyp_walk: NULL graph");
    CYP_REQUIRE(h_root_cells && h_out_seq, "cyp_walk: NULL buffer");
    CYP_REQUIRE(n_roots > 0 && n_walkers >= 0 && n_transitions >= 0, "cyp_walk: bad sizes");
    for (int64_t i = 0; i < n_roots; ++i)
        CYP_REQUIRE(h_root_cells[i] >= 0 && h_root_cells[i] < g->n, "cyp_walk: root cell out of range");
    DeviceGuard guard(g->device);
    int32_t *d_roots = nullptr, *d_seq = nullptr;
    double *d_u = nullptr;
    unsigned long long *d_miss = nullptr;
    const size_t seq_bytes = sizeof(int32_t) * static_cast<size_t>(n_walkers) * (n_transitions + 1);
    int rc = CYP_OK;
    auto chk = [&](cudaError_t err, const char *what) {
        if (err != cudaSuccess && rc == CYP_OK)
            rc = fail(err == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_walk: ") + what + ": " + cudaGetErrorString(err));
        return err == cudaSuccess;
    };
    unsigned long long miss = 0;
    if (chk(dev_alloc(&d_roots, sizeof(int32_t) * n_roots), "cudaMalloc roots") &&
        chk(dev_alloc(&d_seq, seq_bytes ? seq_bytes : 4), "cudaMalloc seq") &&
        chk(dev_alloc(&d_miss, sizeof(unsigned long long)), "cudaMalloc counter") &&
        chk(cudaMemset(d_miss, 0, sizeof(unsigned long long)), "cudaMemset") &&
        chk(cudaMemcpy(d_roots, h_root_cells, sizeof(int32_t) * n_roots, cudaMemcpyHostToDevice), "H2D roots")) {
        bool ok = true;
        if (h_uniforms && n_walkers * n_transitions > 0) {
            const size_t ub = sizeof(double) * static_cast<size_t>(n_walkers) * n_transitions;
            ok = chk(dev_alloc(&d_u, ub), "cudaMalloc uniforms") && chk(cudaMemcpy(d_u, h_uniforms, ub, cudaMemcpyHostToDevice), "H2D uniforms");
        }
        if (ok) {
            rc = cyp_walk_device(g, d_roots, n_roots, sims_per_root, walker_begin, n_walkers, n_transitions, seed, round,
                                 d_u, d_seq, d_miss, 0, nullptr);
            if (rc == CYP_OK) {
                chk(cudaDeviceSynchronize(), "walk kernel");
                chk(cudaMemcpy(h_out_seq, d_seq, seq_bytes, cudaMemcpyDeviceToHost), "D2H seq");
                chk(cudaMemcpy(&miss, d_miss, sizeof(miss), cudaMemcpyDeviceToHost), "D2H counter");
            }
        }
    }
    dev_free(d_roots);
    dev_free(d_seq);
    dev_free(d_u);
    dev_free(d_miss);
    if (n_nocrossing) *n_nocrossing = static_cast<int64_t>(miss);
    if (rc == CYP_OK && miss) return fail(CYP_E_NOCROSSING, "cyp_walk: a uniform exceeded its row's last CDF value (reference: IndexError at markov_chain_sampler.py:49)");
    return rc;
}
