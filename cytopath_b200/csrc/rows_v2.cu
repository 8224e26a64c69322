// K1b: the fused-row layout of common.cuh ("v2"), built from K1's output.
//
// Why this layout (profiles/r01_gather_ceilings.log, r01_footprint_sweep_*.log): at atlas scale K2 is bound
// by DRAM random accesses -- per step one 128-byte row of 16-bit codes (L2 hit rate ~55 %) plus one 32-byte
// sector holding a 16-byte edge record out of a 1 GB array (hit rate ~15 %); the edge-record gathers alone
// are ~30 G/s against a measured ceiling of ~40 G random sector reads/s.  Transition probabilities are
// skewed (scvelo: exp(cosine * scale)), so a row's few most probable entries take most of the steps: their
// targets are stored in the slack of the row's own block and those steps need no second gather at all.
// 12-bit codes (two planes: 8 + 4 bits) make room for that without growing the block; the remaining
// targets shrink from 16-byte records to 4-byte row words.
//
// The kernel is one warp per row: codes from the CDF K1 wrote (bit-equal to matrix_prep, :11-31), targets
// translated to row words, the m most probable entries selected by m warp-argmax rounds.
#include <cstdlib>

#include "common.cuh"

namespace cyp {

namespace {

template <typename CdfT>
__device__ __forceinline__ int code12(CdfT c, int *bad) {
    if constexpr (sizeof(CdfT) == 8) {
        if (!(c >= 0.0)) { *bad = 1; return 0; }
        const double x = floor(c * 4096.0);
        return static_cast<int>(x > 4095.0 ? 4095.0 : x);
    } else {
        // the stored value is fl32(m / 1000) exactly (:31).  The code is 4 m: m <= 1001 needs 10 bits, and shifted
        // by two the upper-8-bit plane (what the binary search runs on) tells 251 levels apart instead of 63
        const float m = rintf(__fmul_rn(c, 1000.0f));
        if (!(m >= 0.0f && m <= 1023.0f)) { *bad = 1; return 0; }
        return static_cast<int>(m) << 2;
    }
}

template <typename CdfT>
__global__ void __launch_bounds__(256)
build_rows_v2_kernel(int64_t row_begin, int64_t row_end, int64_t n_cells, const RowInfo *__restrict__ rows, const CdfT *__restrict__ cdf, const double *__restrict__ val,
                     const int32_t *__restrict__ col, const uint32_t *__restrict__ word, const uint32_t *__restrict__ ebase,
                     int extra, unsigned char *__restrict__ rows2, uint32_t *__restrict__ edge4, int *__restrict__ bad_flag) {
    const int64_t r = row_begin + ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= row_end) return;
    const RowInfo ri = rows[r];
    const int len = static_cast<int>(ri.len);
    const int64_t off = static_cast<int64_t>(ri.off_units) * kRowAlign;
    const uint32_t w = word[r];
    unsigned char *blk = rows2 + static_cast<size_t>(w & kWordOffMask) * 32;
    const int used = v2_used_bytes(len);
    const int nsect = v2_block_sectors(len, extra);          // the word's field saturates at kV2MaxSectors
    int m = (nsect * 32 - used) / 4;
    if (m > len) m = len;
    if (m > 64) m = 64;
    uint32_t *e4 = edge4 + static_cast<size_t>(ebase[r]) * 4;
    int bad = 0;
    // codes and targets
    for (int j0 = 0; j0 < len; j0 += 32) {
        const int j = j0 + lane;
        int code = 0;
        if (j < len) {
            code = code12<CdfT>(cdf[off + j], &bad);
            blk[kV2Header + j] = static_cast<unsigned char>(code >> 4);
            const int32_t cc = col[off + j];
            e4[j] = (cc >= 0 && cc < n_cells) ? word[cc] : 0u;         // a column out of range fails the create call (graph.cu)
        }
        const int lo = code & 15;
        const int lo_next = __shfl_down_sync(0xffffffffu, lo, 1);       // lane parity == entry parity (j0 is a multiple of 32)
        if ((lane & 1) == 0 && j < len) blk[kV2Header + len + (j >> 1)] = static_cast<unsigned char>(lo | ((j + 1 < len ? lo_next : 0) << 4));
    }
    // the m most probable entries among the first 64 (ties: the earlier entry)
    double p0 = (lane < len) ? val[off + lane] : -1.0;
    double p1 = (lane + 32 < len) ? val[off + lane + 32] : -1.0;
    unsigned long long mask = 0ull;
    for (int k = 0; k < m; ++k) {
        double best = p0 >= p1 ? p0 : p1;
        int bj = p0 >= p1 ? lane : lane + 32;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        mask |= 1ull << bj;
        if (bj == lane) p0 = -2.0;
        if (bj == lane + 32) p1 = -2.0;
    }
    uint32_t *inl = reinterpret_cast<uint32_t *>(blk + used);
    for (int j = lane; j < 64 && j < len; j += 32)
        if ((mask >> j) & 1ull) {
            const int32_t cc = col[off + j];
            inl[__popcll(mask & ((1ull << j) - 1ull))] = (cc >= 0 && cc < n_cells) ? word[cc] : 0u;
        }
    if (lane == 0) {
        *reinterpret_cast<int32_t *>(blk) = static_cast<int32_t>(r);
        *reinterpret_cast<uint16_t *>(blk + 4) = static_cast<uint16_t>(len);
        blk[6] = 0;
        blk[7] = static_cast<unsigned char>(m);
        *reinterpret_cast<unsigned long long *>(blk + 8) = mask;
        *reinterpret_cast<uint32_t *>(blk + 16) = ebase[r];
    }
    if (bad) atomicOr(bad_flag, 1);
}

}  // namespace

// ---- host side: cyp_graph_create drives these (graph.cu), chunk by chunk behind the CSR upload ----

// The fused-row plan from the device-side row layout (layout.cu).  Returns false when the layout does not apply.
void v2_candidates(int *extra0, int *extra1) {
    // Extra sectors of inlined targets per row.  More inlining means fewer second gathers but more bytes per step:
    // on B200 it pays while the fused rows stay L2 resident (C3, 100k cells: 87 vs 79 G walker-steps/s) and costs
    // at atlas scale, where DRAM bytes per step are what counts (C4: 51 vs 57 G).  CYP_V2_EXTRA overrides.
    *extra0 = 0;
    *extra1 = 1;
    if (const char *e = getenv("CYP_V2_EXTRA")) {
        const int x = atoi(e);
        if (x >= 0) *extra0 = *extra1 = x;
    }
}

bool v2_plan(const cyp_graph *g, const RowLayoutSummary &rows, const int extra[2], V2Build *b) {
    b->planned = false;
    if (g->max_len > 65535) return false;
    if (const char *off = getenv("CYP_NO_V2")) if (off[0] == '1') return false;
    const int64_t n = g->n;
    // one extra sector per row while the fused rows then still fit 32 MB (L2 resident), else none
    const int which = (extra[0] == extra[1]) ? 0 : (((rows.sectors[0] + n) * 32 <= (32ll << 20)) ? 1 : 0);
    const int64_t sectors = rows.sectors[which], units = rows.units4;
    if (sectors > static_cast<int64_t>(kWordOffMask) - kV2MaxSectors || units >= (1ll << 32) - 1) return false;
    const int max_sectors = rows.max_sectors[which];
    const int64_t *hist = rows.hist[which];
    // slot of the step kernel: the smallest power of two (1..8 sectors) that holds 97 % of the rows; the few longer
    // rows of a ragged graph are then searched in global memory instead of making every slot as large as the longest
    int slot = max_sectors;
    for (int cand = 1; cand <= 8; cand <<= 1) {
        int64_t covered = 0;
        for (int k = 0; k <= cand && k <= kV2MaxSectors; ++k) covered += hist[k];
        if (cand >= max_sectors || covered * 100 >= n * 97) { slot = cand < max_sectors ? cand : max_sectors; break; }
    }
    b->slot_sectors = slot;
    {   // blocks shorter than the slot: when they are rare the step kernel copies whole slots unconditionally
        int64_t shorter = 0;
        for (int k = 0; k < slot && k <= kV2MaxSectors; ++k) shorter += hist[k];
        b->uniform = shorter * 10 <= n ? 1 : 0;
    }
    b->which = which;
    b->extra = extra[which];
    b->sectors = sectors;
    b->units = units;
    b->max_sectors = max_sectors;
    b->rows_bytes = static_cast<size_t>(sectors + kV2MaxSectors + 1) * 32;     // slack: a warp may copy a full slot past the last block
    b->planned = true;
    return true;
}

void v2_release(cyp_graph *g, V2Build *b) {
    dev_free(b->d_ebase);
    dev_free(b->d_bad);
    b->d_ebase = nullptr;
    b->d_bad = nullptr;
    if (!g->v2_ok) {
        dev_free(g->d_rows2_base); dev_free(g->d_edge4); dev_free(g->d_word);
        g->d_rows2 = nullptr; g->d_rows2_base = nullptr; g->d_edge4 = nullptr; g->d_word = nullptr;
    }
}

// Device arrays (default stream).
cudaError_t v2_alloc(cyp_graph *g, V2Build *b) {
    // CYP_ROWS2_OFFSET_KB (developer switch, multiples of 4): shifts the fused rows inside their allocation; which L2
    // slices serve the hot rows depends on it (scripts/placement_scan.py)
    size_t offset = 0;
    if (const char *o = getenv("CYP_ROWS2_OFFSET_KB")) offset = static_cast<size_t>(atol(o)) << 10;
    cudaError_t e = dev_alloc(&g->d_rows2_base, b->rows_bytes + offset);
    g->d_rows2 = g->d_rows2_base ? g->d_rows2_base + offset : nullptr;
    if (e == cudaSuccess) e = dev_alloc(&g->d_edge4, sizeof(uint32_t) * 4 * static_cast<size_t>(b->units + 1));
    if (e == cudaSuccess) e = dev_alloc(&g->d_word, sizeof(uint32_t) * g->n);
    if (e == cudaSuccess) e = dev_alloc(&b->d_ebase, sizeof(uint32_t) * g->n);
    if (e == cudaSuccess) e = dev_alloc(&b->d_bad, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_bad, 0, sizeof(int), nullptr);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_rows2, 0, b->rows_bytes, nullptr);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_edge4, 0, sizeof(uint32_t) * 4 * static_cast<size_t>(b->units + 1), nullptr);   // row padding
    return e;
}

// K1b for rows [row_begin, row_end) on `stream` (K1 has finished these rows on the same stream).
cudaError_t v2_launch(cyp_graph *g, V2Build *b, int64_t row_begin, int64_t row_end, cudaStream_t stream) {
    const int64_t rows = row_end - row_begin;
    if (rows <= 0) return cudaSuccess;
    const int threads = 256;
    const unsigned grid = static_cast<unsigned>((rows * 32 + threads - 1) / threads);
    if (g->cdf_mode == CYP_CDF_EXACT)
        build_rows_v2_kernel<double><<<grid, threads, 0, stream>>>(row_begin, row_end, g->n, g->d_row, static_cast<const double *>(g->d_cdf), g->d_val,
                                                                   g->d_col, g->d_word, b->d_ebase, b->extra, g->d_rows2, g->d_edge4, b->d_bad);
    else
        build_rows_v2_kernel<float><<<grid, threads, 0, stream>>>(row_begin, row_end, g->n, g->d_row, static_cast<const float *>(g->d_cdf), g->d_val,
                                                                  g->d_col, g->d_word, b->d_ebase, b->extra, g->d_rows2, g->d_edge4, b->d_bad);
    return cudaGetLastError();
}

// After the build stream has been synchronised: commits the layout (g->v2_ok = 1) unless a code did not fit.
cudaError_t v2_finish(cyp_graph *g, V2Build *b, bool codes_ok) {
    int bad = 0;
    cudaError_t e = cudaMemcpy(&bad, b->d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    if (!bad && codes_ok) {
        g->v2_sectors = b->sectors;
        g->v2_edge_units = b->units;
        g->v2_max_sectors = b->max_sectors;
        g->v2_slot_sectors = b->slot_sectors;
        g->v2_uniform = b->uniform;
        g->rows2_bytes = static_cast<int64_t>(b->rows_bytes);
        g->device_bytes += static_cast<int64_t>(b->rows_bytes + sizeof(uint32_t) * 4 * (b->units + 1) + sizeof(uint32_t) * g->n);
        g->v2_ok = 1;
    }
    v2_release(g, b);
    return cudaSuccess;
}

}  // namespace cyp
