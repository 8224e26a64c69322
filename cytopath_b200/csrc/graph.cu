// a1: matrix_prep (reference cytopath/markov_chain_sampler.py:11-31) on the device.
//
// K1 `build_cdf_kernel`: segmented inclusive prefix sum over the CSR values.  The sum of a
// row is taken strictly left to right in float64 (np.cumsum, :19), so the result is
// bit-identical to numpy; a tree scan would re-associate the additions and is not used.
// Reference mode then narrows to float32 (:23,:28) and rounds to 3 decimals with
// rintf(c*1000)/1000 in float32 (np.round(decimals=3), :31).  Output goes to the
// sector-aligned row layout described in common.cuh.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cyp {

static thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

int current_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_dev = dev;
    }
    return cached;
}

// Blocks of 64 MB and more are additionally kept in an exact-size cache when they are freed: the CUDA pool hands
// memory back quickly, but for multi-GB requests it often has to re-map physical pages into a new contiguous
// virtual range (3-5 ms per 2.5 GB block at C4, and far worse when several processes of one box do it at
// once), while sampling() and every step of a benchmark ask for the same few sizes again and again.  The cache is
// bounded (6 GB and 16 blocks per device) and ages out (a block nobody asked for during 512 allocations is freed);
// cyp_release_cached_memory() empties it and trims the pool.
namespace {
struct BigBlock { void *p; size_t bytes; int dev; uint64_t stamp; };
std::mutex g_big_mutex;
std::vector<BigBlock> g_big_free;
std::unordered_map<void *, BigBlock> g_big_live;
uint64_t g_alloc_clock = 0;                          // dev_alloc calls so far
constexpr size_t kBigBlock = 64ull << 20;
constexpr size_t kBigCacheEntries = 16;
constexpr size_t kBigCacheBytes = 6ull << 30;        // per device: what a notebook that shares the GPU may lose to the cache
constexpr uint64_t kBigCacheAge = 512;               // a cached block nobody asked for during this many allocations goes back

void big_free_now(const BigBlock &b) {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != b.dev) cudaSetDevice(b.dev);
    cudaFreeAsync(b.p, nullptr);
    if (cur != b.dev && cur >= 0) cudaSetDevice(cur);
}

void big_cache_flush(int dev) {            // dev < 0: all devices.  Caller holds the mutex.
    std::vector<BigBlock> keep;
    for (const BigBlock &b : g_big_free) {
        if (dev >= 0 && b.dev != dev) { keep.push_back(b); continue; }
        big_free_now(b);
    }
    g_big_free.swap(keep);
}

// the cap and the age-out.  Caller holds the mutex.
void big_cache_trim(int dev) {
    std::vector<BigBlock> keep;
    size_t bytes = 0, entries = 0;
    for (auto it = g_big_free.rbegin(); it != g_big_free.rend(); ++it) {          // newest first
        const bool mine = it->dev == dev;
        const bool stale = g_alloc_clock - it->stamp > kBigCacheAge;
        if (mine && (stale || bytes + it->bytes > kBigCacheBytes || entries + 1 > kBigCacheEntries)) { big_free_now(*it); continue; }
        if (mine) { bytes += it->bytes; ++entries; }
        keep.push_back(*it);
    }
    g_big_free.assign(keep.rbegin(), keep.rend());
}
}  // namespace

size_t dev_cached_bytes(int dev) {
    std::lock_guard<std::mutex> lock(g_big_mutex);
    size_t total = 0;
    for (const BigBlock &b : g_big_free) if (b.dev == dev) total += b.bytes;
    return total;
}

void dev_release_cached(int dev) {
    std::lock_guard<std::mutex> lock(g_big_mutex);
    big_cache_flush(dev);
}

cudaError_t dev_alloc(void **p, size_t bytes) {
    static std::atomic<bool> pool_ready[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !pool_ready[dev].load()) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready[dev].store(true);
    }
    *p = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        ++g_alloc_clock;
        if (bytes >= kBigBlock) {
            for (size_t i = 0; i < g_big_free.size(); ++i) {
                if (g_big_free[i].dev == dev && g_big_free[i].bytes == bytes) {
                    *p = g_big_free[i].p;
                    g_big_live[*p] = g_big_free[i];
                    g_big_free.erase(g_big_free.begin() + i);
                    return cudaSuccess;
                }
            }
        }
        if ((g_alloc_clock & 63) == 0 && !g_big_free.empty()) big_cache_trim(dev);
    }
    e = cudaMallocAsync(p, bytes ? bytes : 1, nullptr);
    if (e == cudaErrorMemoryAllocation) {          // give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lock(g_big_mutex);
            big_cache_flush(dev);
        }
        cudaStreamSynchronize(nullptr);
        e = cudaMallocAsync(p, bytes ? bytes : 1, nullptr);
    }
    if (e == cudaSuccess && bytes >= kBigBlock) {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        g_big_live[*p] = BigBlock{*p, bytes, dev, g_alloc_clock};
    }
    return e;
}
void dev_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        auto it = g_big_live.find(p);
        if (it != g_big_live.end()) {
            BigBlock b = it->second;
            g_big_live.erase(it);
            b.stamp = g_alloc_clock;
            g_big_free.push_back(b);              // reusable in default-stream order, like cudaFreeAsync(p, 0)
            big_cache_trim(b.dev);                // at most kBigCacheBytes / kBigCacheEntries per device stay
            return;
        }
    }
    cudaFreeAsync(p, nullptr);
}

// One CTA handles kK1Threads consecutive rows.  The CTA's slice of the CSR value array is
// contiguous in HBM, so it is staged through shared memory with coalesced loads; each
// thread then walks its own row sequentially in shared memory (the serial float64 chain,
// overwriting the values with their running sums); finally the CTA's slice of the padded
// output, again contiguous, is written with coalesced stores.  Slices that do not fit the
// staging buffer (very long rows) take the direct per-thread path.
constexpr int kK1Threads = 128;

template <typename CdfT>
__device__ __forceinline__ CdfT finish_cdf(double acc) {
    if constexpr (sizeof(CdfT) == 8) {
        return acc;
    } else {
        const float c = __double2float_rn(acc);                               // :23/:28 float32 store
        return __fdiv_rn(rintf(__fmul_rn(c, 1000.0f)), 1000.0f);             // :31 np.round(decimals=3)
    }
}

// 16-bit code of a CDF value (what walk_tma's q16 flavour searches):
//   exact mode      q = min(65535, floor(c * 65536)); c is then known to lie in [q, q+1) / 65536
//                   (or above 65535/65536), the kernel resolves the rare ambiguous compare in float64;
//   reference mode  q = m, where the stored float32 value is exactly fl32(m / 1000) (:31): lossless.
// Values outside the representable range clear bit 1 of the flag word (q16 unusable).
template <typename CdfT>
__device__ __forceinline__ uint16_t cdf_code(double acc, int *flags) {
    if constexpr (sizeof(CdfT) == 8) {
        if (!(acc >= 0.0)) { atomicOr(flags, 2); return 0; }
        const double x = floor(acc * 65536.0);
        return static_cast<uint16_t>(x > 65535.0 ? 65535.0 : x);
    } else {
        const float m = rintf(__fmul_rn(__double2float_rn(acc), 1000.0f));
        if (!(m >= 0.0f && m <= 65535.0f)) { atomicOr(flags, 2); return 0; }
        return static_cast<uint16_t>(m);
    }
}

template <typename CdfT, typename ValT>
__global__ void __launch_bounds__(kK1Threads)
build_cdf_kernel(int64_t row_begin, int64_t n, const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ raw_col,
                 const ValT *__restrict__ raw_val, const RowInfo *__restrict__ rows, int align_elems,
                 int64_t total_units, CdfT *__restrict__ cdf, uint16_t *__restrict__ q16, int32_t *__restrict__ col,
                 EdgeRec *__restrict__ edge, double *__restrict__ val, int smem_elems, int *__restrict__ negative_flag,
                 double *__restrict__ block_sum_min, double *__restrict__ block_sum_max) {
    extern __shared__ double s_val[];                 // [smem_elems] raw values, then cumulative sums in place
    __shared__ int64_t s_src[kK1Threads + 1];         // CSR offsets of the CTA's rows
    __shared__ int64_t s_dst[kK1Threads + 1];         // padded offsets of the CTA's rows
    const int64_t r0 = row_begin + static_cast<int64_t>(blockIdx.x) * kK1Threads;     // row_begin is a multiple of kK1Threads
    const int64_t r1 = min(r0 + kK1Threads, n);
    const int64_t block_id = r0 / kK1Threads;
    const int nrows = static_cast<int>(r1 - r0);
    for (int i = threadIdx.x; i <= nrows; i += kK1Threads) {
        const int64_t r = r0 + i;
        s_src[i] = row_ptr[r];
        s_dst[i] = (r < n) ? static_cast<int64_t>(rows[r].off_units) * align_elems : total_units * align_elems;
    }
    __syncthreads();
    const int64_t src0 = s_src[0], cnt = s_src[nrows] - src0;
    const int64_t dst0 = s_dst[0], dcnt = s_dst[nrows] - dst0;
    double row_sum = 0.0;                              // this thread's row: last CDF value before narrowing
    if (cnt <= smem_elems) {
        bool neg = false, zero = false;
        for (int64_t i = threadIdx.x; i < cnt; i += kK1Threads) {
            const double v = static_cast<double>(raw_val[src0 + i]);      // float32 input widens exactly
            neg |= v < 0.0;
            zero |= v == 0.0;
            s_val[i] = v;
        }
        if (neg) atomicOr(negative_flag, 1);
        if (zero) atomicOr(negative_flag, 16);                            // an explicit zero: the CSR is not canonical
        __syncthreads();
        if (threadIdx.x < nrows) {
            const int64_t a = s_src[threadIdx.x] - src0, b = s_src[threadIdx.x + 1] - src0;
            double acc = 0.0;
            for (int64_t j = a; j < b; ++j) {
                acc = __dadd_rn(acc, s_val[j]);        // sequential: no re-association, no FMA
                s_val[j] = acc;
            }
            row_sum = acc;
        }
        __syncthreads();
        for (int64_t p = threadIdx.x; p < dcnt; p += kK1Threads) {
            int lo = 0, hi = nrows;                    // last row with s_dst[row] <= dst0 + p
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_dst[mid] - dst0 <= p) lo = mid; else hi = mid;
            }
            const int64_t j = p - (s_dst[lo] - dst0);
            const int64_t len = s_src[lo + 1] - s_src[lo];
            if (j < len) {
                const int64_t q = s_src[lo] + j;
                const int32_t c = raw_col[q];
                if (c < 0 || c >= n) atomicOr(negative_flag, 4);          // not a cell: the create call fails
                if (j > 0 && raw_col[q - 1] >= c) atomicOr(negative_flag, 8);      // columns not strictly increasing: not canonical
                cdf[dst0 + p] = finish_cdf<CdfT>(s_val[q - src0]);
                const uint16_t code = cdf_code<CdfT>(s_val[q - src0], negative_flag);     // also checks the code range
                col[dst0 + p] = c;
                if (q16) {
                    const RowInfo tr = rows[c];
                    q16[dst0 + p] = code;
                    edge[dst0 + p] = EdgeRec{c, tr.off_units, tr.len, 0u};
                }
                val[dst0 + p] = static_cast<double>(raw_val[q]);
            } else {                                   // padding never matches: u >= 0 > -1
                cdf[dst0 + p] = static_cast<CdfT>(-1);
                col[dst0 + p] = -1;
                if (q16) {
                    q16[dst0 + p] = 0;
                    edge[dst0 + p] = EdgeRec{-1, 0u, 0u, 0u};
                }
                val[dst0 + p] = 0.0;
            }
        }
    } else if (threadIdx.x < nrows) {
        const int64_t a = s_src[threadIdx.x], len = s_src[threadIdx.x + 1] - a;
        const int64_t dst = s_dst[threadIdx.x];
        const int64_t len_pad = (len + align_elems - 1) / align_elems * align_elems;
        double acc = 0.0;
        for (int64_t j = 0; j < len; ++j) {
            const double v = static_cast<double>(raw_val[a + j]);
            if (v < 0.0) atomicOr(negative_flag, 1);
            if (v == 0.0) atomicOr(negative_flag, 16);
            acc = __dadd_rn(acc, v);
            const int32_t c = raw_col[a + j];
            if (c < 0 || c >= n) atomicOr(negative_flag, 4);
            if (j > 0 && raw_col[a + j - 1] >= c) atomicOr(negative_flag, 8);
            cdf[dst + j] = finish_cdf<CdfT>(acc);
            const uint16_t code = cdf_code<CdfT>(acc, negative_flag);
            col[dst + j] = c;
            if (q16) {
                const RowInfo tr = rows[c];
                q16[dst + j] = code;
                edge[dst + j] = EdgeRec{c, tr.off_units, tr.len, 0u};
            }
            val[dst + j] = v;
        }
        row_sum = acc;
        for (int64_t j = len; j < len_pad; ++j) {
            cdf[dst + j] = static_cast<CdfT>(-1);
            col[dst + j] = -1;
            if (q16) {
                q16[dst + j] = 0;
                edge[dst + j] = EdgeRec{-1, 0u, 0u, 0u};
            }
            val[dst + j] = 0.0;
        }
    }
    // range of the row sums of this CTA (the normalisation check of markov_chain_sampler.py:182-187 runs on them)
    __shared__ double s_min[kK1Threads / 32], s_max[kK1Threads / 32];
    __syncthreads();
    double lo = (threadIdx.x < nrows) ? row_sum : DBL_MAX, hi = (threadIdx.x < nrows) ? row_sum : -DBL_MAX;
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kK1Threads / 32; ++w) { lo = fmin(lo, s_min[w]); hi = fmax(hi, s_max[w]); }
        block_sum_min[block_id] = lo;
        block_sum_max[block_id] = hi;
    }
}

template <typename CdfT>
__global__ void unpad_cdf_kernel(int64_t n, const int64_t *__restrict__ row_ptr, const RowInfo *__restrict__ rows,
                                 int align_elems, const CdfT *__restrict__ cdf, CdfT *__restrict__ out) {
    // one warp per row
    const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const int64_t a = row_ptr[r], len = row_ptr[r + 1] - a;
    const int64_t src = static_cast<int64_t>(rows[r].off_units) * align_elems;
    for (int64_t j = lane; j < len; j += 32) out[a + j] = cdf[src + j];
}

// 16-bit CDF codes + edge records (what walk.cu / walk_tma.cu / walk_pp.cu read), derived from K1's output.
template <typename CdfT>
__global__ void __launch_bounds__(256)
build_edge_records_kernel(int64_t padded, const RowInfo *__restrict__ rows, const CdfT *__restrict__ cdf, const int32_t *__restrict__ col,
                          uint16_t *__restrict__ q16, EdgeRec *__restrict__ edge) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= padded) return;
    const int32_t c = col[p];
    if (c >= 0) {
        int unused = 0;
        const RowInfo tr = rows[c];
        if constexpr (sizeof(CdfT) == 8) {
            q16[p] = cdf_code<double>(cdf[p], &unused);
        } else {
            const float m = rintf(__fmul_rn(cdf[p], 1000.0f));          // the stored value is fl32(m / 1000)
            q16[p] = (m >= 0.0f && m <= 65535.0f) ? static_cast<uint16_t>(m) : 0;
        }
        edge[p] = EdgeRec{c, tr.off_units, tr.len, 0u};
    } else {
        q16[p] = 0;
        edge[p] = EdgeRec{-1, 0u, 0u, 0u};
    }
}

int ensure_edge_records(const cyp_graph *cg) {
    cyp_graph *g = const_cast<cyp_graph *>(cg);
    if (g->d_edge) return CYP_OK;
    DeviceGuard guard(g->device);
    uint16_t *q16 = nullptr;
    EdgeRec *edge = nullptr;
    cudaError_t e = dev_alloc(&q16, sizeof(uint16_t) * g->padded);
    if (e == cudaSuccess) e = dev_alloc(&edge, sizeof(EdgeRec) * g->padded);
    if (e == cudaSuccess) {
        const unsigned grid = static_cast<unsigned>((g->padded + 255) / 256);
        if (g->cdf_mode == CYP_CDF_EXACT)
            build_edge_records_kernel<double><<<grid, 256>>>(g->padded, g->d_row, static_cast<const double *>(g->d_cdf), g->d_col, q16, edge);
        else
            build_edge_records_kernel<float><<<grid, 256>>>(g->padded, g->d_row, static_cast<const float *>(g->d_cdf), g->d_col, q16, edge);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    }
    if (e != cudaSuccess) {
        dev_free(q16);
        dev_free(edge);
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("ensure_edge_records: ") + cudaGetErrorString(e));
    }
    g->d_q16 = q16;
    g->d_edge = edge;
    g->device_bytes += static_cast<int64_t>((sizeof(uint16_t) + sizeof(EdgeRec)) * g->padded);
    return CYP_OK;
}

GraphWire graph_wire(const cyp_graph *g) {
    GraphWire w{};
    w.cdf_mode = g->cdf_mode; w.max_len = g->max_len; w.q16_ok = g->q16_ok; w.monotone = g->monotone;
    w.v2_ok = g->v2_ok; w.v2_max_sectors = g->v2_max_sectors; w.v2_slot_sectors = g->v2_slot_sectors; w.v2_uniform = g->v2_uniform;
    w.has_edge_records = g->d_edge != nullptr ? 1 : 0;
    w.noncanonical = g->noncanonical;
    w.n = g->n; w.nnz = g->nnz; w.padded = g->padded; w.v2_sectors = g->v2_sectors; w.v2_edge_units = g->v2_edge_units;
    w.rows2_bytes = g->rows2_bytes;
    w.mean_len = g->mean_len; w.row_sum_min = g->row_sum_min; w.row_sum_max = g->row_sum_max;
    w.build_ms = g->build_ms;
    return w;
}

void graph_device_arrays(const cyp_graph *g, std::vector<std::pair<void *, size_t>> *out) {
    const size_t cdf_elem = (g->cdf_mode == CYP_CDF_EXACT) ? 8 : 4;
    const size_t padded = static_cast<size_t>(g->padded), n = static_cast<size_t>(g->n);
    out->clear();
    out->emplace_back(g->d_row, sizeof(RowInfo) * n);
    out->emplace_back(g->d_cdf, cdf_elem * padded);
    out->emplace_back(g->d_col, sizeof(int32_t) * padded);
    out->emplace_back(g->d_val, sizeof(double) * padded);
    out->emplace_back(g->d_row_ptr, sizeof(int64_t) * (n + 1));
    if (g->v2_ok) {
        out->emplace_back(g->d_rows2, static_cast<size_t>(g->rows2_bytes));
        out->emplace_back(g->d_edge4, sizeof(uint32_t) * 4 * static_cast<size_t>(g->v2_edge_units + 1));
        out->emplace_back(g->d_word, sizeof(uint32_t) * n);
    }
    if (g->d_edge) {
        out->emplace_back(g->d_q16, sizeof(uint16_t) * padded);
        out->emplace_back(g->d_edge, sizeof(EdgeRec) * padded);
    }
}

static void graph_free(cyp_graph *g);

int graph_alloc_like(const GraphWire &w, int device, cyp_graph **out) {
    *out = nullptr;
    DeviceGuard guard(device);
    cyp_graph *g = new cyp_graph();
    g->device = device;
    g->cdf_mode = w.cdf_mode; g->n = w.n; g->nnz = w.nnz; g->padded = w.padded; g->max_len = w.max_len; g->mean_len = w.mean_len;
    g->q16_ok = w.q16_ok; g->monotone = w.monotone; g->build_ms = w.build_ms; g->noncanonical = w.noncanonical;
    g->row_sum_min = w.row_sum_min; g->row_sum_max = w.row_sum_max;
    g->v2_ok = w.v2_ok; g->v2_max_sectors = w.v2_max_sectors; g->v2_slot_sectors = w.v2_slot_sectors; g->v2_uniform = w.v2_uniform;
    g->v2_sectors = w.v2_sectors; g->v2_edge_units = w.v2_edge_units; g->rows2_bytes = w.rows2_bytes;
    const size_t cdf_elem = (w.cdf_mode == CYP_CDF_EXACT) ? 8 : 4;
    const size_t padded = static_cast<size_t>(w.padded), n = static_cast<size_t>(w.n);
    cudaError_t e = dev_alloc(&g->d_row, sizeof(RowInfo) * n);
    if (e == cudaSuccess) e = dev_alloc(&g->d_cdf, cdf_elem * padded);
    if (e == cudaSuccess) e = dev_alloc(&g->d_col, sizeof(int32_t) * padded);
    if (e == cudaSuccess) e = dev_alloc(&g->d_val, sizeof(double) * padded);
    if (e == cudaSuccess) e = dev_alloc(&g->d_row_ptr, sizeof(int64_t) * (n + 1));
    g->device_bytes = static_cast<int64_t>(sizeof(RowInfo) * n + sizeof(int64_t) * (n + 1) + (cdf_elem + sizeof(int32_t) + sizeof(double)) * padded);
    if (e == cudaSuccess && w.v2_ok) {
        e = dev_alloc(&g->d_rows2_base, static_cast<size_t>(w.rows2_bytes));
        g->d_rows2 = g->d_rows2_base;
        if (e == cudaSuccess) e = dev_alloc(&g->d_edge4, sizeof(uint32_t) * 4 * static_cast<size_t>(w.v2_edge_units + 1));
        if (e == cudaSuccess) e = dev_alloc(&g->d_word, sizeof(uint32_t) * n);
        g->device_bytes += static_cast<int64_t>(w.rows2_bytes + sizeof(uint32_t) * 4 * (w.v2_edge_units + 1) + sizeof(uint32_t) * n);
    }
    if (e == cudaSuccess && w.has_edge_records) {
        e = dev_alloc(&g->d_q16, sizeof(uint16_t) * padded);
        if (e == cudaSuccess) e = dev_alloc(&g->d_edge, sizeof(EdgeRec) * padded);
        g->device_bytes += static_cast<int64_t>((sizeof(uint16_t) + sizeof(EdgeRec)) * padded);
    }
    if (e != cudaSuccess) {
        graph_free(g);
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("graph_alloc_like: ") + cudaGetErrorString(e));
    }
    *out = g;
    return CYP_OK;
}

}  // namespace cyp

using namespace cyp;

extern "C" int cyp_version(void) { return CYP_VERSION; }
extern "C" const char *cyp_last_error(void) { return g_last_error.c_str(); }

extern "C" int cyp_release_cached_memory(int device) {
    dev_release_cached(device);
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    for (int d = 0; d < ndev; ++d) {
        if (device >= 0 && d != device) continue;
        DeviceGuard guard(d);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
    }
    return CYP_OK;
}

extern "C" int cyp_device_count(int *n) {
    CYP_REQUIRE(n != nullptr, "cyp_device_count: n is NULL");
    CYP_CUDA(cudaGetDeviceCount(n));
    return CYP_OK;
}

namespace cyp {
static void graph_free(cyp_graph *g) {
    if (!g) return;
    DeviceGuard guard(g->device);
    dev_free(g->d_row);
    dev_free(g->d_cdf);
    dev_free(g->d_q16);
    dev_free(g->d_col);
    dev_free(g->d_edge);
    dev_free(g->d_val);
    dev_free(g->d_row_ptr);
    dev_free(g->d_rows2_base);
    dev_free(g->d_edge4);
    dev_free(g->d_word);
    delete g;
}
}  // namespace cyp

// cyp_graph_create: upload + K0 + K1 + K1b as a pipeline.  The CSR arrays travel in chunks of rows on a copy stream, the
// row offsets first; K0 (layout.cu) lays out the rows on the device as soon as the offsets have landed; K1 and K1b of a
// chunk start as soon as the chunk has, so at atlas scale (628 MB of CSR over PCIe) the kernels hide behind the upload.
static int graph_create_impl(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, const int32_t *h_col,
                             const void *h_val, bool val_f32, int cdf_mode, const RawCsr *raw, cyp_graph **out) {
    struct RawGuard {                                            // `raw` is ours whatever happens: freed here until the build adopts it
        const RawCsr *raw; int device; bool adopted;
        ~RawGuard() { if (raw && !adopted) { DeviceGuard guard(device); dev_free(raw->d_col); dev_free(raw->d_val); } }
    } raw_guard{raw, device, false};
    CYP_REQUIRE(out != nullptr, "cyp_graph_create: out is NULL");
    *out = nullptr;
    CYP_REQUIRE(n_cells > 0 && nnz > 0, "cyp_graph_create: empty matrix");
    CYP_REQUIRE(n_cells < (1ll << 31), "cyp_graph_create: cell ids must fit int32 (the reference stores int32, :38)");
    CYP_REQUIRE(h_row_ptr && (raw ? (raw->d_col && raw->d_val) : (h_col && h_val)), "cyp_graph_create: NULL CSR array");
    CYP_REQUIRE(cdf_mode == CYP_CDF_REFERENCE || cdf_mode == CYP_CDF_EXACT, "cyp_graph_create: bad cdf_mode");
    CYP_REQUIRE(h_row_ptr[0] == 0 && h_row_ptr[n_cells] == nnz, "cyp_graph_create: row_ptr does not span nnz");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_graph_create: no such CUDA device");
    DeviceGuard guard(device);

    cyp_graph *g = new cyp_graph();
    g->device = device;
    g->cdf_mode = cdf_mode;
    g->n = n_cells;
    g->nnz = nnz;
    g->align_elems = kRowAlign;
    const int A = g->align_elems;
    const size_t cdf_elem = (cdf_mode == CYP_CDF_EXACT) ? 8 : 4;
    const size_t val_elem = val_f32 ? sizeof(float) : sizeof(double);

    constexpr int kMaxChunks = 8;
    int32_t *d_raw_col = nullptr;
    void *d_raw_val = nullptr;
    double *d_block_min = nullptr, *d_block_max = nullptr;
    int *d_flag = nullptr;
    cudaStream_t s_copy = nullptr, s_aux = nullptr, s_comp = nullptr;
    cudaEvent_t ev_alloc = nullptr, ev_small = nullptr, ev_chunk[kMaxChunks] = {}, ev_t[kMaxChunks][2] = {};
    V2Build v2;
    RowLayoutScratch layout;
    auto release_tmp = [&]() {
        if (s_copy) cudaStreamSynchronize(s_copy);
        if (s_aux) cudaStreamSynchronize(s_aux);
        if (s_comp) cudaStreamSynchronize(s_comp);
        dev_free(d_raw_col); dev_free(d_raw_val); dev_free(d_flag); dev_free(d_block_min); dev_free(d_block_max);
        layout_release(&layout);
        d_raw_col = nullptr; d_raw_val = nullptr; d_flag = nullptr; d_block_min = d_block_max = nullptr;
        if (ev_alloc) cudaEventDestroy(ev_alloc);
        if (ev_small) cudaEventDestroy(ev_small);
        for (int c = 0; c < kMaxChunks; ++c) {
            if (ev_chunk[c]) cudaEventDestroy(ev_chunk[c]);
            if (ev_t[c][0]) cudaEventDestroy(ev_t[c][0]);
            if (ev_t[c][1]) cudaEventDestroy(ev_t[c][1]);
        }
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_aux) cudaStreamDestroy(s_aux);
        if (s_comp) cudaStreamDestroy(s_comp);
        ev_alloc = ev_small = nullptr; s_copy = s_aux = s_comp = nullptr;
    };
    auto cleanup = [&](int code, const std::string &msg) {
        release_tmp();
        v2_release(g, &v2);
        graph_free(g);
        return fail(code, msg);
    };
#define G_CUDA(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return cleanup(_e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, \
                                              std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

    // ---- 1. raw CSR on its way: the row offsets first, then chunks of rows (multiples of kK1Threads rows, so K1's CTAs
    // never straddle a chunk)
    const int n_chunks = nnz < (4 << 20) ? 1 : kMaxChunks;
    int64_t chunk_rows = (n_cells + n_chunks - 1) / n_chunks;
    chunk_rows = (chunk_rows + kK1Threads - 1) / kK1Threads * kK1Threads;
    G_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    G_CUDA(cudaStreamCreateWithFlags(&s_aux, cudaStreamNonBlocking));
    G_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    G_CUDA(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
    G_CUDA(cudaEventCreateWithFlags(&ev_small, cudaEventDisableTiming));
    G_CUDA(dev_alloc(&g->d_row_ptr, sizeof(int64_t) * (n_cells + 1)));
    if (raw) {                                                   // already on the device: ours from here on
        d_raw_col = raw->d_col;
        d_raw_val = raw->d_val;
        raw_guard.adopted = true;
    } else {
        G_CUDA(dev_alloc(&d_raw_col, sizeof(int32_t) * nnz));
        G_CUDA(dev_alloc(&d_raw_val, val_elem * nnz));
    }
    G_CUDA(cudaEventRecord(ev_alloc, nullptr));                  // device memory comes from the default stream's pool
    G_CUDA(cudaStreamWaitEvent(s_copy, ev_alloc, 0));
    G_CUDA(cudaMemcpyAsync(g->d_row_ptr, h_row_ptr, sizeof(int64_t) * (n_cells + 1), cudaMemcpyHostToDevice, s_copy));
    G_CUDA(cudaEventRecord(ev_small, s_copy));
    G_CUDA(cudaStreamWaitEvent(s_aux, ev_small, 0));             // K0 runs behind the row offsets, beside the chunks
    int used_chunks = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t rb = std::min<int64_t>(static_cast<int64_t>(c) * chunk_rows, n_cells), re = std::min<int64_t>(rb + chunk_rows, n_cells);
        if (rb >= re) break;
        const int64_t a = h_row_ptr[rb], b = h_row_ptr[re];
        if (!(a >= 0 && b >= a && b <= nnz)) return cleanup(CYP_E_INVALID, "cyp_graph_create: row_ptr not monotone");
        if (b > a && !raw) {
            G_CUDA(cudaMemcpyAsync(d_raw_col + a, h_col + a, sizeof(int32_t) * (b - a), cudaMemcpyHostToDevice, s_copy));
            G_CUDA(cudaMemcpyAsync(static_cast<char *>(d_raw_val) + val_elem * a, static_cast<const char *>(h_val) + val_elem * a, val_elem * (b - a),
                                   cudaMemcpyHostToDevice, s_copy));
        }
        G_CUDA(cudaEventCreateWithFlags(&ev_chunk[c], cudaEventDisableTiming));
        G_CUDA(cudaEventRecord(ev_chunk[c], s_copy));             // (with `raw`: the row offsets only)
        used_chunks = c + 1;
    }
    // ---- 2. K0, while the copies run: the row layout on the device (layout.cu); the host reads the totals
    int extra[2];
    v2_candidates(&extra[0], &extra[1]);
    RowLayoutSummary shape{};
    G_CUDA(layout_measure(n_cells, A, g->d_row_ptr, extra[0], extra[1], &layout, s_aux, &shape));
    if (shape.bad) return cleanup(CYP_E_INVALID, "cyp_graph_create: row_ptr not monotone");
    const int64_t units = shape.units16;
    if (units >= (1ll << 32)) return cleanup(CYP_E_INVALID, "cyp_graph_create: matrix too large for 32-bit row units");
    g->padded = units * A + 4 * A;                     // slack so vector loads of a last partial chunk stay in bounds
    g->max_len = shape.max_len;
    g->mean_len = static_cast<double>(nnz) / static_cast<double>(n_cells);
    const bool want_v2 = v2_plan(g, shape, extra, &v2);

    // ---- 3. the other device arrays (default stream)
    G_CUDA(dev_alloc(&g->d_row, sizeof(RowInfo) * n_cells));
    G_CUDA(dev_alloc(&g->d_cdf, cdf_elem * g->padded));
    G_CUDA(dev_alloc(&g->d_col, sizeof(int32_t) * g->padded));
    G_CUDA(dev_alloc(&g->d_val, sizeof(double) * g->padded));
    G_CUDA(dev_alloc(&d_flag, sizeof(int)));
    const unsigned total_blocks = static_cast<unsigned>((n_cells + kK1Threads - 1) / kK1Threads);
    G_CUDA(dev_alloc(&d_block_min, sizeof(double) * total_blocks));
    G_CUDA(dev_alloc(&d_block_max, sizeof(double) * total_blocks));
    G_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), nullptr));
    g->device_bytes = sizeof(RowInfo) * n_cells + sizeof(int64_t) * (n_cells + 1) + (cdf_elem + sizeof(int32_t) + sizeof(double)) * g->padded;
    // slack region: never-matching padding
    G_CUDA(cudaMemsetAsync(static_cast<char *>(g->d_cdf) + cdf_elem * units * A, 0xFF, cdf_elem * 4 * A, nullptr));   // NaN: u <= NaN is false
    G_CUDA(cudaMemsetAsync(g->d_col + units * A, 0xFF, sizeof(int32_t) * 4 * A, nullptr));
    G_CUDA(cudaMemsetAsync(g->d_val + units * A, 0, sizeof(double) * 4 * A, nullptr));
    if (want_v2) G_CUDA(v2_alloc(g, &v2));
    G_CUDA(cudaEventRecord(ev_alloc, nullptr));                  // device memory comes from the default stream's pool

    // RowInfo, row words and target bases of every row: written on the device from K0's prefix sums
    G_CUDA(cudaStreamWaitEvent(s_aux, ev_alloc, 0));
    G_CUDA(layout_write(&layout, g->d_row_ptr, v2.which, g->d_row, want_v2 ? g->d_word : nullptr, want_v2 ? v2.d_ebase : nullptr, s_aux));
    G_CUDA(cudaEventRecord(ev_small, s_aux));

    // ---- 4. K1 (+ K1b) per chunk
    // staging buffer: 1.5x a typical 128-row slice (so several CTAs fit an SM), at most 160 KB
    int smem_elems = static_cast<int>(1.5 * kK1Threads * g->mean_len) + 256;
    if (smem_elems < 2048) smem_elems = 2048;
    if (smem_elems > 20 * 1024) smem_elems = 20 * 1024;
    const int smem_bytes = smem_elems * static_cast<int>(sizeof(double));
    G_CUDA(cudaStreamWaitEvent(s_comp, ev_alloc, 0));
    G_CUDA(cudaStreamWaitEvent(s_comp, ev_small, 0));
    if (raw && raw->ready) G_CUDA(cudaStreamWaitEvent(s_comp, raw->ready, 0));
#define K1_LAUNCH(CDFT, VALT)                                                                                               \
    do {                                                                                                                    \
        G_CUDA(cudaFuncSetAttribute(build_cdf_kernel<CDFT, VALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
        build_cdf_kernel<CDFT, VALT><<<grid, kK1Threads, smem_bytes, s_comp>>>(                                             \
            rb, n_cells, g->d_row_ptr, d_raw_col, static_cast<const VALT *>(d_raw_val), g->d_row, A, units,                 \
            static_cast<CDFT *>(g->d_cdf), nullptr, g->d_col, nullptr, g->d_val, smem_elems, d_flag, d_block_min, d_block_max); \
    } while (0)
    for (int c = 0; c < used_chunks; ++c) {
        const int64_t rb = static_cast<int64_t>(c) * chunk_rows, re = std::min<int64_t>(rb + chunk_rows, n_cells);
        const unsigned grid = static_cast<unsigned>((re - rb + kK1Threads - 1) / kK1Threads);
        G_CUDA(cudaEventCreate(&ev_t[c][0]));
        G_CUDA(cudaEventCreate(&ev_t[c][1]));
        G_CUDA(cudaStreamWaitEvent(s_comp, ev_chunk[c], 0));
        G_CUDA(cudaEventRecord(ev_t[c][0], s_comp));
        if (cdf_mode == CYP_CDF_EXACT) {
            if (val_f32) K1_LAUNCH(double, float); else K1_LAUNCH(double, double);
        } else {
            if (val_f32) K1_LAUNCH(float, float); else K1_LAUNCH(float, double);
        }
        G_CUDA(cudaGetLastError());
        if (want_v2) G_CUDA(v2_launch(g, &v2, rb, re, s_comp));
        G_CUDA(cudaEventRecord(ev_t[c][1], s_comp));
    }
#undef K1_LAUNCH
    G_CUDA(cudaStreamSynchronize(s_comp));
    g->build_ms = 0.f;
    for (int c = 0; c < used_chunks; ++c) {
        float ms = 0.f;
        G_CUDA(cudaEventElapsedTime(&ms, ev_t[c][0], ev_t[c][1]));
        g->build_ms += ms;
    }

    // ---- 5. what K1 found out
    int negative = 0;
    G_CUDA(cudaMemcpy(&negative, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (negative & 4) return cleanup(CYP_E_INVALID, "cyp_graph_create: column index out of range [0, n_cells)");
    g->monotone = (negative & 1) ? 0 : 1;
    g->q16_ok = (negative & 3) == 0 ? 1 : 0;
    g->noncanonical = (negative >> 3) & 3;
    {
        std::vector<double> bmin(total_blocks), bmax(total_blocks);
        G_CUDA(cudaMemcpy(bmin.data(), d_block_min, sizeof(double) * total_blocks, cudaMemcpyDeviceToHost));
        G_CUDA(cudaMemcpy(bmax.data(), d_block_max, sizeof(double) * total_blocks, cudaMemcpyDeviceToHost));
        g->row_sum_min = bmin[0];
        g->row_sum_max = bmax[0];
        for (unsigned b = 1; b < total_blocks; ++b) {
            g->row_sum_min = std::min(g->row_sum_min, bmin[b]);
            g->row_sum_max = std::max(g->row_sum_max, bmax[b]);
        }
    }
    if (want_v2) G_CUDA(v2_finish(g, &v2, g->q16_ok && g->monotone));
    release_tmp();
#undef G_CUDA
    // the 16-bit codes and 16-byte edge records of the other step kernels: now if they are the default for this
    // graph, else on first use (ensure_edge_records)
    if (!g->v2_ok) {
        const int rc2 = ensure_edge_records(g);
        if (rc2 != CYP_OK) {
            graph_free(g);
            return rc2;
        }
    }
    *out = g;
    return CYP_OK;
}

extern "C" int cyp_graph_create(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                                const int32_t *h_col, const double *h_val, int cdf_mode, cyp_graph **out) {
    return graph_create_impl(device, n_cells, nnz, h_row_ptr, h_col, h_val, false, cdf_mode, nullptr, out);
}

extern "C" int cyp_graph_create_f32(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                                    const int32_t *h_col, const float *h_val, int cdf_mode, cyp_graph **out) {
    return graph_create_impl(device, n_cells, nnz, h_row_ptr, h_col, h_val, true, cdf_mode, nullptr, out);
}

int cyp::graph_create_raw(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, bool val_f32, int cdf_mode, const RawCsr &raw,
                          cyp_graph **out) {
    return graph_create_impl(device, n_cells, nnz, h_row_ptr, nullptr, nullptr, val_f32, cdf_mode, &raw, out);
}

extern "C" int cyp_graph_row_sum_range(const cyp_graph *g, double *min_sum, double *max_sum) {
    CYP_REQUIRE(g != nullptr && min_sum != nullptr && max_sum != nullptr, "cyp_graph_row_sum_range: NULL argument");
    *min_sum = g->row_sum_min;
    *max_sum = g->row_sum_max;
    return CYP_OK;
}

extern "C" int cyp_graph_destroy(cyp_graph *g) {
    graph_free(g);
    return CYP_OK;
}

extern "C" int cyp_graph_info(const cyp_graph *g, int64_t *n_cells, int64_t *nnz, int32_t *max_row_nnz,
                              int *cdf_mode, int *device, int64_t *device_bytes) {
    CYP_REQUIRE(g != nullptr, "cyp_graph_info: NULL graph");
    if (n_cells) *n_cells = g->n;
    if (nnz) *nnz = g->nnz;
    if (max_row_nnz) *max_row_nnz = g->max_len;
    if (cdf_mode) *cdf_mode = g->cdf_mode;
    if (device) *device = g->device;
    if (device_bytes) *device_bytes = g->device_bytes;
    return CYP_OK;
}

extern "C" int cyp_graph_canonical(const cyp_graph *g, int *columns_strictly_increasing, int *no_explicit_zeros) {
    CYP_REQUIRE(g != nullptr, "cyp_graph_canonical: NULL graph");
    if (columns_strictly_increasing) *columns_strictly_increasing = (g->noncanonical & 1) ? 0 : 1;
    if (no_explicit_zeros) *no_explicit_zeros = (g->noncanonical & 2) ? 0 : 1;
    return CYP_OK;
}

extern "C" int cyp_graph_layout(const cyp_graph *g, int *fused, int64_t *row_block_bytes, int64_t *target_word_bytes,
                                int32_t *max_block_bytes) {
    CYP_REQUIRE(g != nullptr, "cyp_graph_layout: NULL graph");
    if (fused) *fused = g->v2_ok;
    if (row_block_bytes) *row_block_bytes = g->v2_ok ? g->v2_sectors * 32 : 0;
    if (target_word_bytes) *target_word_bytes = g->v2_ok ? g->v2_edge_units * 16 : 0;
    if (max_block_bytes) *max_block_bytes = g->v2_ok ? g->v2_max_sectors * 32 : 0;
    return CYP_OK;
}

extern "C" int cyp_graph_fetch_fused(const cyp_graph *g, uint8_t *h_rows, uint32_t *h_words, uint32_t *h_targets) {
    CYP_REQUIRE(g != nullptr && h_rows && h_words && h_targets, "cyp_graph_fetch_fused: NULL argument");
    CYP_REQUIRE(g->v2_ok, "cyp_graph_fetch_fused: this graph has no fused rows");
    DeviceGuard guard(g->device);
    CYP_CUDA(cudaMemcpy(h_rows, g->d_rows2, static_cast<size_t>(g->v2_sectors) * 32, cudaMemcpyDeviceToHost));
    CYP_CUDA(cudaMemcpy(h_words, g->d_word, sizeof(uint32_t) * g->n, cudaMemcpyDeviceToHost));
    CYP_CUDA(cudaMemcpy(h_targets, g->d_edge4, sizeof(uint32_t) * 4 * static_cast<size_t>(g->v2_edge_units), cudaMemcpyDeviceToHost));
    return CYP_OK;
}

extern "C" int cyp_graph_placement(const cyp_graph *g, float *probe_ms_before, float *probe_ms_after, int *moves) {
    CYP_REQUIRE(g != nullptr, "cyp_graph_placement: NULL graph");
    if (probe_ms_before) *probe_ms_before = g->placement_ms_before;
    if (probe_ms_after) *probe_ms_after = g->placement_ms_after;
    if (moves) *moves = g->placement_moves;
    return CYP_OK;
}

extern "C" int cyp_graph_build_ms(const cyp_graph *g, float *ms) {
    CYP_REQUIRE(g != nullptr && ms != nullptr, "cyp_graph_build_ms: NULL argument");
    *ms = g->build_ms;
    return CYP_OK;
}

extern "C" int cyp_graph_fetch_cdf(const cyp_graph *g, void *h_cdf) {
    CYP_REQUIRE(g != nullptr && h_cdf != nullptr, "cyp_graph_fetch_cdf: NULL argument");
    DeviceGuard guard(g->device);
    const size_t elem = (g->cdf_mode == CYP_CDF_EXACT) ? 8 : 4;
    void *d_out = nullptr;
    CYP_CUDA(dev_alloc(&d_out, elem * g->nnz));
    const int threads = 256;
    const unsigned grid = static_cast<unsigned>((g->n * 32 + threads - 1) / threads);
    if (g->cdf_mode == CYP_CDF_EXACT)
        unpad_cdf_kernel<double><<<grid, threads>>>(g->n, g->d_row_ptr, g->d_row, g->align_elems,
                                                   static_cast<const double *>(g->d_cdf), static_cast<double *>(d_out));
    else
        unpad_cdf_kernel<float><<<grid, threads>>>(g->n, g->d_row_ptr, g->d_row, g->align_elems,
                                                  static_cast<const float *>(g->d_cdf), static_cast<float *>(d_out));
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(h_cdf, d_out, elem * g->nnz, cudaMemcpyDeviceToHost);
    dev_free(d_out);
    if (e != cudaSuccess) return fail(CYP_E_CUDA, std::string("cyp_graph_fetch_cdf: ") + cudaGetErrorString(e));
    return CYP_OK;
}
