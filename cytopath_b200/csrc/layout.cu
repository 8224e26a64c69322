// K0: the row layout of a graph, computed on the device from the uploaded CSR row offsets.
//
// matrix_prep (cytopath/markov_chain_sampler.py:11-31) pads every row to the longest one; here every row starts at a
// multiple of 16 elements in the padded arrays (RowInfo) and owns a block of 32-byte sectors in the fused rows plus a
// run of 4-word units in the target array (row word, edge base: common.cuh).  All three are prefix sums over the row
// lengths.  The host used to walk the n+1 offsets three times and upload 16 bytes per row from pageable memory: ~9 ms
// at 1M cells, hidden behind the 12 ms CSR upload of a single GPU but not behind the 1.5 ms slice a rank of an 8-GPU
// group uploads.  On the device it is three kernels and four cub scans, ~0.1 ms, and the host only reads the totals.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace cyp {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
row_shape_kernel(int64_t n, const int64_t *__restrict__ row_ptr, int align_elems, int extra0, int extra1, int64_t *__restrict__ units16,
                 int64_t *__restrict__ sect0, int64_t *__restrict__ sect1, int64_t *__restrict__ units4, RowLayoutSummary *__restrict__ sum) {
    __shared__ unsigned int s_hist[2][kV2MaxSectors + 1];
    __shared__ int s_max[3];
    __shared__ int s_bad;
    for (int i = threadIdx.x; i < 2 * (kV2MaxSectors + 1); i += kThreads) (&s_hist[0][0])[i] = 0u;
    if (threadIdx.x < 3) s_max[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int64_t r = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (r < n) {
        const int64_t len64 = row_ptr[r + 1] - row_ptr[r];
        const bool bad = len64 < 0 || len64 > (1ll << 30);
        const int len = bad ? 0 : static_cast<int>(len64);
        const int n0 = v2_block_sectors(len, extra0), n1 = v2_block_sectors(len, extra1);
        units16[r] = (len + align_elems - 1) / align_elems;
        sect0[r] = n0;
        sect1[r] = n1;
        units4[r] = (len + 3) / 4;
        atomicAdd(&s_hist[0][n0 < kV2MaxSectors ? n0 : kV2MaxSectors], 1u);
        atomicAdd(&s_hist[1][n1 < kV2MaxSectors ? n1 : kV2MaxSectors], 1u);
        atomicMax(&s_max[0], len);
        atomicMax(&s_max[1], n0);
        atomicMax(&s_max[2], n1);
        if (bad) s_bad = 1;
    } else if (r == n) {                                         // the scans run over n + 1 entries: entry n receives the total
        units16[r] = 0; sect0[r] = 0; sect1[r] = 0; units4[r] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (kV2MaxSectors + 1); i += kThreads) {
        const unsigned int c = (&s_hist[0][0])[i];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(&sum->hist[0][0]) + i, static_cast<unsigned long long>(c));
    }
    if (threadIdx.x == 0) {
        atomicMax(&sum->max_len, s_max[0]);
        atomicMax(&sum->max_sectors[0], s_max[1]);
        atomicMax(&sum->max_sectors[1], s_max[2]);
        if (s_bad) atomicOr(&sum->bad, 1);
    }
}

__global__ void row_totals_kernel(int64_t n, const int64_t *__restrict__ units16, const int64_t *__restrict__ sect0, const int64_t *__restrict__ sect1,
                                  const int64_t *__restrict__ units4, RowLayoutSummary *__restrict__ sum) {
    sum->units16 = units16[n];
    sum->sectors[0] = sect0[n];
    sum->sectors[1] = sect1[n];
    sum->units4 = units4[n];
}

__global__ void __launch_bounds__(kThreads)
row_write_kernel(int64_t n, const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ units16, const int64_t *__restrict__ sect,
                 const int64_t *__restrict__ units4, int extra, RowInfo *__restrict__ rows, uint32_t *__restrict__ word, uint32_t *__restrict__ ebase) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (r >= n) return;
    const int len = static_cast<int>(row_ptr[r + 1] - row_ptr[r]);
    RowInfo ri;
    ri.off_units = static_cast<uint32_t>(units16[r]);
    ri.len = static_cast<uint32_t>(len);
    rows[r] = ri;
    if (word) {
        // the word's sector field saturates: a longer block is never copied whole (slots hold <= 8 sectors), the step
        // kernel searches it in HBM and takes its length from the header
        const int nsect = v2_block_sectors(len, extra);
        word[r] = (static_cast<uint32_t>(nsect < kV2MaxSectors ? nsect : kV2MaxSectors) << kWordOffBits) | static_cast<uint32_t>(sect[r]);
        ebase[r] = static_cast<uint32_t>(units4[r]);
    }
}

}  // namespace

void layout_release(RowLayoutScratch *s) {
    dev_free(s->d_scan);
    dev_free(s->d_tmp);
    dev_free(s->d_sum);
    *s = RowLayoutScratch{};
}

// Row shapes + prefix sums on `stream` (d_row_ptr must be complete there); the summary comes back to the host and the
// stream is synchronised.  extra0 / extra1: the two candidate numbers of extra inline sectors per row (rows_v2.cu).
cudaError_t layout_measure(int64_t n, int align_elems, const int64_t *d_row_ptr, int extra0, int extra1, RowLayoutScratch *s,
                           cudaStream_t stream, RowLayoutSummary *h_out) {
    const size_t m = static_cast<size_t>(n) + 1;
    s->n = n;
    s->extra[0] = extra0;
    s->extra[1] = extra1;
    cudaError_t e = dev_alloc(&s->d_scan, sizeof(int64_t) * 4 * m);
    if (e == cudaSuccess) e = dev_alloc(&s->d_sum, sizeof(RowLayoutSummary));
    if (e != cudaSuccess) return e;
    int64_t *u16 = s->d_scan, *s0 = u16 + m, *s1 = s0 + m, *u4 = s1 + m;
    size_t tmp = 0;
    e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, u16, u16, static_cast<int>(m), stream);
    if (e != cudaSuccess) return e;
    s->tmp_bytes = tmp;
    e = dev_alloc(&s->d_tmp, tmp);
    if (e != cudaSuccess) return e;
    // the three allocations come from the default stream's pool: order `stream` behind them
    cudaEvent_t ev = nullptr;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(ev, nullptr);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(s->d_sum, 0, sizeof(RowLayoutSummary), stream);
    if (e != cudaSuccess) return e;
    const unsigned grid = static_cast<unsigned>((m + kThreads - 1) / kThreads);
    row_shape_kernel<<<grid, kThreads, 0, stream>>>(n, d_row_ptr, align_elems, extra0, extra1, u16, s0, s1, u4, s->d_sum);
    for (int64_t *a : {u16, s0, s1, u4}) {
        e = cub::DeviceScan::ExclusiveSum(s->d_tmp, tmp, a, a, static_cast<int>(m), stream);
        if (e != cudaSuccess) return e;
    }
    row_totals_kernel<<<1, 1, 0, stream>>>(n, u16, s0, s1, u4, s->d_sum);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, s->d_sum, sizeof(RowLayoutSummary), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    return e;
}

// RowInfo of every row and, when d_word is given, the row words / target bases for candidate `which` of layout_measure.
cudaError_t layout_write(const RowLayoutScratch *s, const int64_t *d_row_ptr, int which, RowInfo *d_row, uint32_t *d_word, uint32_t *d_ebase,
                         cudaStream_t stream) {
    const size_t m = static_cast<size_t>(s->n) + 1;
    const int64_t *u16 = s->d_scan, *sect = u16 + m * (which ? 2 : 1), *u4 = u16 + 3 * m;
    const unsigned grid = static_cast<unsigned>((s->n + kThreads - 1) / kThreads);
    row_write_kernel<<<grid, kThreads, 0, stream>>>(s->n, d_row_ptr, u16, sect, u4, s->extra[which ? 1 : 0], d_row, d_word, d_ebase);
    return cudaGetLastError();
}

}  // namespace cyp
