// transition_score (reference cytopath/markov_chain_sampler.py:50,:52):
//   score[row] = sum_t -log(float32(T[c_t, c_t+1])), float32.
// Evaluated only for the rows the caller passes (the rows sampling() returns), one warp per
// row.  T[c_t, c_t+1] is found by binary search in the sorted column indices of row c_t.
// numpy computes the float32 log with a SIMD polynomial whose last bit depends on the host
// CPU, so this kernel is compared with a relative tolerance (tests: 2e-6), not bit-exactly:
// each term is log() in float64 rounded once to float32, the row sum is taken in float64 and
// rounded once.
#include "common.cuh"

namespace cyp {

__global__ void __launch_bounds__(256)
transition_scores_kernel(int64_t n_rows, int32_t n_cols, const int32_t *__restrict__ seq, const RowInfo *__restrict__ rows,
                         int align_elems, const int32_t *__restrict__ col, const double *__restrict__ val,
                         float *__restrict__ score, unsigned long long *__restrict__ n_missing) {
    const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int32_t *row = seq + r * static_cast<int64_t>(n_cols);
    double acc = 0.0;
    unsigned missing = 0;
    for (int t = lane; t + 1 < n_cols; t += 32) {
        const int32_t cur = row[t], nxt = row[t + 1];
        const RowInfo ri = rows[cur];
        const int64_t off = static_cast<int64_t>(ri.off_units) * align_elems;
        int lo = 0, hi = static_cast<int>(ri.len) - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int32_t c = col[off + mid];
            if (c == nxt) { pos = mid; break; }
            if (c < nxt) lo = mid + 1; else hi = mid - 1;
        }
        if (pos < 0) { ++missing; continue; }              // not an edge of T: the reference would take log(0)
        const float p32 = __double2float_rn(val[off + pos]);   // prob_state is a float32 array (:37,:50)
        acc += static_cast<double>(__double2float_rn(-log(static_cast<double>(p32))));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    missing = __reduce_add_sync(0xffffffffu, missing);
    if (lane == 0) {
        score[r] = missing ? __int_as_float(0x7f800000) : __double2float_rn(acc);    // -log(0) = inf
        if (missing) atomicAdd(n_missing, static_cast<unsigned long long>(missing));
    }
}

}  // namespace cyp

using namespace cyp;

extern "C" int cyp_transition_scores(const cyp_graph *g, const int32_t *h_seq, int64_t n_rows, int32_t n_steps,
                                     float *h_score) {
    CYP_REQUIRE(g != nullptr && h_seq != nullptr && h_score != nullptr, "cyp_transition_scores: NULL argument");
    CYP_REQUIRE(n_rows >= 0 && n_steps >= 1, "cyp_transition_scores: bad sizes");
    if (n_rows == 0) return CYP_OK;
    const size_t count = static_cast<size_t>(n_rows) * n_steps;
    for (size_t i = 0; i < count; ++i)
        CYP_REQUIRE(h_seq[i] >= 0 && h_seq[i] < g->n, "cyp_transition_scores: cell index out of range");
    DeviceGuard guard(g->device);
    int32_t *d_seq = nullptr;
    float *d_score = nullptr;
    unsigned long long *d_missing = nullptr;
    cudaError_t e = dev_alloc(&d_seq, sizeof(int32_t) * count);
    if (e == cudaSuccess) e = dev_alloc(&d_score, sizeof(float) * n_rows);
    if (e == cudaSuccess) e = dev_alloc(&d_missing, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(d_missing, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpy(d_seq, h_seq, sizeof(int32_t) * count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const unsigned grid = static_cast<unsigned>((n_rows * 32 + 255) / 256);
        transition_scores_kernel<<<grid, 256>>>(n_rows, n_steps, d_seq, g->d_row, g->align_elems, g->d_col, g->d_val, d_score,
                                                d_missing);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h_score, d_score, sizeof(float) * n_rows, cudaMemcpyDeviceToHost);
    dev_free(d_seq);
    dev_free(d_score);
    dev_free(d_missing);
    if (e != cudaSuccess) return fail(CYP_E_CUDA, std::string("cyp_transition_scores: ") + cudaGetErrorString(e));
    return CYP_OK;
}
