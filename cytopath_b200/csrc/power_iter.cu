// a5: iterate_state_probability (reference cytopath/markov_chain_sampler.py:62-87) on the device.
//
// K4: p <- p @ T repeated max_iter times, recording eps_i = cosine distance(p_i, stationary) (:73-76).
// The matrix arrives as CSR (as stored in adata.uns) and is transposed on the device with a stable
// radix sort of the edges by column (cub::DeviceRadixSort, a library call on a one-off preparation
// step; scipy's single-threaded csr_tocsc took 1.3 s of a 6 s sampling() call at 1M cells).  Output j
// is then the dot product of p with column j, accumulated strictly in increasing row order with
// separate multiply and add: the same order and rounding as scipy's csc_matvec, which is what
// `ndarray @ csr_matrix` (:73) runs.  The three dot
// products of the cosine are reduced in a fixed order (per-block partials, then one block), so the
// result is deterministic; they differ from BLAS' association by O(1e-16), which only matters if a
// |difference of eps| lands within that distance of `tol` (:55-58).  The host derives max_steps from
// the eps history exactly like check_convergence_criteria (:54-60).
#include <cub/device/device_radix_sort.cuh>
#include <vector>

#include "common.cuh"

namespace cyp {

constexpr int kPiThreads = 256;

// One thread per column, reading its run of the CSC arrays straight from HBM: 0.40 ms per iteration at C4 (612 MB, 23 % of
// the HBM roofline: 32 lanes touch 32 different sectors per load, the L1 keeps the rest of each sector for the next
// iterations of the lane).  Tried in round 2 and dropped: one CTA per 64 columns forming the products with coalesced loads
// into shared memory and adding them up per column from there -- 0.70 ms per iteration (profiles/r02_k2_experiments.md, 6).
__global__ void __launch_bounds__(kPiThreads)
spmv_t_kernel(int64_t n, const int64_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx,
              const double *__restrict__ val, const double *__restrict__ p, double *__restrict__ p_new) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * kPiThreads + threadIdx.x;
    if (j >= n) return;
    double acc = 0.0;
    for (int64_t e = col_ptr[j]; e < col_ptr[j + 1]; ++e) acc = __dadd_rn(acc, __dmul_rn(val[e], p[row_idx[e]]));
    p_new[j] = acc;
}

// row id and identity permutation of every edge (one thread per row)
__global__ void __launch_bounds__(kPiThreads)
expand_rows_kernel(int64_t n, const int64_t *__restrict__ row_ptr, int32_t *__restrict__ row_of_edge, int32_t *__restrict__ edge_id) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kPiThreads + threadIdx.x;
    if (i >= n) return;
    for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        row_of_edge[e] = static_cast<int32_t>(i);
        edge_id[e] = static_cast<int32_t>(e);
    }
}

// after the stable sort by column: gather row ids and values, and mark where each column starts
__global__ void __launch_bounds__(kPiThreads)
gather_csc_kernel(int64_t nnz, const int32_t *__restrict__ sorted_col, const int32_t *__restrict__ perm,
                  const int32_t *__restrict__ row_of_edge, const double *__restrict__ val, int32_t *__restrict__ row_idx,
                  double *__restrict__ val_t, int64_t *__restrict__ col_ptr, int64_t n) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kPiThreads + threadIdx.x;
    if (k >= nnz) return;
    const int32_t e = perm[k];
    row_idx[k] = row_of_edge[e];
    val_t[k] = val[e];
    const int32_t c = sorted_col[k];
    const int32_t prev = (k == 0) ? -1 : sorted_col[k - 1];
    for (int32_t j = prev + 1; j <= c; ++j) col_ptr[j] = k;          // columns prev+1 .. c start here (empty ones included)
    if (k == nnz - 1)
        for (int64_t j = c + 1; j <= n; ++j) col_ptr[j] = nnz;
}

// partial[b] = {sum p*s, sum p*p, sum s*s} over the block's slice
__global__ void __launch_bounds__(kPiThreads)
dots_partial_kernel(int64_t n, const double *__restrict__ p, const double *__restrict__ s, double *__restrict__ partial) {
    __shared__ double sh[3][kPiThreads];
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kPiThreads + threadIdx.x;
    const double a = (i < n) ? p[i] : 0.0, b = (i < n) ? s[i] : 0.0;
    sh[0][threadIdx.x] = a * b;
    sh[1][threadIdx.x] = a * a;
    sh[2][threadIdx.x] = b * b;
    __syncthreads();
    for (int d = kPiThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d)
            for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[3 * static_cast<int64_t>(blockIdx.x) + threadIdx.x] = sh[threadIdx.x][0];
}

// one block: fixed-order sum of the partials, eps = 1 - uv / sqrt(uu * vv), clipped to [0, 2] like scipy
__global__ void __launch_bounds__(kPiThreads)
dots_final_kernel(int64_t n_blocks, const double *__restrict__ partial, double *__restrict__ eps_out) {
    __shared__ double sh[3][kPiThreads];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t b = threadIdx.x; b < n_blocks; b += kPiThreads)
        for (int k = 0; k < 3; ++k) acc[k] += partial[3 * b + k];
    for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int d = kPiThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d)
            for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double dist = 1.0 - sh[0][0] / sqrt(sh[1][0] * sh[2][0]);
        dist = fmin(fmax(dist, 0.0), 2.0);
        *eps_out = dist;
    }
}

}  // namespace cyp

using namespace cyp;

namespace {

// K4 on device arrays: compact CSR (d_row_ptr [n+1], d_col / d_val [nnz]) already resident on the current device
int power_iteration_device(int64_t n_cells, int64_t nnz, const int64_t *d_row_ptr, const int32_t *d_col, const double *d_val,
                           const double *h_init, const double *h_stationary, int32_t max_iter, double *h_eps_history, double *h_final_state,
                           double *h_history = nullptr) {
    int64_t *d_col_ptr = nullptr;
    int32_t *d_col_sorted = nullptr, *d_edge = nullptr, *d_perm = nullptr, *d_row_of = nullptr, *d_row_idx = nullptr;
    double *d_val_t = nullptr, *d_p = nullptr, *d_q = nullptr, *d_s = nullptr, *d_partial = nullptr, *d_eps = nullptr;
    void *d_tmp = nullptr;
    const unsigned blocks = static_cast<unsigned>((n_cells + kPiThreads - 1) / kPiThreads);
    const unsigned eblocks = static_cast<unsigned>((nnz + kPiThreads - 1) / kPiThreads);
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(dev_alloc(&d_col_ptr, sizeof(int64_t) * (n_cells + 1)));
    step(dev_alloc(&d_col_sorted, sizeof(int32_t) * nnz));
    step(dev_alloc(&d_edge, sizeof(int32_t) * nnz));
    step(dev_alloc(&d_perm, sizeof(int32_t) * nnz));
    step(dev_alloc(&d_row_of, sizeof(int32_t) * nnz));
    step(dev_alloc(&d_row_idx, sizeof(int32_t) * nnz));
    step(dev_alloc(&d_val_t, sizeof(double) * nnz));
    step(dev_alloc(&d_p, sizeof(double) * n_cells));
    step(dev_alloc(&d_q, sizeof(double) * n_cells));
    step(dev_alloc(&d_s, sizeof(double) * n_cells));
    step(dev_alloc(&d_partial, sizeof(double) * 3 * blocks));
    step(dev_alloc(&d_eps, sizeof(double) * (max_iter > 0 ? max_iter : 1)));
    size_t tmp_bytes = 0;
    int key_bits = 1;
    while ((1ll << key_bits) < n_cells) ++key_bits;
    if (e == cudaSuccess)
        step(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_col, d_col_sorted, d_edge, d_perm, static_cast<int>(nnz), 0, key_bits));
    step(dev_alloc(&d_tmp, tmp_bytes));
    if (e == cudaSuccess) {
        step(cudaMemcpy(d_p, h_init, sizeof(double) * n_cells, cudaMemcpyHostToDevice));
        step(cudaMemcpy(d_s, h_stationary, sizeof(double) * n_cells, cudaMemcpyHostToDevice));
    }
    if (e == cudaSuccess) {
        // transpose: stable sort of the edges by column keeps rows ascending inside every column
        expand_rows_kernel<<<blocks, kPiThreads>>>(n_cells, d_row_ptr, d_row_of, d_edge);
        step(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_col, d_col_sorted, d_edge, d_perm, static_cast<int>(nnz), 0, key_bits));
        gather_csc_kernel<<<eblocks, kPiThreads>>>(nnz, d_col_sorted, d_perm, d_row_of, d_val, d_row_idx, d_val_t, d_col_ptr, n_cells);
        for (int32_t i = 0; i < max_iter; ++i) {
            if (h_history) step(cudaMemcpy(h_history + static_cast<int64_t>(i) * n_cells, d_p, sizeof(double) * n_cells, cudaMemcpyDeviceToHost));   // :71 history[i] = p
            spmv_t_kernel<<<blocks, kPiThreads>>>(n_cells, d_col_ptr, d_row_idx, d_val_t, d_p, d_q);
            dots_partial_kernel<<<blocks, kPiThreads>>>(n_cells, d_q, d_s, d_partial);
            dots_final_kernel<<<1, kPiThreads>>>(blocks, d_partial, d_eps + i);
            double *t = d_p; d_p = d_q; d_q = t;
        }
        step(cudaGetLastError());
        if (max_iter > 0) step(cudaMemcpy(h_eps_history, d_eps, sizeof(double) * max_iter, cudaMemcpyDeviceToHost));
        if (h_final_state) step(cudaMemcpy(h_final_state, d_p, sizeof(double) * n_cells, cudaMemcpyDeviceToHost));
    }
    dev_free(d_col_ptr); dev_free(d_col_sorted); dev_free(d_edge); dev_free(d_perm);
    dev_free(d_row_of); dev_free(d_row_idx); dev_free(d_val_t); dev_free(d_p); dev_free(d_q); dev_free(d_s);
    dev_free(d_partial); dev_free(d_eps); dev_free(d_tmp);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_power_iteration: ") + cudaGetErrorString(e));
    return CYP_OK;
}

// padded rows of a graph -> compact CSR order (one warp per row)
__global__ void __launch_bounds__(kPiThreads)
unpad_graph_kernel(int64_t n, const int64_t *__restrict__ row_ptr, const RowInfo *__restrict__ rows, int align_elems,
                   const int32_t *__restrict__ col, const double *__restrict__ val, int32_t *__restrict__ col_out, double *__restrict__ val_out) {
    const int64_t r = (static_cast<int64_t>(blockIdx.x) * kPiThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const int64_t a = row_ptr[r], len = row_ptr[r + 1] - a;
    const int64_t src = static_cast<int64_t>(rows[r].off_units) * align_elems;
    for (int64_t j = lane; j < len; j += 32) {
        col_out[a + j] = col[src + j];
        val_out[a + j] = val[src + j];
    }
}

}  // namespace

// Same on the matrix a graph already holds in HBM (no second upload of the CSR): valid when the matrix stored in
// adata.uns[matrix_key] IS the canonical matrix the graph was built from (the usual case; float32 values were widened
// exactly, like scipy widens them for `p @ T`).
extern "C" int cyp_graph_power_iteration(const cyp_graph *g, const double *h_init, const double *h_stationary, int32_t max_iter,
                                         double *h_eps_history, double *h_final_state) {
    CYP_REQUIRE(g && h_init && h_stationary && h_eps_history, "cyp_graph_power_iteration: NULL argument");
    CYP_REQUIRE(g->nnz < (1ll << 31) && max_iter >= 0, "cyp_graph_power_iteration: bad sizes");
    DeviceGuard guard(g->device);
    int32_t *d_col = nullptr;
    double *d_val = nullptr;
    cudaError_t e = dev_alloc(&d_col, sizeof(int32_t) * g->nnz);
    if (e == cudaSuccess) e = dev_alloc(&d_val, sizeof(double) * g->nnz);
    int rc = CYP_OK;
    if (e == cudaSuccess) {
        unpad_graph_kernel<<<static_cast<unsigned>((g->n * 32 + kPiThreads - 1) / kPiThreads), kPiThreads>>>(
            g->n, g->d_row_ptr, g->d_row, g->align_elems, g->d_col, g->d_val, d_col, d_val);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) rc = power_iteration_device(g->n, g->nnz, g->d_row_ptr, d_col, d_val, h_init, h_stationary, max_iter, h_eps_history, h_final_state);
    dev_free(d_col);
    dev_free(d_val);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_graph_power_iteration: ") + cudaGetErrorString(e));
    return rc;
}

extern "C" int cyp_power_iteration(int device, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr,
                                   const int32_t *h_col, const double *h_val, const double *h_init,
                                   const double *h_stationary, int32_t max_iter, double *h_eps_history,
                                   double *h_final_state, double *h_history) {
    CYP_REQUIRE(h_row_ptr && h_col && h_val && h_init && h_stationary && h_eps_history, "cyp_power_iteration: NULL argument");
    CYP_REQUIRE(n_cells > 0 && nnz > 0 && nnz < (1ll << 31) && max_iter >= 0, "cyp_power_iteration: bad sizes");
    CYP_REQUIRE(h_row_ptr[0] == 0 && h_row_ptr[n_cells] == nnz, "cyp_power_iteration: row_ptr does not span nnz");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_power_iteration: no such CUDA device");
    DeviceGuard guard(device);
    int64_t *d_row_ptr = nullptr;
    int32_t *d_col = nullptr;
    double *d_val = nullptr;
    cudaError_t e = dev_alloc(&d_row_ptr, sizeof(int64_t) * (n_cells + 1));
    if (e == cudaSuccess) e = dev_alloc(&d_col, sizeof(int32_t) * nnz);
    if (e == cudaSuccess) e = dev_alloc(&d_val, sizeof(double) * nnz);
    if (e == cudaSuccess) e = cudaMemcpy(d_row_ptr, h_row_ptr, sizeof(int64_t) * (n_cells + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_col, h_col, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_val, h_val, sizeof(double) * nnz, cudaMemcpyHostToDevice);
    int rc = CYP_OK;
    if (e == cudaSuccess) rc = power_iteration_device(n_cells, nnz, d_row_ptr, d_col, d_val, h_init, h_stationary, max_iter, h_eps_history, h_final_state, h_history);
    dev_free(d_row_ptr); dev_free(d_col); dev_free(d_val);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_power_iteration: ") + cudaGetErrorString(e));
    return rc;
}
