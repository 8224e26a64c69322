// f4: directionality score of the cells in the neighbourhoods along a mean trajectory
// (reference cytopath/trajectory_estimation/cyto_path.py:155-243, `directionality_score`, with the numba
// cosine of cytopath/utils.py:14-39).  For a cell c in the neighbourhood of step j of trajectory i:
//     s(dir) = mean over the entries (c, b, w) of the masked velocity graph of  cos(dir, x_b - x_c) * expm1(w)
//     score  = s(forward)                         at the first step        (:191-203)
//              max(s(forward), s(backward))       in between               (:205-221)
//              s(backward)                        at the last step         (:223-235)
// with forward = coords[j+1] - coords[j], backward = coords[j] - coords[j-1]; cos := 1 when either vector is zero
// (utils.py:24-27); a cell without entries scores nan (np.average of nothing).  The reference loops over steps with
// joblib (:239) and over cells and neighbours in Python.
//
// Kernel: one warp per (trajectory, step, cell) instance; lanes stride over the cell's entries, each computes the
// displacement to its neighbour once and both cosines from it; the warp sums.  Everything is float64; expm1 is
// applied on the host by numpy in the dtype of the graph, as the reference does (:203), so its rounding is the
// reference's.  Gather bound (positions of the neighbours), tiny next to the sampler.
#include <cmath>

#include "common.cuh"

namespace cyp {

__global__ void __launch_bounds__(256)
directionality_kernel(int64_t n_inst, int dims, const double *__restrict__ pos, const int64_t *__restrict__ row_ptr,
                      const int32_t *__restrict__ col, const double *__restrict__ w, const double *__restrict__ dirs,
                      const int32_t *__restrict__ cell, const int32_t *__restrict__ fwd, const int32_t *__restrict__ bwd,
                      double *__restrict__ score) {
    const int64_t inst = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (inst >= n_inst) return;
    const int32_t c = cell[inst];
    const int32_t f = fwd[inst], b = bwd[inst];
    const double *xc = pos + static_cast<int64_t>(c) * dims;
    const double *df = f >= 0 ? dirs + static_cast<int64_t>(f) * dims : nullptr;
    const double *db = b >= 0 ? dirs + static_cast<int64_t>(b) * dims : nullptr;
    double ff = 0.0, bb = 0.0;                       // |forward|^2, |backward|^2
    for (int k = 0; k < dims; ++k) {
        if (df) ff += df[k] * df[k];
        if (db) bb += db[k] * db[k];
    }
    const int64_t a = row_ptr[c], e = row_ptr[c + 1];
    double sf = 0.0, sb = 0.0;
    for (int64_t q = a + lane; q < e; q += 32) {
        const double *xb = pos + static_cast<int64_t>(col[q]) * dims;
        double vv = 0.0, fv = 0.0, bv = 0.0;
        for (int k = 0; k < dims; ++k) {
            const double d = xb[k] - xc[k];
            vv += d * d;
            if (df) fv += df[k] * d;
            if (db) bv += db[k] * d;
        }
        const double wq = w[q];
        if (df) sf += ((ff != 0.0 && vv != 0.0) ? fv / sqrt(ff * vv) : 1.0) * wq;
        if (db) sb += ((bb != 0.0 && vv != 0.0) ? bv / sqrt(bb * vv) : 1.0) * wq;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sf += __shfl_xor_sync(0xffffffffu, sf, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) {
        const double cnt = static_cast<double>(e - a);
        const double mf = sf / cnt, mb = sb / cnt;   // 0 / 0 = nan for a cell without entries, like np.average
        double out;
        if (f >= 0 && b >= 0) out = (mb > mf) ? mb : mf;          // Python's max(forward, backward) (:221)
        else out = f >= 0 ? mf : mb;
        score[inst] = out;
    }
}

}  // namespace cyp

using namespace cyp;

extern "C" int cyp_directionality_scores(int device, int64_t n_cells, int32_t dims, const double *h_pos, const int64_t *h_row_ptr,
                                         const int32_t *h_col, const double *h_w, int64_t n_dirs, const double *h_dirs,
                                         int64_t n_inst, const int32_t *h_cell, const int32_t *h_fwd, const int32_t *h_bwd,
                                         double *h_score, float *kernel_ms) {
    CYP_REQUIRE(n_cells > 0 && dims > 0 && n_dirs >= 0 && n_inst >= 0, "cyp_directionality_scores: bad sizes");
    CYP_REQUIRE(h_pos && h_row_ptr && (n_inst == 0 || (h_cell && h_fwd && h_bwd && h_score)), "cyp_directionality_scores: NULL argument");
    if (kernel_ms) *kernel_ms = 0.f;
    if (n_inst == 0) return CYP_OK;
    const int64_t nnz = h_row_ptr[n_cells];
    CYP_REQUIRE(h_row_ptr[0] == 0 && nnz >= 0 && (nnz == 0 || (h_col && h_w)), "cyp_directionality_scores: bad CSR");
    for (int64_t i = 0; i < n_inst; ++i) {
        CYP_REQUIRE(h_cell[i] >= 0 && h_cell[i] < n_cells, "cyp_directionality_scores: cell out of range");
        CYP_REQUIRE(h_fwd[i] < n_dirs && h_bwd[i] < n_dirs && (h_fwd[i] >= 0 || h_bwd[i] >= 0), "cyp_directionality_scores: bad direction index");
    }
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_directionality_scores: no such CUDA device");
    DeviceGuard guard(device);
    double *d_pos = nullptr, *d_w = nullptr, *d_dirs = nullptr, *d_score = nullptr;
    int64_t *d_row_ptr = nullptr;
    int32_t *d_col = nullptr, *d_cell = nullptr, *d_fwd = nullptr, *d_bwd = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [&](auto **dst, const void *src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = dev_alloc(dst, bytes ? bytes : 1);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up(&d_pos, h_pos, sizeof(double) * n_cells * dims);
    up(&d_row_ptr, h_row_ptr, sizeof(int64_t) * (n_cells + 1));
    up(&d_col, h_col, sizeof(int32_t) * nnz);
    up(&d_w, h_w, sizeof(double) * nnz);
    up(&d_dirs, h_dirs, sizeof(double) * n_dirs * dims);
    up(&d_cell, h_cell, sizeof(int32_t) * n_inst);
    up(&d_fwd, h_fwd, sizeof(int32_t) * n_inst);
    up(&d_bwd, h_bwd, sizeof(int32_t) * n_inst);
    if (e == cudaSuccess) e = dev_alloc(&d_score, sizeof(double) * n_inst);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        const unsigned grid = static_cast<unsigned>((n_inst * 32 + 255) / 256);
        cudaEventRecord(e0);
        directionality_kernel<<<grid, 256>>>(n_inst, dims, d_pos, d_row_ptr, d_col, d_w, d_dirs, d_cell, d_fwd, d_bwd, d_score);
        e = cudaGetLastError();
        cudaEventRecord(e1);
        if (e == cudaSuccess) e = cudaMemcpy(h_score, d_score, sizeof(double) * n_inst, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    dev_free(d_pos); dev_free(d_row_ptr); dev_free(d_col); dev_free(d_w); dev_free(d_dirs);
    dev_free(d_cell); dev_free(d_fwd); dev_free(d_bwd); dev_free(d_score);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_directionality_scores: ") + cudaGetErrorString(e));
    return CYP_OK;
}
