// f2: pairwise Hausdorff distances between sampled trajectories — the step right after
// sampling (reference cytopath/trajectory_estimation/sample_clustering.py:57-71, `distance_matrix`):
//   D[i, j] = hausdorff_distance(chain_j, chain_i)   for all pairs of S chains of L points in d dimensions,
// where hausdorff_distance (PyPI `hausdorff`, a numba loop; absent from this image, algorithm restated in
// oracle/hausdorff_port.py) = max( max_p min_q dist(a_p, b_q), max_q min_p dist(a_p, b_q) ).
// The reference runs S(S+1)/2 Python-level calls of an O(L^2 d) loop.
//
// Kernel: one CTA per pair (i <= j).  A directed pass gives every thread one point of X (coordinates in
// shared memory, odd row stride: conflict-free) and streams Y through shared memory in transposed chunks,
// so the four Y values a thread needs per coordinate are two broadcast 16-byte loads; each thread keeps the
// running minimum of its point in registers, the CTA reduces the maximum.  The pass runs twice with the
// roles swapped.  All arithmetic is float64.  Bound: FP64 pipe (S^2 L^2 d flops), not memory.
#include <cfloat>

#include "common.cuh"

namespace cyp {

constexpr int kHdThreads = 128;     // points of X per tile
constexpr int kHdChunk = 64;        // points of Y per shared-memory chunk

enum HdMetric : int { kHdEuclidean = 0, kHdManhattan = 1, kHdChebyshev = 2, kHdCosine = 3 };

template <int METRIC>
__device__ __forceinline__ double directed_hausdorff(const double *__restrict__ X, const double *__restrict__ Y, int L, int d,
                                                     double *sX, double *sYT, double *sRed) {
    const int t = threadIdx.x;
    const int strideX = d | 1;
    double result = 0.0;
    for (int x0 = 0; x0 < L; x0 += kHdThreads) {
        const int nx = min(kHdThreads, L - x0);
        __syncthreads();
        for (int e = t; e < nx * d; e += kHdThreads) sX[(e / d) * strideX + (e % d)] = X[static_cast<int64_t>(x0) * d + e];
        double xnorm = 0.0;
        double minv = DBL_MAX;
        for (int y0 = 0; y0 < L; y0 += kHdChunk) {
            const int ny = min(kHdChunk, L - y0);
            __syncthreads();
            for (int e = t; e < kHdChunk * d; e += kHdThreads) {     // transposed: sYT[k][q]; points beyond ny repeat the last one
                const int q = e / d, k = e % d;
                sYT[k * kHdChunk + q] = Y[static_cast<int64_t>(y0 + min(q, ny - 1)) * d + k];
            }
            __syncthreads();
            if (t < nx) {
                const double *xp = sX + t * strideX;
                if (METRIC == kHdCosine && y0 == 0)
                    for (int k = 0; k < d; ++k) xnorm += xp[k] * xp[k];
                for (int q = 0; q < ny; q += 4) {
                    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0;
                    for (int k = 0; k < d; ++k) {
                        const double x = xp[k];
                        const double2 y01 = *reinterpret_cast<const double2 *>(sYT + k * kHdChunk + q);
                        const double2 y23 = *reinterpret_cast<const double2 *>(sYT + k * kHdChunk + q + 2);
                        if (METRIC == kHdEuclidean) {
                            const double t0 = x - y01.x, t1 = x - y01.y, t2 = x - y23.x, t3 = x - y23.y;
                            a0 += t0 * t0; a1 += t1 * t1; a2 += t2 * t2; a3 += t3 * t3;
                        } else if (METRIC == kHdManhattan) {
                            a0 += fabs(x - y01.x); a1 += fabs(x - y01.y); a2 += fabs(x - y23.x); a3 += fabs(x - y23.y);
                        } else if (METRIC == kHdChebyshev) {
                            a0 = fmax(a0, fabs(x - y01.x)); a1 = fmax(a1, fabs(x - y01.y));
                            a2 = fmax(a2, fabs(x - y23.x)); a3 = fmax(a3, fabs(x - y23.y));
                        } else {
                            a0 += x * y01.x; a1 += x * y01.y; a2 += x * y23.x; a3 += x * y23.y;
                            n0 += y01.x * y01.x; n1 += y01.y * y01.y; n2 += y23.x * y23.x; n3 += y23.y * y23.y;
                        }
                    }
                    if (METRIC == kHdCosine) {
                        a0 = 1.0 - a0 / sqrt(xnorm * n0); a1 = 1.0 - a1 / sqrt(xnorm * n1);
                        a2 = 1.0 - a2 / sqrt(xnorm * n2); a3 = 1.0 - a3 / sqrt(xnorm * n3);
                    }
                    minv = fmin(fmin(minv, fmin(a0, a1)), fmin(a2, a3));   // padded points repeat a real one: harmless
                }
            }
        }
        if (t < nx) result = fmax(result, minv);
    }
    // CTA max
    for (int o = 16; o > 0; o >>= 1) result = fmax(result, __shfl_xor_sync(0xffffffffu, result, o));
    __syncthreads();
    if ((t & 31) == 0) sRed[t >> 5] = result;
    __syncthreads();
    double r = sRed[0];
    for (int w = 1; w < kHdThreads / 32; ++w) r = fmax(r, sRed[w]);
    return (METRIC == kHdEuclidean) ? sqrt(r) : r;     // sqrt is monotone: min/max commute with it
}

template <int METRIC>
__global__ void __launch_bounds__(kHdThreads)
hausdorff_kernel(const double *__restrict__ coords, int64_t S, int L, int d, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    double *sX = smem;
    double *sYT = sX + kHdThreads * (d | 1) + ((kHdThreads * (d | 1)) & 1);    // keep 16-byte alignment
    __shared__ double sRed[kHdThreads / 32];
    // pair index -> (i, j), i <= j, row-major over the upper triangle
    const int64_t pidx = blockIdx.x;
    int64_t i = static_cast<int64_t>((2.0 * S + 1.0 - sqrt((2.0 * S + 1.0) * (2.0 * S + 1.0) - 8.0 * static_cast<double>(pidx))) * 0.5);
    while (i > 0 && i * S - i * (i - 1) / 2 > pidx) --i;
    while ((i + 1) * S - (i + 1) * i / 2 <= pidx) ++i;
    const int64_t j = i + (pidx - (i * S - i * (i - 1) / 2));
    const double *A = coords + i * static_cast<int64_t>(L) * d;
    const double *B = coords + j * static_cast<int64_t>(L) * d;
    double h = 0.0;
    if (i != j) {
        const double hab = directed_hausdorff<METRIC>(A, B, L, d, sX, sYT, sRed);
        const double hba = directed_hausdorff<METRIC>(B, A, L, d, sX, sYT, sRed);
        h = fmax(hab, hba);
    }
    if (threadIdx.x == 0) {
        out[i * S + j] = h;
        out[j * S + i] = h;
    }
}

}  // namespace cyp

using namespace cyp;

extern "C" int cyp_hausdorff_matrix(int device, const double *h_coords, int64_t n_chains, int32_t chain_len, int32_t dim,
                                    int metric, double *h_out, float *kernel_ms) {
    CYP_REQUIRE(h_coords && h_out, "cyp_hausdorff_matrix: NULL argument");
    CYP_REQUIRE(n_chains > 0 && chain_len > 0 && dim > 0, "cyp_hausdorff_matrix: bad sizes");
    CYP_REQUIRE(metric >= 0 && metric <= 3, "cyp_hausdorff_matrix: unknown metric");
    CYP_REQUIRE(n_chains * (n_chains + 1) / 2 < (1ll << 31), "cyp_hausdorff_matrix: too many chains for one launch");
    const size_t smem = sizeof(double) * (static_cast<size_t>(kHdThreads) * (dim | 1) + 1 + static_cast<size_t>(kHdChunk) * dim);
    CYP_REQUIRE(smem <= 220 * 1024, "cyp_hausdorff_matrix: dimension too large for the shared-memory tiles");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_hausdorff_matrix: no such CUDA device");
    DeviceGuard guard(device);
    double *d_coords = nullptr, *d_out = nullptr;
    const size_t n_in = static_cast<size_t>(n_chains) * chain_len * dim, n_out = static_cast<size_t>(n_chains) * n_chains;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = dev_alloc(&d_coords, sizeof(double) * n_in);
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(dev_alloc(&d_out, sizeof(double) * n_out));
    step(cudaMemcpy(d_coords, h_coords, sizeof(double) * n_in, cudaMemcpyHostToDevice));
    step(cudaEventCreate(&e0));
    step(cudaEventCreate(&e1));
    if (e == cudaSuccess) {
        const unsigned grid = static_cast<unsigned>(n_chains * (n_chains + 1) / 2);
        step(cudaEventRecord(e0));
#define HD_LAUNCH(M)                                                                                              \
    step(cudaFuncSetAttribute(hausdorff_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    if (e == cudaSuccess) hausdorff_kernel<M><<<grid, kHdThreads, smem>>>(d_coords, n_chains, chain_len, dim, d_out);
        switch (metric) {
            case kHdEuclidean: HD_LAUNCH(kHdEuclidean) break;
            case kHdManhattan: HD_LAUNCH(kHdManhattan) break;
            case kHdChebyshev: HD_LAUNCH(kHdChebyshev) break;
            default: HD_LAUNCH(kHdCosine) break;
        }
#undef HD_LAUNCH
        step(cudaGetLastError());
        step(cudaEventRecord(e1));
        step(cudaMemcpy(h_out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost));
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && kernel_ms) *kernel_ms = ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    dev_free(d_coords);
    dev_free(d_out);
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_hausdorff_matrix: ") + cudaGetErrorString(e));
    return CYP_OK;
}
