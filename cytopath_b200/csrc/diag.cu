// Diagnostics for the measurement contract (bench.py): the random-gather ceiling of the device the step kernel is
// compared with.  K2 is a stream of DEPENDENT random gathers of one row block per walker-step; this kernel issues
// the same gathers with the dependency removed (8 independent gathers in flight per lane group, index arithmetic of
// ~6 instructions), out of a working set of the same size.  Its rate is the bound no walker kernel on this layout can
// beat; K2's rows/s divided by it is `roofline.gather_ceiling_frac`.
#include "common.cuh"

namespace cyp {
namespace {

__device__ __forceinline__ uint4 ldg_na(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t xorshift32(uint32_t &s) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}

// LPG lanes (16 bytes each) read one gather of 16 * LPG bytes; ILP gathers in flight per lane group
template <int LPG, int ILP>
__global__ void __launch_bounds__(1024) gather_ceiling_kernel(const uint4 *__restrict__ data, uint32_t n_slots, int iters, uint32_t *__restrict__ sink) {
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = gthread / LPG;
    const int sub = static_cast<int>(gthread % LPG);
    uint32_t state = group * 2654435761u + 12345u;
    if (state == 0) state = 1;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const uint32_t slot = static_cast<uint32_t>((static_cast<uint64_t>(xorshift32(state)) * n_slots) >> 32);
            v[k] = ldg_na(data + static_cast<uint64_t>(slot) * LPG + sub);
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int LPG>
cudaError_t run_ceiling(const uint4 *d, uint32_t n_slots, int sms, uint32_t *sink, float *best_ms, double *gathers) {
    constexpr int ILP = 8, THREADS = 1024;
    const int iters = 4000 / ILP;
    cudaEvent_t a, b;
    cudaError_t e = cudaEventCreate(&a);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&b);
    if (e != cudaSuccess) { cudaEventDestroy(a); return e; }
    *best_ms = 1e30f;
    for (int r = 0; r < 4 && e == cudaSuccess; ++r) {          // first repetition warms up
        cudaEventRecord(a);
        gather_ceiling_kernel<LPG, ILP><<<sms, THREADS>>>(d, n_slots, iters, sink);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < *best_ms) *best_ms = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *gathers = static_cast<double>(sms) * THREADS / LPG * iters * ILP;
    return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace
}  // namespace cyp

using namespace cyp;

extern "C" int cyp_gather_ceiling(int device, int64_t working_set_bytes, int32_t gather_bytes, double *gathers_per_s) {
    CYP_REQUIRE(gathers_per_s != nullptr, "cyp_gather_ceiling: NULL argument");
    CYP_REQUIRE(working_set_bytes >= (1 << 20) && working_set_bytes <= (64ll << 30), "cyp_gather_ceiling: working set out of range");
    CYP_REQUIRE(gather_bytes == 32 || gather_bytes == 64 || gather_bytes == 128 || gather_bytes == 256, "cyp_gather_ceiling: gathers of 32, 64, 128 or 256 bytes");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_gather_ceiling: no such CUDA device");
    DeviceGuard guard(device);
    int sms = 0;
    CYP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint4 *d = nullptr;
    uint32_t *sink = nullptr;
    cudaError_t e = dev_alloc(&d, static_cast<size_t>(working_set_bytes));
    if (e == cudaSuccess) e = dev_alloc(&sink, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(d, 0x5a, static_cast<size_t>(working_set_bytes));
    float ms = 0.f;
    double gathers = 0.0;
    const uint32_t n_slots = static_cast<uint32_t>(working_set_bytes / gather_bytes);
    if (e == cudaSuccess) {
        switch (gather_bytes) {
            case 32: e = run_ceiling<2>(d, n_slots, sms, sink, &ms, &gathers); break;
            case 64: e = run_ceiling<4>(d, n_slots, sms, sink, &ms, &gathers); break;
            case 128: e = run_ceiling<8>(d, n_slots, sms, sink, &ms, &gathers); break;
            default: e = run_ceiling<16>(d, n_slots, sms, sink, &ms, &gathers); break;
        }
    }
    dev_free(d);
    dev_free(sink);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string("cyp_gather_ceiling: ") + cudaGetErrorString(e));
    *gathers_per_s = gathers / (ms * 1e-3);
    return CYP_OK;
}
