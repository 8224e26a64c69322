// Random-gather ceilings of one B200: what the memory system sustains for independent random reads of
// S contiguous bytes (S = 32 .. 512) out of a working set of W bytes, as G gathers/s and TB/s.
// K2 (the walker step kernel) is a stream of dependent random gathers (one row of CDF codes, then one edge
// record per step); this program measures the ceiling for that access pattern with the dependency removed
// (every thread group has `ILP` independent gathers in flight), so K2's gather rate can be put next to it.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench scripts/gather_bench.cu
//   ./gather_bench            (prints a table; ~10 s)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint4 ldg_na(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t xorshift32(uint32_t &s) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}

// "fast" flavour: the index costs ~6 integer instructions (xorshift32 + multiply-shift), so that the memory system
// and not the index arithmetic is what the kernel measures.  Every lane of a gather's group computes the same index.
template <int LPG, int ILP>
__global__ void __launch_bounds__(1024) gather_fast_kernel(const uint4 *__restrict__ data, uint32_t n_slots, int iters, uint32_t *__restrict__ sink) {
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

// "cpasync" flavour: the same gathers, but through cp.async (LDGSTS, L1 bypass) into shared memory, ILP 16-byte copies
// per lane in flight, then cp.async.wait_group 0 -- the data path K2 uses for its rows.
template <int LPG, int ILP>
__global__ void __launch_bounds__(1024) gather_cpasync_kernel(const uint4 *__restrict__ data, uint32_t n_slots, int iters, uint32_t *__restrict__ sink) {
    extern __shared__ uint4 sbuf[];                                  // [blockDim.x * ILP]
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = gthread / LPG;
    const int sub = static_cast<int>(gthread % LPG);
    uint32_t state = group * 2654435761u + 12345u;
    if (state == 0) state = 1;
    uint32_t acc = 0;
    uint4 *mine = sbuf + static_cast<size_t>(threadIdx.x) * ILP;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(mine));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const uint32_t slot = static_cast<uint32_t>((static_cast<uint64_t>(xorshift32(state)) * n_slots) >> 32);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 16), "l"(data + static_cast<uint64_t>(slot) * LPG + sub) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        acc ^= mine[it % ILP].x;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// LPG lanes (16 bytes each) read one gather of S = 16 * LPG bytes; ILP gathers in flight per lane group.
template <int LPG, int ILP>
__global__ void __launch_bounds__(1024) gather_kernel(const uint4 *__restrict__ data, uint64_t n_slots /* working set / S */, int iters,
                                                      uint32_t *__restrict__ sink) {
    const uint64_t gthread = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t group = gthread / LPG;
    const int sub = static_cast<int>(gthread % LPG);
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const uint64_t h = mix((group * static_cast<uint64_t>(iters) + it) * ILP + k + 0x9E3779B97F4A7C15ull);
            const uint64_t slot = h % n_slots;
            v[k] = ldg_na(data + slot * LPG + sub);
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// Same, but each lane group follows a dependent chain (next slot derived from the loaded data): the latency-bound
// shape of a walker.  CHAINS independent chains per lane group.
template <int LPG, int CHAINS>
__global__ void __launch_bounds__(1024) chase_kernel(const uint4 *__restrict__ data, uint64_t n_slots, int iters, uint32_t *__restrict__ sink) {
    const uint64_t gthread = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t group = gthread / LPG;
    const int sub = static_cast<int>(gthread % LPG);
    uint64_t slot[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) slot[k] = mix(group * CHAINS + k) % n_slots;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            const uint4 v = ldg_na(data + slot[k] * LPG + sub);
            uint32_t x = v.x ^ v.y ^ v.z ^ v.w;
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, o);     // the group agrees on the next slot
            acc ^= x;
            slot[k] = mix(slot[k] + x + it) % n_slots;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <typename K>
static double time_kernel(K launch, int reps = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t max_bytes = 4ull << 30;
    uint4 *data; uint32_t *sink;
    CK(cudaMalloc(&data, max_bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(data, 0x5a, max_bytes));
    printf("SMs %d\n", sms);
    printf("%-10s %6s %8s %5s %5s %10s %9s\n", "kind", "S", "set", "thr", "ilp", "G gath/s", "TB/s");
    const size_t sets[] = {32ull << 20, 128ull << 20, 320ull << 20, 1200ull << 20, 3800ull << 20};
#define RUN_G(LPG, ILP, THREADS)                                                                                   \
    for (size_t ws : sets) {                                                                                        \
        const int iters = 2000 / ILP;                                                                               \
        const uint64_t n_slots = ws / (16 * LPG);                                                                   \
        const double ms = time_kernel([&] { gather_kernel<LPG, ILP><<<sms, THREADS>>>(data, n_slots, iters, sink); }); \
        const double gathers = static_cast<double>(sms) * THREADS / LPG * iters * ILP;                              \
        printf("%-10s %6d %6zuMB %5d %5d %10.2f %9.3f\n", "indep", 16 * LPG, ws >> 20, THREADS, ILP, gathers / ms / 1e6, gathers * 16 * LPG / ms / 1e9); \
    }
#define RUN_F(LPG, ILP, THREADS)                                                                                   \
    for (size_t ws : sets) {                                                                                        \
        const int iters = 4000 / ILP;                                                                               \
        const uint32_t n_slots = static_cast<uint32_t>(ws / (16 * LPG));                                            \
        const double ms = time_kernel([&] { gather_fast_kernel<LPG, ILP><<<sms, THREADS>>>(data, n_slots, iters, sink); }); \
        const double gathers = static_cast<double>(sms) * THREADS / LPG * iters * ILP;                              \
        printf("%-10s %6d %6zuMB %5d %5d %10.2f %9.3f\n", "fast", 16 * LPG, ws >> 20, THREADS, ILP, gathers / ms / 1e6, gathers * 16 * LPG / ms / 1e9); \
    }
#define RUN_A(LPG, ILP, THREADS)                                                                                   \
    for (size_t ws : sets) {                                                                                        \
        const int iters = 4000 / ILP;                                                                               \
        const uint32_t n_slots = static_cast<uint32_t>(ws / (16 * LPG));                                            \
        const size_t smem = static_cast<size_t>(THREADS) * ILP * 16;                                                \
        CK(cudaFuncSetAttribute(gather_cpasync_kernel<LPG, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
        const double ms = time_kernel([&] { gather_cpasync_kernel<LPG, ILP><<<sms, THREADS, smem>>>(data, n_slots, iters, sink); }); \
        const double gathers = static_cast<double>(sms) * THREADS / LPG * iters * ILP;                              \
        printf("%-10s %6d %6zuMB %5d %5d %10.2f %9.3f\n", "cpasync", 16 * LPG, ws >> 20, THREADS, ILP, gathers / ms / 1e6, gathers * 16 * LPG / ms / 1e9); \
    }
    RUN_A(8, 1, 1024) RUN_A(8, 2, 1024) RUN_A(8, 4, 1024) RUN_A(8, 8, 1024) RUN_A(2, 8, 1024)
    RUN_F(8, 1, 1024) RUN_F(8, 2, 1024)
    RUN_F(2, 8, 1024) RUN_F(4, 8, 1024) RUN_F(8, 8, 1024) RUN_F(8, 4, 1024) RUN_F(8, 16, 512) RUN_F(16, 8, 1024)
    RUN_G(1, 8, 1024) RUN_G(2, 8, 1024) RUN_G(4, 8, 1024) RUN_G(8, 8, 1024) RUN_G(16, 8, 1024) RUN_G(32, 8, 1024)
    RUN_G(2, 4, 1024) RUN_G(8, 4, 1024) RUN_G(2, 16, 1024) RUN_G(8, 16, 1024)
#define RUN_C(LPG, CH, THREADS)                                                                                    \
    for (size_t ws : sets) {                                                                                        \
        const int iters = 1000;                                                                                     \
        const uint64_t n_slots = ws / (16 * LPG);                                                                   \
        const double ms = time_kernel([&] { chase_kernel<LPG, CH><<<sms, THREADS>>>(data, n_slots, iters, sink); }); \
        const double gathers = static_cast<double>(sms) * THREADS / LPG * iters * CH;                               \
        printf("%-10s %6d %6zuMB %5d %5d %10.2f %9.3f\n", "chase", 16 * LPG, ws >> 20, THREADS, CH, gathers / ms / 1e6, gathers * 16 * LPG / ms / 1e9); \
    }
    RUN_C(1, 1, 1024) RUN_C(1, 2, 1024) RUN_C(1, 4, 1024) RUN_C(2, 4, 1024) RUN_C(8, 4, 1024) RUN_C(8, 8, 1024)
    // mixed like K2 at C4: per step one 128-byte row gather (1.3 MB.. set) and one 32-byte sector gather, dependent
    return 0;
}
