// a6 / (e): the reference's joblib fan-out (cytopath/markov_chain_sampler.py:306-319) as a GROUP of devices that
// share one sampling() run.
//
// The walkers of a round are cut into contiguous id blocks, one per rank (lower rank = lower ids); the graph is
// replicated.  Because the uniform stream is keyed by (walker, step, round) the result does not depend on the cut,
// and because a kept sequence is stored as its 16-byte (round, walker, slot) record (session.cu) the ranks exchange
// records, not rows:
//   graph      every rank uploads 1/world of the host CSR over its own PCIe link, ncclAllGather completes it on every
//              device and each builds the graph itself (cyp_group_graph_create_sharded); or, when only one rank holds the
//              matrix, built there and ncclBroadcast of its device arrays (1.6 GB at 1M cells, a few ms over NVLink) --
//              never one full PCIe upload per GPU;
//   per round  ncclAllGather of every rank's (hash_lo, hash_hi, walker) records -> each rank drops its sequences that
//              also occur under a lower walker id elsewhere (first occurrence, :330-331; equality is VERIFIED by
//              replaying both walkers, see session.cu) -> ncclAllGather of the survivors' (walker, slot), so that
//              every rank holds the global index of kept sequences in accumulation order (:335-340);
//   selection  the rows the final draw picks (:395-403) are split over the ranks, replayed there, and gathered with
//              ncclAllGather of the int32 sequences (north-star item 4).
// A group lives either in ONE process that drives N devices (ncclCommInitAll, one host thread per device inside a
// call: the boundary is a notebook function call, there is no launcher) or in N processes of one GPU each
// (ncclCommInitRank; the 128-byte id travels through whatever the host already has, e.g. torch.distributed).
// NCCL is bound at run time (dlopen): a single-GPU run neither needs nor loads it.
#include <dlfcn.h>
#include <nccl.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#include "session.cuh"

namespace cyp {

// transition_scores_kernel lives in scores.cu
__global__ void transition_scores_kernel(int64_t n_rows, int32_t n_cols, const int32_t *__restrict__ seq, const RowInfo *__restrict__ rows,
                                         int align_elems, const int32_t *__restrict__ col, const double *__restrict__ val,
                                         float *__restrict__ score, unsigned long long *__restrict__ n_missing);

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string where;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return CYP_OK;
    // the copy the process already has (torch's, when torch is imported) wins; then CYP_NCCL_PATH; then the system's
    const char *env = getenv("CYP_NCCL_PATH");
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    std::string where = "libnccl.so.2 (already loaded)";
    if (!h && env && env[0]) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); where = env; }
    if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); where = "libnccl.so.2"; }
    if (!h) return fail(CYP_E_INVALID, std::string("multi-GPU groups need NCCL: libnccl.so.2 could not be loaded (") + dlerror() + "); set CYP_NCCL_PATH");
#define CYP_SYM(field, name)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));            \
    if (!g_nccl.field) return fail(CYP_E_INVALID, std::string("NCCL symbol missing: ") + name);
    CYP_SYM(GetVersion, "ncclGetVersion") CYP_SYM(GetUniqueId, "ncclGetUniqueId") CYP_SYM(CommInitRank, "ncclCommInitRank")
    CYP_SYM(CommInitAll, "ncclCommInitAll") CYP_SYM(CommDestroy, "ncclCommDestroy") CYP_SYM(AllGather, "ncclAllGather")
    CYP_SYM(Broadcast, "ncclBroadcast") CYP_SYM(AllReduce, "ncclAllReduce") CYP_SYM(GroupStart, "ncclGroupStart")
    CYP_SYM(GroupEnd, "ncclGroupEnd") CYP_SYM(GetErrorString, "ncclGetErrorString")
#undef CYP_SYM
    g_nccl.where = where;
    g_nccl.handle = h;
    return CYP_OK;
}

#define G_NCCL(expr)                                                                                          \
    do {                                                                                                      \
        ncclResult_t _r = (expr);                                                                             \
        if (_r != ncclSuccess) return fail(CYP_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)
#define G_CUDA(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? CYP_E_NOMEM : CYP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

constexpr int kThreads = 256;
inline unsigned blocks_for(int64_t n, int per_block) { return static_cast<unsigned>((n + per_block - 1) / per_block); }

__global__ void __launch_bounds__(kThreads) fill_i32_kernel(int64_t n, int32_t v, int32_t *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i < n) out[i] = v;
}
__global__ void __launch_bounds__(kThreads) iota_u32_kernel(int64_t n, uint32_t *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i < n) out[i] = static_cast<uint32_t>(i);
}
// pick k = the off[k]-th kept sequence (accumulation order) that ends on end point slot[k]: its position in the global index
__global__ void __launch_bounds__(kThreads)
locate_picks_kernel(int64_t n_pick, const int32_t *__restrict__ slot, const int64_t *__restrict__ off, int64_t n_rows,
                    const uint32_t *__restrict__ sorted_slot, const uint32_t *__restrict__ order, int64_t *__restrict__ row, int *__restrict__ bad) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (k >= n_pick) return;
    const uint32_t s = static_cast<uint32_t>(slot[k]);
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_slot[mid] < s) lo = mid + 1; else hi = mid;
    }
    const int64_t pos = lo + off[k];
    if (off[k] < 0 || pos >= n_rows || sorted_slot[pos] != s) { *bad = 1; row[k] = 0; return; }
    row[k] = order[pos];
}
// order-sensitive 128-bit digest of an index of kept sequences
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(kThreads)
digest_kernel(int64_t n, const int32_t *__restrict__ round, const int64_t *__restrict__ walker, const int32_t *__restrict__ slot,
              unsigned long long *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    uint64_t a = 0, b = 0;
    if (i < n) {
        const uint64_t x = mix(static_cast<uint64_t>(walker[i]) ^ (static_cast<uint64_t>(static_cast<uint32_t>(round[i])) << 48));
        const uint64_t y = mix(static_cast<uint64_t>(static_cast<uint32_t>(slot[i])) + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(i + 1));
        a = mix(x ^ y);
        b = mix(x + 3 * y + static_cast<uint64_t>(i));
    }
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, static_cast<unsigned long long>(a));
        atomicAdd(out + 1, static_cast<unsigned long long>(b));
    }
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace
}  // namespace cyp

using namespace cyp;

struct cyp_group {
    int world = 1, n_local = 1, first_rank = 0;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;        // [n_local]; empty when world == 1
    std::vector<cudaStream_t> streams;    // [n_local] collectives run here
    std::vector<cyp_graph *> graphs;      // [n_local], owned
    std::vector<cyp_session *> sessions;  // [n_local], owned
    // global index of kept sequences in accumulation order, replicated on every local device (world > 1; a group of
    // one device reads its session's index in place)
    struct Index { int32_t *d_round = nullptr; int64_t *d_walker = nullptr; int32_t *d_slot = nullptr; int64_t cap = 0; };
    std::vector<Index> index;
    int64_t g_rows = 0;
    int32_t n_end_slots = 0, max_steps = 0;
    int unique = 1;
    int64_t n_roots = 0;
};

namespace {

// runs fn(i) for every local device, in parallel host threads when there are several; first failure wins
template <typename Fn>
int for_each_local(cyp_group *g, Fn fn) {
    if (g->n_local == 1) return fn(0);
    std::vector<int> rc(g->n_local, CYP_OK);
    std::vector<std::string> msg(g->n_local);
    std::vector<std::thread> th;
    for (int i = 0; i < g->n_local; ++i)
        th.emplace_back([&, i]() {
            cudaSetDevice(g->devices[i]);
            rc[i] = fn(i);
            if (rc[i] != CYP_OK) msg[i] = cyp_last_error();       // the message is thread-local: carry it over
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < g->n_local; ++i)
        if (rc[i] != CYP_OK) return fail(rc[i], "device " + std::to_string(g->devices[i]) + ": " + msg[i]);
    return CYP_OK;
}

int sync_streams(cyp_group *g) {
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        G_CUDA(cudaStreamSynchronize(g->streams[i]));
    }
    return CYP_OK;
}
int sync_default(cyp_group *g) {
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        G_CUDA(cudaStreamSynchronize(nullptr));
    }
    return CYP_OK;
}

// in-place all-gather of `count` elements per rank: every local device's buffer holds world * count elements and its own
// segment has been written
int all_gather_inplace(cyp_group *g, const std::vector<void *> &bufs, size_t count, ncclDataType_t type, size_t elem) {
    G_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < g->n_local; ++i) {
        char *base = static_cast<char *>(bufs[i]);
        G_NCCL(g_nccl.AllGather(base + static_cast<size_t>(g->first_rank + i) * count * elem, base, count, type, g->comms[i], g->streams[i]));
    }
    G_NCCL(g_nccl.GroupEnd());
    return CYP_OK;
}

int group_finish_create(cyp_group *g) {
    g->streams.resize(g->n_local, nullptr);
    g->graphs.assign(g->n_local, nullptr);
    g->sessions.assign(g->n_local, nullptr);
    g->index.resize(g->n_local);
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        G_CUDA(cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking));
    }
    if (g->world > 1) {
        // NCCL connects its transports lazily, per collective and algorithm, at first use.  Do it now, before the graph and
        // the round buffers are allocated: when it happens after them K2 runs 25 % slower for the rest of the process
        // (profiles/r01_multi_gpu_probe.md).  So: every collective the group uses, in a small and a large size.
        const char *warm = getenv("CYP_GROUP_WARM");
        const bool full = !(warm && warm[0] == '0');
        size_t big = full ? (8u << 20) : 8;                   // bytes per rank of the large all-gather
        if (full && warm && atoi(warm) > 1) big = static_cast<size_t>(atoi(warm)) << 20;       // developer switch: MB per rank
        std::vector<void *> bufs(g->n_local, nullptr);
        for (int i = 0; i < g->n_local; ++i) {
            DeviceGuard guard(g->devices[i]);
            G_CUDA(cudaMalloc(&bufs[i], big * g->world));
            G_CUDA(cudaMemset(bufs[i], 0, big * g->world));
        }
        int rc = all_gather_inplace(g, bufs, 1, ncclInt64, sizeof(int64_t));
        if (rc == CYP_OK && full) rc = all_gather_inplace(g, bufs, big / 8, ncclInt64, sizeof(int64_t));
        if (rc == CYP_OK && full) rc = all_gather_inplace(g, bufs, big / 4, ncclInt32, sizeof(int32_t));
        if (rc == CYP_OK && full) {
            ncclResult_t r = g_nccl.GroupStart();
            for (int i = 0; i < g->n_local && r == ncclSuccess; ++i) r = g_nccl.Broadcast(bufs[i], bufs[i], 128, ncclChar, 0, g->comms[i], g->streams[i]);
            if (r == ncclSuccess) r = g_nccl.GroupEnd();
            if (r == ncclSuccess) r = g_nccl.GroupStart();
            for (int i = 0; i < g->n_local && r == ncclSuccess; ++i) r = g_nccl.Broadcast(bufs[i], bufs[i], big * g->world, ncclChar, 0, g->comms[i], g->streams[i]);
            if (r == ncclSuccess) r = g_nccl.GroupEnd();
            if (r != ncclSuccess) rc = fail(CYP_E_CUDA, std::string("NCCL warm-up: ") + g_nccl.GetErrorString(r));
        }
        if (rc == CYP_OK) rc = sync_streams(g);
        for (int i = 0; i < g->n_local; ++i) {
            DeviceGuard guard(g->devices[i]);
            cudaFree(bufs[i]);
        }
        if (rc != CYP_OK) return rc;
    }
    return CYP_OK;
}

}  // namespace

extern "C" int cyp_group_create_local(int n_devices, const int *devices, cyp_group **out) {
    CYP_REQUIRE(out != nullptr, "cyp_group_create_local: out is NULL");
    *out = nullptr;
    CYP_REQUIRE(n_devices >= 1 && devices != nullptr, "cyp_group_create_local: no devices");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    for (int i = 0; i < n_devices; ++i) {
        CYP_REQUIRE(devices[i] >= 0 && devices[i] < ndev, "cyp_group_create_local: no such CUDA device");
        for (int j = 0; j < i; ++j) CYP_REQUIRE(devices[j] != devices[i], "cyp_group_create_local: a device is listed twice");
    }
    cyp_group *g = new cyp_group();
    g->world = g->n_local = n_devices;
    g->first_rank = 0;
    g->devices.assign(devices, devices + n_devices);
    int rc = CYP_OK;
    if (n_devices > 1) {
        rc = load_nccl();
        if (rc == CYP_OK) {
            g->comms.resize(n_devices);
            ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), n_devices, devices);
            if (r != ncclSuccess) {
                g->comms.clear();
                rc = fail(CYP_E_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
            }
        }
    }
    if (rc == CYP_OK) rc = group_finish_create(g);
    if (rc != CYP_OK) {
        const std::string keep = cyp_last_error();
        cyp_group_destroy(g);
        return fail(rc, keep);
    }
    *out = g;
    return CYP_OK;
}

extern "C" int cyp_group_unique_id(uint8_t *id128) {
    CYP_REQUIRE(id128 != nullptr, "cyp_group_unique_id: NULL buffer");
    const int rc = load_nccl();
    if (rc != CYP_OK) return rc;
    ncclUniqueId id;
    G_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return CYP_OK;
}

extern "C" int cyp_group_create_rank(const uint8_t *id128, int rank, int world, int device, cyp_group **out) {
    CYP_REQUIRE(out != nullptr, "cyp_group_create_rank: out is NULL");
    *out = nullptr;
    CYP_REQUIRE(world >= 1 && rank >= 0 && rank < world, "cyp_group_create_rank: bad rank");
    CYP_REQUIRE(world == 1 || id128 != nullptr, "cyp_group_create_rank: NULL id");
    int ndev = 0;
    CYP_CUDA(cudaGetDeviceCount(&ndev));
    CYP_REQUIRE(device >= 0 && device < ndev, "cyp_group_create_rank: no such CUDA device");
    cyp_group *g = new cyp_group();
    g->world = world;
    g->n_local = 1;
    g->first_rank = rank;
    g->devices.assign(1, device);
    int rc = CYP_OK;
    if (world > 1) {
        rc = load_nccl();
        if (rc == CYP_OK) {
            DeviceGuard guard(device);
            ncclUniqueId id;
            memcpy(&id, id128, 128);
            g->comms.resize(1);
            ncclResult_t r = g_nccl.CommInitRank(&g->comms[0], world, id, rank);
            if (r != ncclSuccess) {
                g->comms.clear();
                rc = fail(CYP_E_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
            }
        }
    }
    if (rc == CYP_OK) rc = group_finish_create(g);
    if (rc != CYP_OK) {
        const std::string keep = cyp_last_error();
        cyp_group_destroy(g);
        return fail(rc, keep);
    }
    *out = g;
    return CYP_OK;
}

extern "C" int cyp_group_destroy(cyp_group *g) {
    if (!g) return CYP_OK;
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        if (i < static_cast<int>(g->sessions.size())) cyp_session_destroy(g->sessions[i]);
        if (i < static_cast<int>(g->graphs.size())) cyp_graph_destroy(g->graphs[i]);
        if (i < static_cast<int>(g->index.size())) {
            dev_free(g->index[i].d_round); dev_free(g->index[i].d_walker); dev_free(g->index[i].d_slot);
        }
        if (i < static_cast<int>(g->streams.size()) && g->streams[i]) cudaStreamDestroy(g->streams[i]);
        if (i < static_cast<int>(g->comms.size()) && g->comms[i]) g_nccl.CommDestroy(g->comms[i]);
    }
    delete g;
    return CYP_OK;
}

// Frees the graphs, sessions and indices but keeps the communicators: a group is expensive to form (ncclCommInit*),
// so a host that calls sampling() repeatedly keeps its group and releases the device memory between calls.
extern "C" int cyp_group_release(cyp_group *g) {
    if (!g) return CYP_OK;
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        if (g->sessions[i]) { cyp_session_destroy(g->sessions[i]); g->sessions[i] = nullptr; }
        if (g->graphs[i]) { cyp_graph_destroy(g->graphs[i]); g->graphs[i] = nullptr; }
        dev_free(g->index[i].d_round); dev_free(g->index[i].d_walker); dev_free(g->index[i].d_slot);
        g->index[i] = cyp_group::Index{};
    }
    g->g_rows = 0;
    return CYP_OK;
}

extern "C" int cyp_group_info(const cyp_group *g, int *world, int *n_local, int *first_rank, int *nccl_version) {
    CYP_REQUIRE(g != nullptr, "cyp_group_info: NULL group");
    if (world) *world = g->world;
    if (n_local) *n_local = g->n_local;
    if (first_rank) *first_rank = g->first_rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    return CYP_OK;
}

extern "C" cyp_graph *cyp_group_graph(const cyp_group *g, int local_index) {
    if (!g || local_index < 0 || local_index >= g->n_local) return nullptr;
    return g->graphs[local_index];
}
extern "C" cyp_session *cyp_group_session(const cyp_group *g, int local_index) {
    if (!g || local_index < 0 || local_index >= g->n_local) return nullptr;
    return g->sessions[local_index];
}

// ---- graph: build once, broadcast the device arrays -------------------------------------------------------
extern "C" int cyp_group_graph_create(cyp_group *g, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, const int32_t *h_col,
                                      const void *h_val, int val_is_f32, int cdf_mode, int root_rank, float *ms_build, float *ms_broadcast) {
    CYP_REQUIRE(g != nullptr, "cyp_group_graph_create: NULL group");
    CYP_REQUIRE(root_rank >= 0 && root_rank < g->world, "cyp_group_graph_create: bad root rank");
    for (int i = 0; i < g->n_local; ++i) {
        if (g->sessions[i]) { cyp_session_destroy(g->sessions[i]); g->sessions[i] = nullptr; }
        if (g->graphs[i]) { cyp_graph_destroy(g->graphs[i]); g->graphs[i] = nullptr; }
    }
    const int root_local = root_rank - g->first_rank;
    const bool have_root = root_local >= 0 && root_local < g->n_local;
    const double t0 = now_ms();
    if (have_root) {
        const int rc = val_is_f32 ? cyp_graph_create_f32(g->devices[root_local], n_cells, nnz, h_row_ptr, h_col, static_cast<const float *>(h_val), cdf_mode, &g->graphs[root_local])
                                  : cyp_graph_create(g->devices[root_local], n_cells, nnz, h_row_ptr, h_col, static_cast<const double *>(h_val), cdf_mode, &g->graphs[root_local]);
        if (rc != CYP_OK && g->world == 1) return rc;
        // with other ranks waiting in the broadcast a failure must reach them too: n = -1 in the wire header
        if (rc != CYP_OK) g->graphs[root_local] = nullptr;
    }
    const double t1 = now_ms();
    if (ms_build) *ms_build = static_cast<float>(t1 - t0);
    if (ms_broadcast) *ms_broadcast = 0.f;
    if (g->world == 1) return CYP_OK;
    // 1. the host-side fields
    GraphWire wire{};
    wire.n = -1;
    std::string root_error;
    if (have_root) {
        if (g->graphs[root_local]) wire = graph_wire(g->graphs[root_local]);
        else root_error = cyp_last_error();
    }
    std::vector<void *> hdr(g->n_local, nullptr);
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        G_CUDA(dev_alloc(&hdr[i], sizeof(GraphWire)));
        if (i == root_local) G_CUDA(cudaMemcpy(hdr[i], &wire, sizeof(GraphWire), cudaMemcpyHostToDevice));
    }
    int rc = sync_default(g);
    if (rc != CYP_OK) return rc;
    G_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < g->n_local; ++i)
        G_NCCL(g_nccl.Broadcast(hdr[i], hdr[i], sizeof(GraphWire), ncclChar, root_rank, g->comms[i], g->streams[i]));
    G_NCCL(g_nccl.GroupEnd());
    rc = sync_streams(g);
    if (rc != CYP_OK) return rc;
    {
        DeviceGuard guard(g->devices[0]);
        G_CUDA(cudaMemcpy(&wire, hdr[0], sizeof(GraphWire), cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        dev_free(hdr[i]);
    }
    if (wire.n < 0) return fail(CYP_E_INVALID, have_root ? root_error : std::string("cyp_group_graph_create: the root rank could not build the graph"));
    // 2. the device arrays
    for (int i = 0; i < g->n_local; ++i) {
        if (i == root_local) continue;
        rc = graph_alloc_like(wire, g->devices[i], &g->graphs[i]);
        if (rc != CYP_OK) return rc;
    }
    rc = sync_default(g);
    if (rc != CYP_OK) return rc;
    std::vector<std::vector<std::pair<void *, size_t>>> arrays(g->n_local);
    for (int i = 0; i < g->n_local; ++i) graph_device_arrays(g->graphs[i], &arrays[i]);
    {
        G_NCCL(g_nccl.GroupStart());
        for (size_t a = 0; a < arrays[0].size(); ++a)
            for (int i = 0; i < g->n_local; ++i)
                G_NCCL(g_nccl.Broadcast(arrays[i][a].first, arrays[i][a].first, arrays[i][a].second, ncclChar, root_rank, g->comms[i], g->streams[i]));
        G_NCCL(g_nccl.GroupEnd());
        rc = sync_streams(g);
        if (rc != CYP_OK) return rc;
    }
    if (ms_broadcast) *ms_broadcast = static_cast<float>(now_ms() - t1);
    return CYP_OK;
}

// The same graph, when EVERY rank holds the host CSR (torchrun: each process has the AnnData; one process with N devices:
// the one host copy): each rank uploads 1/world of the column and value arrays over its own PCIe link, an in-place
// ncclAllGather over NVLink completes them on every device, and every device runs K1 / K1b itself (3 ms at 1M cells).
// The upload, the longest part of building a graph, shrinks with the number of GPUs; the 1.6 GB of built arrays never travel.
extern "C" int cyp_group_graph_create_sharded(cyp_group *g, int64_t n_cells, int64_t nnz, const int64_t *h_row_ptr, const int32_t *h_col,
                                              const void *h_val, int val_is_f32, int cdf_mode, float *ms_upload, float *ms_gather,
                                              float *ms_total) {
    CYP_REQUIRE(g != nullptr, "cyp_group_graph_create_sharded: NULL group");
    CYP_REQUIRE(n_cells > 0 && nnz > 0 && h_row_ptr && h_col && h_val, "cyp_group_graph_create_sharded: every rank passes the CSR");
    if (ms_upload) *ms_upload = 0.f;
    if (ms_gather) *ms_gather = 0.f;
    if (ms_total) *ms_total = 0.f;
    const double t0 = now_ms();
    if (g->world == 1) {
        const int rc = cyp_group_graph_create(g, n_cells, nnz, h_row_ptr, h_col, h_val, val_is_f32, cdf_mode, 0, ms_total, nullptr);
        if (ms_upload && ms_total) *ms_upload = *ms_total;
        return rc;
    }
    for (int i = 0; i < g->n_local; ++i) {
        if (g->sessions[i]) { cyp_session_destroy(g->sessions[i]); g->sessions[i] = nullptr; }
        if (g->graphs[i]) { cyp_graph_destroy(g->graphs[i]); g->graphs[i] = nullptr; }
    }
    const int W = g->world;
    const size_t val_elem = val_is_f32 ? sizeof(float) : sizeof(double);
    // ---- 1. do all ranks hold the same matrix?  (a mismatch would otherwise hang the all-gather or build different graphs)
    {
        constexpr int kWords = 4;
        uint64_t h = 0x243F6A8885A308D3ull;
        auto fold = [&](uint64_t x) { h = (h ^ x) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; };
        for (int k = 0; k < 64; ++k) {
            const int64_t r = (n_cells * k) / 64, e = (nnz * k) / 64;
            fold(static_cast<uint64_t>(h_row_ptr[r]));
            fold(static_cast<uint64_t>(static_cast<uint32_t>(h_col[e])));
            uint64_t bits = 0;
            memcpy(&bits, static_cast<const char *>(h_val) + val_elem * e, val_elem);
            fold(bits);
        }
        fold(static_cast<uint64_t>(h_row_ptr[n_cells]));
        const int64_t mine[kWords] = {n_cells, nnz, static_cast<int64_t>((val_is_f32 ? 256 : 0) | cdf_mode), static_cast<int64_t>(h)};
        std::vector<void *> d_chk(g->n_local, nullptr);
        struct FreeBufs { cyp_group *g; std::vector<void *> &a; ~FreeBufs() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(a[i]); } } } free_bufs{g, d_chk};
        for (int i = 0; i < g->n_local; ++i) {
            DeviceGuard guard(g->devices[i]);
            G_CUDA(dev_alloc(&d_chk[i], sizeof(int64_t) * kWords * W));
            G_CUDA(cudaMemcpy(static_cast<int64_t *>(d_chk[i]) + static_cast<size_t>(g->first_rank + i) * kWords, mine, sizeof(mine), cudaMemcpyHostToDevice));
        }
        int rc = all_gather_inplace(g, d_chk, kWords, ncclInt64, sizeof(int64_t));
        if (rc == CYP_OK) rc = sync_streams(g);
        if (rc != CYP_OK) return rc;
        std::vector<int64_t> all(static_cast<size_t>(kWords) * W);
        {
            DeviceGuard guard(g->devices[0]);
            G_CUDA(cudaMemcpy(all.data(), d_chk[0], sizeof(int64_t) * kWords * W, cudaMemcpyDeviceToHost));
        }
        for (int r = 0; r < W; ++r)
            if (memcmp(all.data() + static_cast<size_t>(r) * kWords, all.data(), sizeof(mine)) != 0)
                return fail(CYP_E_INVALID, "cyp_group_graph_create_sharded: rank " + std::to_string(r) +
                                               " passed a different matrix (or cdf_mode) than rank 0: every rank must pass the same CSR");
    }
    // ---- 2. every rank uploads its slice; in-place all-gather over NVLink
    const int64_t per = (((nnz + W - 1) / W) + 3) & ~int64_t(3);     // elements per rank; slices stay 16-byte aligned
    std::vector<RawCsr> raw(g->n_local);
    std::vector<cudaEvent_t> ev(g->n_local * 3, nullptr);            // per device: start, uploaded, gathered
    auto drop_events = [&]() {
        for (int i = 0; i < g->n_local; ++i) {
            DeviceGuard guard(g->devices[i]);
            for (int k = 0; k < 3; ++k) if (ev[3 * i + k]) cudaEventDestroy(ev[3 * i + k]);
            if (raw[i].ready) cudaEventDestroy(raw[i].ready);
        }
    };
    auto drop_raw = [&]() {
        for (int i = 0; i < g->n_local; ++i) {
            DeviceGuard guard(g->devices[i]);
            cudaStreamSynchronize(g->streams[i]);
            dev_free(raw[i].d_col); dev_free(raw[i].d_val);
            raw[i].d_col = nullptr; raw[i].d_val = nullptr;
        }
    };
    int rc = for_each_local(g, [&](int i) -> int {
        DeviceGuard guard(g->devices[i]);
        const int r = g->first_rank + i;
        cudaStream_t st = g->streams[i];
        for (int k = 0; k < 3; ++k) G_CUDA(cudaEventCreate(&ev[3 * i + k]));
        G_CUDA(cudaEventCreateWithFlags(&raw[i].ready, cudaEventDisableTiming));
        G_CUDA(dev_alloc(&raw[i].d_col, sizeof(int32_t) * per * W));
        G_CUDA(dev_alloc(&raw[i].d_val, val_elem * per * W));
        G_CUDA(cudaEventRecord(raw[i].ready, nullptr));              // device memory comes from the default stream's pool
        G_CUDA(cudaStreamWaitEvent(st, raw[i].ready, 0));
        G_CUDA(cudaEventRecord(ev[3 * i], st));
        const int64_t a = std::min<int64_t>(per * r, nnz), b = std::min<int64_t>(a + per, nnz);
        if (b > a) {
            G_CUDA(cudaMemcpyAsync(raw[i].d_col + a, h_col + a, sizeof(int32_t) * (b - a), cudaMemcpyHostToDevice, st));
            G_CUDA(cudaMemcpyAsync(static_cast<char *>(raw[i].d_val) + val_elem * a, static_cast<const char *>(h_val) + val_elem * a,
                                   val_elem * (b - a), cudaMemcpyHostToDevice, st));
        }
        G_CUDA(cudaEventRecord(ev[3 * i + 1], st));
        return CYP_OK;
    });
    if (rc == CYP_OK) {
        std::vector<void *> cols(g->n_local), vals(g->n_local);
        for (int i = 0; i < g->n_local; ++i) { cols[i] = raw[i].d_col; vals[i] = raw[i].d_val; }
        rc = all_gather_inplace(g, cols, static_cast<size_t>(per), ncclInt32, sizeof(int32_t));
        if (rc == CYP_OK) rc = all_gather_inplace(g, vals, static_cast<size_t>(per) * (val_elem / 4), ncclInt32, sizeof(int32_t));
    }
    if (rc == CYP_OK)
        for (int i = 0; i < g->n_local && rc == CYP_OK; ++i) {
            DeviceGuard guard(g->devices[i]);
            cudaError_t e = cudaEventRecord(ev[3 * i + 2], g->streams[i]);
            if (e == cudaSuccess) e = cudaEventRecord(raw[i].ready, g->streams[i]);
            if (e != cudaSuccess) rc = fail(CYP_E_CUDA, std::string("cyp_group_graph_create_sharded: ") + cudaGetErrorString(e));
        }
    if (rc != CYP_OK) {
        const std::string keep = cyp_last_error();
        drop_raw();
        drop_events();
        return fail(rc, keep);
    }
    // ---- 3. every device builds its graph (the host lays out the rows while the slices travel)
    rc = for_each_local(g, [&](int i) -> int {
        RawCsr mine = raw[i];
        raw[i].d_col = nullptr; raw[i].d_val = nullptr;              // the build owns them now, also when it fails
        return graph_create_raw(g->devices[i], n_cells, nnz, h_row_ptr, val_is_f32 != 0, cdf_mode, mine, &g->graphs[i]);
    });
    if (rc == CYP_OK) {
        DeviceGuard guard(g->devices[0]);
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, ev[0], ev[1]) == cudaSuccess && ms_upload) *ms_upload = a;
        if (cudaEventElapsedTime(&b, ev[1], ev[2]) == cudaSuccess && ms_gather) *ms_gather = b;
    }
    const std::string keep = rc == CYP_OK ? std::string() : std::string(cyp_last_error());
    drop_raw();
    drop_events();
    if (rc != CYP_OK) {
        for (int i = 0; i < g->n_local; ++i)
            if (g->graphs[i]) { cyp_graph_destroy(g->graphs[i]); g->graphs[i] = nullptr; }
        return fail(rc, keep);
    }
    if (ms_total) *ms_total = static_cast<float>(now_ms() - t0);
    return CYP_OK;
}

extern "C" int cyp_group_session_create(cyp_group *g, const int32_t *h_root_cells, int64_t n_roots, const int32_t *h_cell_code,
                                        int32_t n_codes, int32_t min_clusters, const int32_t *h_end_slot_of_cell, int32_t n_end_slots,
                                        int32_t max_steps, int unique, uint64_t seed) {
    CYP_REQUIRE(g != nullptr, "cyp_group_session_create: NULL group");
    for (int i = 0; i < g->n_local; ++i) {
        CYP_REQUIRE(g->graphs[i] != nullptr, "cyp_group_session_create: the group has no graph");
        if (g->sessions[i]) { cyp_session_destroy(g->sessions[i]); g->sessions[i] = nullptr; }
    }
    g->g_rows = 0;
    g->n_end_slots = n_end_slots;
    g->max_steps = max_steps;
    g->unique = unique;
    g->n_roots = n_roots;
    return for_each_local(g, [&](int i) {
        return cyp_session_create(g->graphs[i], h_root_cells, n_roots, h_cell_code, n_codes, min_clusters, h_end_slot_of_cell,
                                  n_end_slots, max_steps, unique, seed, &g->sessions[i]);
    });
}

namespace {

int grow_group_index(cyp_group *g, int i, int64_t need) {
    cyp_group::Index &ix = g->index[i];
    if (need <= ix.cap) return CYP_OK;
    DeviceGuard guard(g->devices[i]);
    int64_t cap = std::max<int64_t>(std::max<int64_t>(need, 4096), ix.cap + ix.cap / 2);
    int32_t *r = nullptr, *s = nullptr;
    int64_t *w = nullptr;
    G_CUDA(dev_alloc(&r, sizeof(int32_t) * cap));
    G_CUDA(dev_alloc(&w, sizeof(int64_t) * cap));
    G_CUDA(dev_alloc(&s, sizeof(int32_t) * cap));
    if (g->g_rows) {
        G_CUDA(cudaMemcpy(r, ix.d_round, sizeof(int32_t) * g->g_rows, cudaMemcpyDeviceToDevice));
        G_CUDA(cudaMemcpy(w, ix.d_walker, sizeof(int64_t) * g->g_rows, cudaMemcpyDeviceToDevice));
        G_CUDA(cudaMemcpy(s, ix.d_slot, sizeof(int32_t) * g->g_rows, cudaMemcpyDeviceToDevice));
    }
    dev_free(ix.d_round); dev_free(ix.d_walker); dev_free(ix.d_slot);
    ix.d_round = r; ix.d_walker = w; ix.d_slot = s; ix.cap = cap;
    return CYP_OK;
}

// pointers to the global index as local device i sees it
void group_index(const cyp_group *g, int i, const int32_t **round, const int64_t **walker, const int32_t **slot) {
    if (g->world == 1) {
        *round = g->sessions[i]->d_k_round; *walker = g->sessions[i]->d_k_walker; *slot = g->sessions[i]->d_k_slot;
    } else {
        *round = g->index[i].d_round; *walker = g->index[i].d_walker; *slot = g->index[i].d_slot;
    }
}

constexpr int kStatWords = 8;     // per rank: n_records, n_walkers, n_pass, n_terminal, n_nocrossing, n_replayed, ms_walk, ms_filter (float bits)

}  // namespace

extern "C" int cyp_group_round(cyp_group *g, uint32_t round, int64_t sims_per_root, int64_t *h_slot_counts, cyp_group_round_stats *stats) {
    CYP_REQUIRE(g != nullptr && stats != nullptr, "cyp_group_round: NULL argument");
    for (int i = 0; i < g->n_local; ++i) CYP_REQUIRE(g->sessions[i] != nullptr, "cyp_group_round: the group has no session");
    CYP_REQUIRE(sims_per_root > 0, "cyp_group_round: bad sims_per_root");
    *stats = cyp_group_round_stats{};
    const int64_t total = g->n_roots * sims_per_root;
    const int W = g->world;
    const double t0 = now_ms();
    // ---- every rank walks its block of the round's walkers
    std::vector<cyp_round_stats> local(g->n_local);
    int rc = for_each_local(g, [&](int i) {
        const int r = g->first_rank + i;
        const int64_t w0 = (total * r) / W, w1 = (total * (r + 1)) / W;
        return cyp_session_round(g->sessions[i], round, sims_per_root, w0, w1, &local[i]);
    });
    const bool nocross = rc == CYP_E_NOCROSSING;
    if (rc != CYP_OK && !(nocross && W > 1)) return rc;            // a missing crossing on one rank must not leave the others in the collective
    const double t1 = now_ms();
    stats->ms_local = static_cast<float>(t1 - t0);
    int64_t row0 = g->g_rows;
    if (W == 1) {
        const cyp_round_stats &s = local[0];
        stats->n_walkers = s.n_walkers; stats->n_pass_clusters = s.n_pass_clusters; stats->n_terminal = s.n_terminal;
        stats->n_kept = s.n_kept; stats->n_nocrossing = s.n_nocrossing; stats->n_replayed = s.n_replayed;
        stats->ms_walk = s.ms_walk; stats->ms_filter = s.ms_filter; stats->n_chunks = s.n_chunks; stats->dedup_passes = s.dedup_passes;
        g->g_rows = g->sessions[0]->kept_rows;
        stats->total_rows = g->g_rows;
    } else {
        // ---- 1. counts and statistics of every rank
        std::vector<void *> d_stat(g->n_local, nullptr);
        std::vector<int64_t> h_stat(static_cast<size_t>(W) * kStatWords, 0);
        auto gather_stats = [&](bool first) -> int {
            for (int i = 0; i < g->n_local; ++i) {
                DeviceGuard guard(g->devices[i]);
                if (!d_stat[i]) G_CUDA(dev_alloc(&d_stat[i], sizeof(int64_t) * W * kStatWords));
                int64_t mine[kStatWords] = {g->sessions[i]->n_records, 0, 0, 0, 0, 0, 0, 0};
                if (first) {
                    const cyp_round_stats &s = local[i];
                    mine[1] = s.n_walkers; mine[2] = s.n_pass_clusters; mine[3] = s.n_terminal; mine[4] = s.n_nocrossing; mine[5] = s.n_replayed;
                    memcpy(&mine[6], &s.ms_walk, sizeof(float));
                    memcpy(&mine[7], &s.ms_filter, sizeof(float));
                }
                G_CUDA(cudaMemcpy(static_cast<int64_t *>(d_stat[i]) + static_cast<size_t>(g->first_rank + i) * kStatWords, mine, sizeof(mine), cudaMemcpyHostToDevice));
            }
            int rc2 = all_gather_inplace(g, d_stat, kStatWords, ncclInt64, sizeof(int64_t));
            if (rc2 != CYP_OK) return rc2;
            rc2 = sync_streams(g);
            if (rc2 != CYP_OK) return rc2;
            DeviceGuard guard(g->devices[0]);
            G_CUDA(cudaMemcpy(h_stat.data(), d_stat[0], sizeof(int64_t) * W * kStatWords, cudaMemcpyDeviceToHost));
            return CYP_OK;
        };
        struct FreeStat { cyp_group *g; std::vector<void *> &p; ~FreeStat() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(p[i]); } } } free_stat{g, d_stat};
        rc = gather_stats(true);
        if (rc != CYP_OK) return rc;
        stats->exchange_bytes += static_cast<int64_t>(sizeof(int64_t)) * W * kStatWords;
        for (int r = 0; r < W; ++r) {
            const int64_t *s = &h_stat[static_cast<size_t>(r) * kStatWords];
            stats->n_walkers += s[1]; stats->n_pass_clusters += s[2]; stats->n_terminal += s[3]; stats->n_nocrossing += s[4]; stats->n_replayed += s[5];
            float a, b;
            memcpy(&a, &s[6], sizeof(float)); memcpy(&b, &s[7], sizeof(float));
            stats->ms_walk = std::max(stats->ms_walk, a); stats->ms_filter = std::max(stats->ms_filter, b);
        }
        for (int i = 0; i < g->n_local; ++i) { stats->n_chunks = std::max(stats->n_chunks, local[i].n_chunks); stats->dedup_passes = std::max(stats->dedup_passes, local[i].dedup_passes); }
        if (stats->n_nocrossing)
            return fail(CYP_E_NOCROSSING, "cyp_group_round: a uniform exceeded its row's last CDF value (reference: IndexError at markov_chain_sampler.py:49)");
        // ---- 2. first occurrence across ranks: all-gather the records, every rank prunes its own
        int64_t n_all = 0, n_max = 0;
        for (int r = 0; r < W; ++r) { n_all += h_stat[static_cast<size_t>(r) * kStatWords]; n_max = std::max(n_max, h_stat[static_cast<size_t>(r) * kStatWords]); }
        if (g->unique && n_all > 0) {
            std::vector<void *> d_pad(g->n_local, nullptr), d_all(g->n_local, nullptr);
            struct FreeBufs { cyp_group *g; std::vector<void *> &a, &b; ~FreeBufs() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(a[i]); dev_free(b[i]); } } } free_bufs{g, d_pad, d_all};
            const size_t seg = static_cast<size_t>(n_max) * 3;
            for (int i = 0; i < g->n_local; ++i) {
                DeviceGuard guard(g->devices[i]);
                G_CUDA(dev_alloc(&d_pad[i], sizeof(uint64_t) * seg * W));
                G_CUDA(dev_alloc(&d_all[i], sizeof(uint64_t) * 3 * n_all));
                const cyp_session *s = g->sessions[i];
                if (s->n_records)
                    G_CUDA(cudaMemcpyAsync(static_cast<uint64_t *>(d_pad[i]) + seg * (g->first_rank + i), s->d_records, sizeof(uint64_t) * 3 * s->n_records,
                                           cudaMemcpyDeviceToDevice, nullptr));
            }
            rc = sync_default(g);
            if (rc != CYP_OK) return rc;
            rc = all_gather_inplace(g, d_pad, seg, ncclUint64, sizeof(uint64_t));
            if (rc != CYP_OK) return rc;
            stats->exchange_bytes += static_cast<int64_t>(sizeof(uint64_t) * seg * W);
            for (int i = 0; i < g->n_local; ++i) {               // padded segments -> one list, rank order = walker order
                DeviceGuard guard(g->devices[i]);
                int64_t at = 0;
                for (int r = 0; r < W; ++r) {
                    const int64_t c = h_stat[static_cast<size_t>(r) * kStatWords];
                    if (c) G_CUDA(cudaMemcpyAsync(static_cast<uint64_t *>(d_all[i]) + 3 * at, static_cast<uint64_t *>(d_pad[i]) + seg * r, sizeof(uint64_t) * 3 * c,
                                                  cudaMemcpyDeviceToDevice, g->streams[i]));
                    at += c;
                }
            }
            rc = sync_streams(g);
            if (rc != CYP_OK) return rc;
            std::vector<int64_t> dropped(g->n_local, 0);
            rc = for_each_local(g, [&](int i) {
                return cyp_session_round_prune(g->sessions[i], static_cast<const uint64_t *>(d_all[i]), n_all, &dropped[i]);
            });
            if (rc != CYP_OK) return rc;
            rc = gather_stats(false);                            // the counts after the prune
            if (rc != CYP_OK) return rc;
        }
        // ---- 3. the survivors' (walker, slot) of every rank -> the global index, in rank order
        int64_t n_kept = 0, k_max = 0;
        for (int r = 0; r < W; ++r) { n_kept += h_stat[static_cast<size_t>(r) * kStatWords]; k_max = std::max(k_max, h_stat[static_cast<size_t>(r) * kStatWords]); }
        stats->n_kept = n_kept;
        if (n_kept > 0) {
            std::vector<void *> d_w(g->n_local, nullptr), d_s(g->n_local, nullptr);
            struct FreeBufs { cyp_group *g; std::vector<void *> &a, &b; ~FreeBufs() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(a[i]); dev_free(b[i]); } } } free_bufs{g, d_w, d_s};
            for (int i = 0; i < g->n_local; ++i) {
                rc = grow_group_index(g, i, g->g_rows + n_kept);
                if (rc != CYP_OK) return rc;
                DeviceGuard guard(g->devices[i]);
                G_CUDA(dev_alloc(&d_w[i], sizeof(int64_t) * k_max * W));
                G_CUDA(dev_alloc(&d_s[i], sizeof(int32_t) * k_max * W));
                const cyp_session *s = g->sessions[i];
                const int64_t m = s->n_records;
                if (m) {
                    G_CUDA(cudaMemcpyAsync(static_cast<int64_t *>(d_w[i]) + k_max * (g->first_rank + i), s->d_k_walker + s->last_row0, sizeof(int64_t) * m, cudaMemcpyDeviceToDevice, nullptr));
                    G_CUDA(cudaMemcpyAsync(static_cast<int32_t *>(d_s[i]) + k_max * (g->first_rank + i), s->d_k_slot + s->last_row0, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, nullptr));
                }
            }
            rc = sync_default(g);
            if (rc != CYP_OK) return rc;
            rc = all_gather_inplace(g, d_w, static_cast<size_t>(k_max), ncclInt64, sizeof(int64_t));
            if (rc != CYP_OK) return rc;
            rc = all_gather_inplace(g, d_s, static_cast<size_t>(k_max), ncclInt32, sizeof(int32_t));
            if (rc != CYP_OK) return rc;
            stats->exchange_bytes += static_cast<int64_t>(12 * k_max * W);
            for (int i = 0; i < g->n_local; ++i) {
                DeviceGuard guard(g->devices[i]);
                cyp_group::Index &ix = g->index[i];
                int64_t at = g->g_rows;
                for (int r = 0; r < W; ++r) {
                    const int64_t c = h_stat[static_cast<size_t>(r) * kStatWords];
                    if (c) {
                        G_CUDA(cudaMemcpyAsync(ix.d_walker + at, static_cast<int64_t *>(d_w[i]) + k_max * r, sizeof(int64_t) * c, cudaMemcpyDeviceToDevice, g->streams[i]));
                        G_CUDA(cudaMemcpyAsync(ix.d_slot + at, static_cast<int32_t *>(d_s[i]) + k_max * r, sizeof(int32_t) * c, cudaMemcpyDeviceToDevice, g->streams[i]));
                    }
                    at += c;
                }
                fill_i32_kernel<<<blocks_for(n_kept, kThreads), kThreads, 0, g->streams[i]>>>(n_kept, static_cast<int32_t>(round), ix.d_round + g->g_rows);
                G_CUDA(cudaGetLastError());
            }
            rc = sync_streams(g);
            if (rc != CYP_OK) return rc;
        }
        g->g_rows += n_kept;
        stats->total_rows = g->g_rows;
    }
    stats->ms_exchange = static_cast<float>(now_ms() - t1);
    if (h_slot_counts) {
        const int32_t *r_, *s_;
        const int64_t *w_;
        group_index(g, 0, &r_, &w_, &s_);
        DeviceGuard guard(g->devices[0]);
        rc = slot_histogram(s_ + row0, g->g_rows - row0, g->n_end_slots, h_slot_counts, nullptr);
        if (rc != CYP_OK) return rc;
    }
    if (nocross) return fail(CYP_E_NOCROSSING, "cyp_group_round: a uniform exceeded its row's last CDF value (reference: IndexError at markov_chain_sampler.py:49)");
    return CYP_OK;
}

extern "C" int cyp_group_rows(const cyp_group *g, int64_t *n_rows) {
    CYP_REQUIRE(g != nullptr && n_rows != nullptr, "cyp_group_rows: NULL argument");
    *n_rows = g->g_rows;
    return CYP_OK;
}

extern "C" int cyp_group_digest(const cyp_group *g, uint64_t *digest2) {
    CYP_REQUIRE(g != nullptr && digest2 != nullptr, "cyp_group_digest: NULL argument");
    digest2[0] = digest2[1] = 0;
    if (g->g_rows == 0) return CYP_OK;
    const int32_t *r_, *s_;
    const int64_t *w_;
    group_index(g, 0, &r_, &w_, &s_);
    DeviceGuard guard(g->devices[0]);
    unsigned long long *d = nullptr;
    G_CUDA(dev_alloc(&d, 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), nullptr);
    if (e == cudaSuccess) {
        digest_kernel<<<blocks_for(g->g_rows, kThreads), kThreads>>>(g->g_rows, r_, w_, s_, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(digest2, d, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    dev_free(d);
    if (e != cudaSuccess) return fail(CYP_E_CUDA, std::string("cyp_group_digest: ") + cudaGetErrorString(e));
    digest2[0] ^= static_cast<uint64_t>(g->g_rows);
    return CYP_OK;
}

extern "C" int cyp_group_fetch_index(const cyp_group *g, int64_t row_begin, int64_t n_rows, int32_t *h_round, int64_t *h_walker, int32_t *h_slot) {
    CYP_REQUIRE(g != nullptr, "cyp_group_fetch_index: NULL group");
    CYP_REQUIRE(row_begin >= 0 && n_rows >= 0 && row_begin + n_rows <= g->g_rows, "cyp_group_fetch_index: bad row range");
    if (n_rows == 0) return CYP_OK;
    const int32_t *r_, *s_;
    const int64_t *w_;
    group_index(g, 0, &r_, &w_, &s_);
    DeviceGuard guard(g->devices[0]);
    if (h_round) G_CUDA(cudaMemcpy(h_round, r_ + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_walker) G_CUDA(cudaMemcpy(h_walker, w_ + row_begin, sizeof(int64_t) * n_rows, cudaMemcpyDeviceToHost));
    if (h_slot) G_CUDA(cudaMemcpy(h_slot, s_ + row_begin, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost));
    return CYP_OK;
}

// ---- final selection (:378-403): picks -> rows of the global index -> replay, sharded over the ranks -> all-gather ----
extern "C" int cyp_group_select(cyp_group *g, int64_t n_pick, const int32_t *h_slot, const int64_t *h_offset, int32_t *h_seq, float *h_score,
                                cyp_group_select_stats *stats) {
    CYP_REQUIRE(g != nullptr && n_pick >= 0, "cyp_group_select: bad argument");
    CYP_REQUIRE(n_pick == 0 || (h_slot && h_offset && h_seq), "cyp_group_select: NULL argument");
    if (stats) *stats = cyp_group_select_stats{};
    if (n_pick == 0) return CYP_OK;
    for (int64_t k = 0; k < n_pick; ++k) CYP_REQUIRE(h_slot[k] >= 0 && h_slot[k] < g->n_end_slots, "cyp_group_select: end-point slot out of range");
    CYP_REQUIRE(g->g_rows > 0 && g->g_rows < (1ll << 32), "cyp_group_select: no kept sequences (or more than 2^32)");
    const int W = g->world;
    const int32_t ncol = g->max_steps;
    const int64_t share = (n_pick + W - 1) / W;                   // picks per rank (the last ranks may have fewer)
    const double t0 = now_ms();
    std::vector<int32_t *> d_rows(g->n_local, nullptr);
    std::vector<float *> d_score(g->n_local, nullptr);
    struct FreeRows { cyp_group *g; std::vector<int32_t *> &a; std::vector<float *> &b; ~FreeRows() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(a[i]); dev_free(b[i]); } } } free_rows{g, d_rows, d_score};
    int rc = for_each_local(g, [&](int i) -> int {
        DeviceGuard guard(g->devices[i]);
        cudaStream_t st = nullptr;
        const cyp_session *s = g->sessions[i];
        const int32_t *r_, *s_;
        const int64_t *w_;
        group_index(g, i, &r_, &w_, &s_);
        const int64_t n = g->g_rows;
        // stable sort of the index positions by end-point slot: inside a slot the accumulation order survives (:386-394)
        uint32_t *d_key = nullptr, *d_key_sorted = nullptr, *d_val = nullptr, *d_order = nullptr;
        int32_t *d_pslot = nullptr;
        int64_t *d_poff = nullptr, *d_prow = nullptr;
        int *d_bad = nullptr;
        void *d_tmp = nullptr;
        ReplayItem *d_items = nullptr;
        std::vector<void *> scratch;
        auto alloc = [&](void **p, size_t bytes) { cudaError_t e = dev_alloc(p, bytes); if (e == cudaSuccess) scratch.push_back(*p); return e; };
        struct FreeScratch { std::vector<void *> &v; ~FreeScratch() { for (void *p : v) dev_free(p); } } free_scratch{scratch};
        G_CUDA(alloc(reinterpret_cast<void **>(&d_key_sorted), sizeof(uint32_t) * n));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_val), sizeof(uint32_t) * n));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_order), sizeof(uint32_t) * n));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_pslot), sizeof(int32_t) * n_pick));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_poff), sizeof(int64_t) * n_pick));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_prow), sizeof(int64_t) * n_pick));
        G_CUDA(alloc(reinterpret_cast<void **>(&d_bad), sizeof(int)));
        G_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
        d_key = reinterpret_cast<uint32_t *>(const_cast<int32_t *>(s_));                 // slots are >= 0: sorted as unsigned
        int bits = 1;
        while ((1ll << bits) < g->n_end_slots) ++bits;
        size_t tmp_bytes = 0;
        G_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key_sorted, d_val, d_order, static_cast<int>(n), 0, bits, st));
        G_CUDA(alloc(&d_tmp, tmp_bytes));
        iota_u32_kernel<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, d_val);
        G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key_sorted, d_val, d_order, static_cast<int>(n), 0, bits, st));
        G_CUDA(cudaMemcpyAsync(d_pslot, h_slot, sizeof(int32_t) * n_pick, cudaMemcpyHostToDevice, st));
        G_CUDA(cudaMemcpyAsync(d_poff, h_offset, sizeof(int64_t) * n_pick, cudaMemcpyHostToDevice, st));
        locate_picks_kernel<<<blocks_for(n_pick, kThreads), kThreads, 0, st>>>(n_pick, d_pslot, d_poff, n, d_key_sorted, d_order, d_prow, d_bad);
        G_CUDA(cudaGetLastError());
        int bad = 0;
        G_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        G_CUDA(cudaStreamSynchronize(st));
        if (bad) return fail(CYP_E_INVALID, "cyp_group_select: a pick lies beyond the kept sequences of its end point");
        // this rank's share of the picks, replayed into its segment of the row buffer
        G_CUDA(dev_alloc(&d_rows[i], sizeof(int32_t) * static_cast<size_t>(share) * W * ncol));
        const int r = g->first_rank + i;
        const int64_t p0 = std::min<int64_t>(n_pick, share * r), p1 = std::min<int64_t>(n_pick, share * (r + 1));
        if (p1 > p0) {
            const int64_t bm = std::min<int64_t>(p1 - p0, session_replay_batch_rows(s));
            G_CUDA(alloc(reinterpret_cast<void **>(&d_items), sizeof(ReplayItem) * bm));
            for (int64_t b0 = p0; b0 < p1; b0 += bm) {
                const int64_t m = std::min(bm, p1 - b0);
                launch_make_items(m, d_prow + b0, reinterpret_cast<const uint64_t *>(w_), 1, r_, 0u, s, d_items, st);
                G_CUDA(cudaGetLastError());
                const int rc2 = session_replay(s, m, d_items, d_rows[i] + b0 * ncol, st);
                if (rc2 != CYP_OK) return rc2;
            }
        }
        G_CUDA(cudaStreamSynchronize(st));
        return CYP_OK;
    });
    if (rc != CYP_OK) return rc;
    const double t1 = now_ms();
    if (W > 1) {
        std::vector<void *> bufs(g->n_local);
        for (int i = 0; i < g->n_local; ++i) bufs[i] = d_rows[i];
        rc = all_gather_inplace(g, bufs, static_cast<size_t>(share) * ncol, ncclInt32, sizeof(int32_t));
        if (rc != CYP_OK) return rc;
        rc = sync_streams(g);
        if (rc != CYP_OK) return rc;
        if (stats) stats->gather_bytes = static_cast<int64_t>(sizeof(int32_t)) * share * W * ncol;
    }
    const double t2 = now_ms();
    {   // scores (:50,:52) of the selected rows and the copy to the host, from the first local device
        DeviceGuard guard(g->devices[0]);
        const cyp_graph *gr = g->graphs[0];
        if (h_score) {
            unsigned long long *d_missing = nullptr;
            G_CUDA(dev_alloc(&d_score[0], sizeof(float) * n_pick));
            G_CUDA(dev_alloc(&d_missing, sizeof(unsigned long long)));
            cudaMemsetAsync(d_missing, 0, sizeof(unsigned long long), nullptr);
            transition_scores_kernel<<<blocks_for(n_pick * 32, 256), 256>>>(n_pick, ncol, d_rows[0], gr->d_row, gr->align_elems, gr->d_col, gr->d_val, d_score[0], d_missing);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpy(h_score, d_score[0], sizeof(float) * n_pick, cudaMemcpyDeviceToHost);
            dev_free(d_missing);
            if (e != cudaSuccess) return fail(CYP_E_CUDA, std::string("cyp_group_select: ") + cudaGetErrorString(e));
        }
        G_CUDA(cudaMemcpy(h_seq, d_rows[0], sizeof(int32_t) * static_cast<size_t>(n_pick) * ncol, cudaMemcpyDeviceToHost));
    }
    if (stats) {
        stats->ms_locate_replay = static_cast<float>(t1 - t0);
        stats->ms_gather = static_cast<float>(t2 - t1);
        stats->ms_fetch = static_cast<float>(now_ms() - t2);
    }
    return CYP_OK;
}

extern "C" int cyp_group_set_chunk_walkers(cyp_group *g, int64_t walkers) {
    CYP_REQUIRE(g != nullptr, "cyp_group_set_chunk_walkers: NULL group");
    for (int i = 0; i < g->n_local; ++i) {
        CYP_REQUIRE(g->sessions[i] != nullptr, "cyp_group_set_chunk_walkers: the group has no session");
        const int rc = cyp_session_set_chunk_walkers(g->sessions[i], walkers);
        if (rc != CYP_OK) return rc;
    }
    return CYP_OK;
}

extern "C" int cyp_group_broadcast_bytes(cyp_group *g, void *h_buf, int64_t n_bytes, int root_rank) {
    CYP_REQUIRE(g != nullptr && n_bytes >= 0 && (n_bytes == 0 || h_buf != nullptr), "cyp_group_broadcast_bytes: bad argument");
    CYP_REQUIRE(root_rank >= 0 && root_rank < g->world, "cyp_group_broadcast_bytes: bad root rank");
    if (g->world == 1 || n_bytes == 0) return CYP_OK;
    // in a single-process group every rank shares h_buf: only the device round trip of rank-per-process groups moves data
    const int root_local = root_rank - g->first_rank;
    std::vector<void *> d(g->n_local, nullptr);
    struct FreeBufs { cyp_group *g; std::vector<void *> &a; ~FreeBufs() { for (int i = 0; i < g->n_local; ++i) { DeviceGuard guard(g->devices[i]); dev_free(a[i]); } } } free_bufs{g, d};
    for (int i = 0; i < g->n_local; ++i) {
        DeviceGuard guard(g->devices[i]);
        G_CUDA(dev_alloc(&d[i], static_cast<size_t>(n_bytes)));
        if (i == root_local) G_CUDA(cudaMemcpy(d[i], h_buf, static_cast<size_t>(n_bytes), cudaMemcpyHostToDevice));
    }
    int rc = sync_default(g);
    if (rc != CYP_OK) return rc;
    G_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < g->n_local; ++i)
        G_NCCL(g_nccl.Broadcast(d[i], d[i], static_cast<size_t>(n_bytes), ncclChar, root_rank, g->comms[i], g->streams[i]));
    G_NCCL(g_nccl.GroupEnd());
    rc = sync_streams(g);
    if (rc != CYP_OK) return rc;
    if (root_local < 0 || root_local >= g->n_local) {
        DeviceGuard guard(g->devices[0]);
        G_CUDA(cudaMemcpy(h_buf, d[0], static_cast<size_t>(n_bytes), cudaMemcpyDeviceToHost));
    }
    return CYP_OK;
}
