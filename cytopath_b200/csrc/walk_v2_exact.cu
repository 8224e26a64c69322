// K2, fused-row flavour, exact (float64-semantics) CDF codes: instantiations of walk_v2.cuh.
#include "walk_v2.cuh"

namespace cyp {
int launch_walk_v2_exact(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    return launch_v2_cpr<v2::kModeExact>(p, mode, c, sm_count, stream);
}
}  // namespace cyp
