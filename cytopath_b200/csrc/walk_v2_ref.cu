// K2, fused-row flavour, reference (float32, 3-decimal) CDF codes: instantiations of walk_v2.cuh.
#include "walk_v2.cuh"

namespace cyp {
int launch_walk_v2_ref(const WalkParams &p, WalkMode mode, const V2Config &c, int sm_count, cudaStream_t stream) {
    return launch_v2_cpr<v2::kModeRef>(p, mode, c, sm_count, stream);
}
}  // namespace cyp
