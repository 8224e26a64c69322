// K2, fused-row flavour: host-side dispatch (the kernel and its launch templates are in walk_v2.cuh; the two CDF
// modes are instantiated in walk_v2_exact.cu / walk_v2_ref.cu).
#include "walk_v2.cuh"

namespace cyp {

int launch_walk_v2(const cyp_graph *g, const WalkParams &p_in, WalkMode mode, int variant, cudaStream_t stream) {
    const int sm_count = current_sm_count();
    V2Config c;
    if (sm_count <= 0 || !v2_config(g, variant, p_in.n_walkers, sm_count, &c)) return kNotApplicable;
    WalkParams p = p_in;
    p.rows2 = g->d_rows2;
    p.edge4 = g->d_edge4;
    p.word = g->d_word;
    p.l1_rows = c.l1_rows;
    return g->cdf_mode == CYP_CDF_EXACT ? launch_walk_v2_exact(p, mode, c, sm_count, stream)
                                        : launch_walk_v2_ref(p, mode, c, sm_count, stream);
}

const char *walk_v2_name(const cyp_graph *g, int variant) {
    static thread_local char buf[160];
    V2Config c;
    if (!v2_config(g, variant, 0, 0, &c)) return nullptr;
    static const char *flavor[] = {"uniform rows", "ragged rows", "ragged rows + long rows in HBM"};
    snprintf(buf, sizeof(buf), "walk_v2_kernel<%s,fused rows,threads<=%d%s,slot=%dB rotated,%s%s>", g->cdf_mode == CYP_CDF_EXACT ? "q12/f64" : "q12/f32",
             c.threads, c.ctas > 1 ? " x 2 CTAs/SM" : "", 16 * c.cpr, flavor[c.flavor], c.l1_rows ? ",L1 rows" : "");
    return buf;
}

}  // namespace cyp
