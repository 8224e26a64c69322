/* CPU oracle for the cytopath sampler hot path in plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, on CSR and without densifying:
 *   - matrix_prep            /root/reference/cytopath/markov_chain_sampler.py:11-31
 *   - markov_sim's selection /root/reference/cytopath/markov_chain_sampler.py:45-49
 * plus the frozen Philox4x32-10 uniform stream documented in oracle/philox.py.
 * Pinned against the real reference through tests/golden (see oracle/__init__.py).
 * The product (cytopath_b200) never links or loads this file.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/liboracle.so oracle/walk.c -lm
 * (no -ffast-math: the float32 round-to-3-decimals must be IEEE exact).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* u(seed, round, walker, t) of the frozen stream (oracle/philox.py). */
double oracle_uniform(uint64_t seed, uint32_t round, uint64_t walker, uint32_t t)
{
    uint32_t ctr[4] = { (uint32_t)walker, (uint32_t)(walker >> 32), t >> 1, round };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t x[4];
    oracle_philox4x32_10(ctr, key, x);
    uint32_t a = (t & 1u) ? x[2] : x[0];
    uint32_t b = (t & 1u) ? x[3] : x[1];
    uint64_t mant = ((uint64_t)(a >> 5) << 26) + (uint64_t)(b >> 6);
    return (double)mant * (1.0 / 9007199254740992.0);
}

/* matrix_prep :18-19 (+ :23,:28,:31 in reference mode).
 * mode 0: out32[i] = round3(float32(cumsum_f64)), mode 1: out64[i] = cumsum_f64. */
void oracle_build_cdf(int64_t n, const int64_t *indptr, const double *data, int mode,
                      float *out32, double *out64)
{
    for (int64_t r = 0; r < n; ++r) {
        double acc = 0.0;
        for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) {
            acc += data[p];                       /* strictly left to right, like np.cumsum */
            if (mode == 1) {
                out64[p] = acc;
            } else {
                volatile float c = (float)acc;    /* :23/:28 float32 store */
                volatile float m = c * 1000.0f;   /* np.round(decimals=3) == rint(x*1000)/1000 in float32 */
                out32[p] = rintf(m) / 1000.0f;
            }
        }
    }
}

/* markov_sim :45-49 for walkers [walker_begin, walker_begin+n_walkers) of one round.
 * start_cells[w] is the root of local walker w.  uniforms, if not NULL, is
 * [n_walkers, n_transitions] float64 and replaces the Philox stream.
 * out_seq is int32 [n_walkers, n_transitions+1].  Returns the number of walker-steps
 * for which no bin satisfied u <= cdf (the reference raises IndexError there, :49);
 * such walkers stay on their current cell. */
int64_t oracle_walk(const int64_t *indptr, const int32_t *indices,
                    const float *cdf32, const double *cdf64,
                    const int32_t *start_cells, int64_t n_walkers, int64_t walker_begin,
                    int32_t n_transitions, uint64_t seed, uint32_t round,
                    const double *uniforms, int32_t *out_seq, int nthreads)
{
    int64_t misses = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : misses)
    for (int64_t w = 0; w < n_walkers; ++w) {
        int32_t *row = out_seq + w * (int64_t)(n_transitions + 1);
        int64_t cur = start_cells[w];
        row[0] = (int32_t)cur;
        for (int32_t t = 0; t < n_transitions; ++t) {
            double u = uniforms ? uniforms[w * (int64_t)n_transitions + t]
                                : oracle_uniform(seed, round, (uint64_t)(walker_begin + w), (uint32_t)t);
            int64_t a = indptr[cur], b = indptr[cur + 1], p;
            if (cdf64) { for (p = a; p < b; ++p) if (u <= cdf64[p]) break; }
            else       { for (p = a; p < b; ++p) if (u <= (double)cdf32[p]) break; }
            if (p == b) { ++misses; }
            else        { cur = indices[p]; }
            row[t + 1] = (int32_t)cur;
        }
    }
    return misses;
}
