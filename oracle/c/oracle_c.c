/*
 * oracle_c.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the two pieces of native third-party code the reference's
 * NMF inpainting path executes and whose exact operation order matters for parity:
 *
 *   1. sklearn's coordinate-descent sweep  `_update_cdnmf_fast`
 *      ($SP/sklearn/decomposition/_cdnmf_fast.pyx:8-38), float32 specialisation,
 *      called from `_update_coordinate_descent` ($SP/sklearn/decomposition/_nmf.py:369-396),
 *      reached from `NMF(...).fit_transform` at main4_NMF_gap.py:62-63,
 *      main4_NMF_mask.py:67-68, main4_NMF.py:83,87.
 *   2. numpy's legacy `RandomState(seed).standard_normal` (MT19937 `init_genrand`,
 *      53-bit doubles, polar Box-Muller with cached second deviate) used by
 *      `_initialize_nmf(init='random')` ($SP/sklearn/decomposition/_nmf.py:296-307).
 *
 * Built with -ffp-contract=off so that `grad += HHt[t,r]*W[i,r]` is a separate
 * multiply and add, as in the x86-64 wheel.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* ---- 1. coordinate-descent sweep (float32) -------------------------------------- */
/* W: n_rows x K row-major (updated in place); G: K x K; B: n_rows x K. Returns the
 * float32 violation accumulated in the reference's order (coordinate-major, row-minor). */
float oracle_cd_sweep_f32(float *W, const float *G, const float *B, ptrdiff_t n_rows,
                          ptrdiff_t K)
{
    float violation = 0.0f;
    for (ptrdiff_t t = 0; t < K; ++t) {
        for (ptrdiff_t i = 0; i < n_rows; ++i) {
            float grad = -B[i * K + t];
            for (ptrdiff_t r = 0; r < K; ++r)
                grad += G[t * K + r] * W[i * K + r];
            float pg = (W[i * K + t] == 0.0f) ? (grad < 0.0f ? grad : 0.0f) : grad;
            violation += fabsf(pg);
            float hess = G[t * K + t];
            if (hess != 0.0f) {
                float v = W[i * K + t] - grad / hess;
                W[i * K + t] = v > 0.0f ? v : 0.0f;
            }
        }
    }
    return violation;
}

/* ---- 2. MT19937 + legacy gaussian ------------------------------------------------ */
typedef struct {
    uint32_t mt[624];
    int pos;
    int has_gauss;
    double gauss;
} oracle_rng;

void oracle_rng_seed(oracle_rng *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->pos = 624;
    s->has_gauss = 0;
    s->gauss = 0.0;
}

static void mt_refill(oracle_rng *s)
{
    uint32_t *mt = s->mt;
    int k;
    uint32_t y;
    for (k = 0; k < 624 - 397; ++k) {
        y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; k < 623; ++k) {
        y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->pos = 0;
}

static uint32_t mt_next32(oracle_rng *s)
{
    if (s->pos == 624) mt_refill(s);
    uint32_t y = s->mt[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_next_double(oracle_rng *s)
{
    uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

double oracle_rng_gauss(oracle_rng *s)
{
    if (s->has_gauss) {
        s->has_gauss = 0;
        double t = s->gauss;
        s->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * mt_next_double(s) - 1.0;
        x2 = 2.0 * mt_next_double(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    s->gauss = f * x1;
    s->has_gauss = 1;
    return f * x2;
}

/* n float64 normals in draw order. */
void oracle_standard_normal(uint32_t seed, double *out, size_t n)
{
    oracle_rng s;
    oracle_rng_seed(&s, seed);
    for (size_t i = 0; i < n; ++i) out[i] = oracle_rng_gauss(&s);
}
