// rng.cu -- numpy.random.RandomState(seed).standard_normal on the device: the initial factors sklearn draws for
// init='random' ($SP/sklearn/decomposition/_nmf.py:296-307 behind main4_NMF_gap.py:62) are
//     H0 = |avg * rng.standard_normal((K, T))|,  then  W0 = |avg * rng.standard_normal((F, K))|,   rng = RandomState(seed),
// i.e. MT19937 seeded by init_genrand, 53-bit doubles from word pairs, and the legacy polar Box-Muller with its cached
// second deviate (numpy/random/src/legacy/legacy-distributions.c: legacy_gauss).  Round 1 drew them on the host (20 ns per
// normal: 1 s for the 1-hour signal, and every rank of a time-sharded run uploaded the whole 159 MB table).  Here one CTA
// regenerates the generator state 624 words at a time -- three data-parallel sub-steps per refill, because word k of the
// new state depends on new word k - 227 only -- and turns every refill into 156 polar attempts in parallel: an attempt
// always consumes exactly four words whether it is accepted or not, so attempt a of refill r owns words 4a..4a+3, and the
// position of its two deviates in the output is a prefix count of the accepted attempts.  A rank of the sharded mode walks
// the same stream and keeps only its frames.
#include "kernels.h"

#ifndef AINMF_EMU          // the emulator build keeps the host generator (api.cu)
namespace ainmf {

constexpr int kRngThreads = 256;
constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t next, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (next & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// out index i of the stream: i < K*T_total -> H0 normal (k = i / T_total, t = i % T_total), kept when t_begin <= t < t_begin +
// t_count at Hn[k * t_count + t - t_begin]; then F*K normals of W0 at Wn[i - K*T_total].
__global__ void __launch_bounds__(kRngThreads)
numpy_normals_kernel(uint32_t seed, long long n_h, long long T_total, long long t_begin, long long t_count, float* __restrict__ Hn,
                     long long n_w, float* __restrict__ Wn) {
    __shared__ uint32_t s_mt[2][kMtN];
    __shared__ int s_cnt[8];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint32_t v = seed;
        s_mt[0][0] = v;
        for (int i = 1; i < kMtN; ++i) { v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i; s_mt[0][i] = v; }   // init_genrand
        s_base = 0;
    }
    __syncthreads();
    const long long n_total = n_h + n_w;
    int cur = 0;
    long long base = 0;
    while (base < n_total) {
        const uint32_t* o = s_mt[cur];
        uint32_t* nw = s_mt[cur ^ 1];
        // ---- refill: new[k] = old[k+397 mod 624 -> new[k-227] for k >= 227] ^ twist(old[k], old[k+1]) ----
        for (int k = tid; k < kMtN - kMtM; k += kRngThreads) nw[k] = mt_twist(o[k], o[k + 1], o[k + kMtM]);
        __syncthreads();
        for (int k = (kMtN - kMtM) + tid; k < 2 * (kMtN - kMtM); k += kRngThreads) nw[k] = mt_twist(o[k], o[k + 1], nw[k - (kMtN - kMtM)]);
        __syncthreads();
        for (int k = 2 * (kMtN - kMtM) + tid; k < kMtN; k += kRngThreads)
            nw[k] = mt_twist(o[k], (k + 1 < kMtN) ? o[k + 1] : nw[0], nw[k - (kMtN - kMtM)]);
        __syncthreads();
        cur ^= 1;
        // ---- 156 polar attempts, four tempered words each ----
        bool acc = false;
        double g0 = 0.0, g1 = 0.0;
        if (tid < kMtN / 4) {
            const uint32_t w0 = mt_temper(nw[4 * tid]), w1 = mt_temper(nw[4 * tid + 1]);
            const uint32_t w2 = mt_temper(nw[4 * tid + 2]), w3 = mt_temper(nw[4 * tid + 3]);
            const double d1 = ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
            const double d2 = ((double)(w2 >> 5) * 67108864.0 + (double)(w3 >> 6)) / 9007199254740992.0;
            const double x1 = 2.0 * d1 - 1.0, x2 = 2.0 * d2 - 1.0;
            const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));    // two roundings, as the C code compiled without FMA
            acc = !(r2 >= 1.0 || r2 == 0.0);
            if (acc) {
                const double f = sqrt(-2.0 * log(r2) / r2);
                g0 = __dmul_rn(f, x2);              // returned first
                g1 = __dmul_rn(f, x1);              // the cached deviate, returned by the next call
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) before += s_cnt[w];
        if (acc) {
            const long long i0 = base + 2LL * before;
            const float v[2] = {(float)g0, (float)g1};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long i = i0 + e;
                if (i < n_h) {
                    const long long k = i / T_total, t = i - k * T_total;
                    if (t >= t_begin && t < t_begin + t_count) Hn[k * t_count + (t - t_begin)] = v[e];
                } else if (i < n_total) {
                    Wn[i - n_h] = v[e];
                }
            }
        }
        int tot = 0;
        for (int w = 0; w < kRngThreads / 32; ++w) tot += s_cnt[w];
        base += 2LL * tot;
        __syncthreads();                            // s_cnt is rewritten by the next refill
    }
}

cudaError_t launch_numpy_normals(uint32_t seed, int K, long long T_total, long long t_begin, long long t_count, float* Hn, int F, float* Wn,
                                 cudaStream_t s) {
    AINMF_LAUNCH(numpy_normals_kernel, dim3(1), dim3(kRngThreads), 0, s, seed, (long long)K * T_total, T_total, t_begin, t_count, Hn,
                 (long long)F * K, Wn);
    return cudaGetLastError();
}

}  // namespace ainmf
#endif  // AINMF_EMU
