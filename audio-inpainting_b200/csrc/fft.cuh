// fft.cuh -- block-cooperative batched complex FFT in shared memory (Stockham autosort, radix 4 with a
// radix-2 tail), plus the half-length real-FFT pre/post steps.  Replaces pocketfft's r2c / c2r as
// called by scipy.signal.stft / istft ($SP/scipy/signal/_spectral_py.py:2395 and :1872).
#pragma once
#include "common.cuh"

namespace ainmf {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Forward DFT (e^{-i...}) of `nfr` independent length-M sequences held in a[nfr][M]; b is a scratch
// buffer of the same size; tw[k] = exp(-2 pi i k / M), k < M.  M is a power of two >= 2.
// All threads of the block must call; returns the buffer that holds the result (a or b).
// Ends with a __syncthreads().
__device__ __forceinline__ float2* block_fft_forward(float2* a, float2* b, const float2* tw, int M, int nfr) {
    for (int Ns = 1; Ns < M;) {
        const bool r4 = (Ns * 4 <= M);
        const int R = r4 ? 4 : 2;
        const int per = M / R;                 // butterflies per sequence
        const int total = per * nfr;
        const int tstride = M / (Ns * R);
        const int lper = 31 - __clz(per);      // per is a power of two
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int fr = idx >> lper;
            const int j = idx & (per - 1);
            const int k = j & (Ns - 1);
            const float2* in = a + fr * M;
            float2* out = b + fr * M;
            const int j0 = (j - k) * R + k;
            const int ti = k * tstride;
            if (r4) {
                float2 v0 = in[j];
                float2 v1 = cmul(in[j + per], tw[ti]);
                float2 v2 = cmul(in[j + 2 * per], tw[2 * ti]);
                float2 v3 = cmul(in[j + 3 * per], tw[3 * ti]);
                float2 s0 = cadd(v0, v2), s1 = csub(v0, v2), s2 = cadd(v1, v3), d = csub(v1, v3);
                float2 s3 = make_float2(d.y, -d.x);         // (v1 - v3) * (-i)
                out[j0] = cadd(s0, s2);
                out[j0 + Ns] = cadd(s1, s3);
                out[j0 + 2 * Ns] = csub(s0, s2);
                out[j0 + 3 * Ns] = csub(s1, s3);
            } else {
                float2 v0 = in[j];
                float2 v1 = cmul(in[j + per], tw[ti]);
                out[j0] = cadd(v0, v1);
                out[j0 + Ns] = csub(v0, v1);
            }
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= R;
    }
    return a;
}

}  // namespace ainmf
