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

// Shared-memory layout of the transforms: element i of sequence fr sits at fr * fft_pitch(M) + fft_pad(i, ps) -- one unused slot
// after every 8 elements, which makes the stride-8 stores of a radix-8 Stockham stage (and the stride-4 / stride-2 ones of the
// tail stages) fall on distinct banks, while consecutive elements stay (almost) consecutive for the reads.
// (n_fft = 4096 keeps the dense layout: its buffers already fill the shared memory of an SM.)
__host__ __device__ __forceinline__ int fft_pad_shift(int M) { return (M >= 2048) ? 31 : 3; }
__host__ __device__ __forceinline__ int fft_pad(int i, int ps) { return i + (i >> ps); }
__host__ __device__ __forceinline__ int fft_pitch(int M) { return M + (M >> fft_pad_shift(M)); }

// Forward DFT (e^{-i...}) of `nfr` independent length-M sequences held in a (padded layout above); b is a scratch
// buffer of the same size; tw[k] = exp(-2 pi i k / M), k < M.  M is a power of two >= 2.  Stockham autosort, radix 8 with a
// radix-4 / radix-2 tail: three passes over shared memory for M = 512 instead of five with radix 4.
// All threads of the block must call; returns the buffer that holds the result (a or b).  Ends with a __syncthreads().
__device__ __forceinline__ float2* block_fft_forward(float2* a, float2* b, const float2* tw, int M, int nfr) {
    const int MP = fft_pitch(M), ps = fft_pad_shift(M);
    for (int Ns = 1; Ns < M;) {
        const int R = (Ns * 8 <= M) ? 8 : ((Ns * 4 <= M) ? 4 : 2);
        const int per = M / R;                 // butterflies per sequence
        const int total = per * nfr;
        const int tstride = M / (Ns * R);
        const int lper = 31 - __clz(per);      // per is a power of two
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int fr = idx >> lper;
            const int j = idx & (per - 1);
            const int k = j & (Ns - 1);
            const float2* in = a + fr * MP;
            float2* out = b + fr * MP;
            const int j0 = (j - k) * R + k;
            const int ti = k * tstride;
            if (R == 8) {
                float2 v[8];
                v[0] = in[fft_pad(j, ps)];
#pragma unroll
                for (int q = 1; q < 8; ++q) v[q] = cmul(in[fft_pad(j + q * per, ps)], tw[q * ti]);
                // two 4-point transforms (even / odd inputs), then the 8-point combination
                const float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), d26 = csub(v[2], v[6]);
                const float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]), a6 = cadd(v[3], v[7]), d37 = csub(v[3], v[7]);
                const float2 a3 = make_float2(d26.y, -d26.x), a7 = make_float2(d37.y, -d37.x);       // * (-i)
                const float2 e0 = cadd(a0, a2), e1 = cadd(a1, a3), e2 = csub(a0, a2), e3 = csub(a1, a3);
                const float2 o0 = cadd(a4, a6), o1 = cadd(a5, a7), d46 = csub(a4, a6), o3 = csub(a5, a7);
                const float h = 0.70710678118654752f;
                const float2 b5 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));                  // * (1 - i)/sqrt 2
                const float2 b6 = make_float2(d46.y, -d46.x);                                          // * (-i)
                const float2 b7 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));                 // * (-1 - i)/sqrt 2
                out[fft_pad(j0, ps)] = cadd(e0, o0);
                out[fft_pad(j0 + Ns, ps)] = cadd(e1, b5);
                out[fft_pad(j0 + 2 * Ns, ps)] = cadd(e2, b6);
                out[fft_pad(j0 + 3 * Ns, ps)] = cadd(e3, b7);
                out[fft_pad(j0 + 4 * Ns, ps)] = csub(e0, o0);
                out[fft_pad(j0 + 5 * Ns, ps)] = csub(e1, b5);
                out[fft_pad(j0 + 6 * Ns, ps)] = csub(e2, b6);
                out[fft_pad(j0 + 7 * Ns, ps)] = csub(e3, b7);
            } else if (R == 4) {
                float2 v0 = in[fft_pad(j, ps)];
                float2 v1 = cmul(in[fft_pad(j + per, ps)], tw[ti]);
                float2 v2 = cmul(in[fft_pad(j + 2 * per, ps)], tw[2 * ti]);
                float2 v3 = cmul(in[fft_pad(j + 3 * per, ps)], tw[3 * ti]);
                float2 s0 = cadd(v0, v2), s1 = csub(v0, v2), s2 = cadd(v1, v3), d = csub(v1, v3);
                float2 s3 = make_float2(d.y, -d.x);         // (v1 - v3) * (-i)
                out[fft_pad(j0, ps)] = cadd(s0, s2);
                out[fft_pad(j0 + Ns, ps)] = cadd(s1, s3);
                out[fft_pad(j0 + 2 * Ns, ps)] = csub(s0, s2);
                out[fft_pad(j0 + 3 * Ns, ps)] = csub(s1, s3);
            } else {
                float2 v0 = in[fft_pad(j, ps)];
                float2 v1 = cmul(in[fft_pad(j + per, ps)], tw[ti]);
                out[fft_pad(j0, ps)] = cadd(v0, v1);
                out[fft_pad(j0 + Ns, ps)] = csub(v0, v1);
            }
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= R;
    }
    return a;
}

}  // namespace ainmf
