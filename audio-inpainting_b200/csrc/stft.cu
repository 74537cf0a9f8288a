// stft.cu -- K1 (framing + window + real FFT + scale + magnitude) and K5 (recombine + inverse real FFT
// + synthesis window + overlap-add + normalise + trim), hand-written for sm_100a.
//
// Follows scipy.signal.stft / istft as called at main4_NMF_gap.py:47,71 (main4_NMF_mask.py:52,76,
// main4_NMF.py:69,93): $SP/scipy/signal/_spectral_py.py:2240-2247 (zero extension by n_fft/2, tail pad),
// :2378-2395 (frames every hop, window multiply, rfft), :2277-2316 (scale 1/sum(w)); inverse :1872-1910
// (irfft, * sum(w), windowed overlap-add, sum w^2 normalisation where > 1e-10, drop n_fft/2 each side).
#include "fft.cuh"
#include "kernels.h"

namespace ainmf {

constexpr int kStftFramesPerPass = 4;    // frames transformed together by one block
constexpr int kStftFramesPerBlock = 8;   // frames owned by one block (shares the waveform segment + tables)
constexpr int kIstftHopsPerBlock = 16;   // output hops owned by one block of the inverse

// -------------------------------------------------------------------------------------------------
// K1.  grid = (ceil(t_count / kStftFramesPerBlock), B).  Local frame lt <-> global frame t_begin + lt.
// Frame t covers clip samples [t*hop - n/2, t*hop + n/2); samples outside [0, N) are zero.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
stft_kernel(const float* __restrict__ x, long long x_stride, long long x_origin, long long x_avail, long long N,
            int n_fft, int hop,
            int t_begin, int t_count, int F, int ldf, const float2* __restrict__ tw_half,
            const float2* __restrict__ tw_full, const float* __restrict__ window, float scale,
            float* __restrict__ V, float2* __restrict__ Z, long long vz_stride) {
    AINMF_DYN_SMEM(smem_raw);
    const int M = n_fft >> 1, lM = 31 - __clz(M);
    const int seg_len = (kStftFramesPerBlock - 1) * hop + n_fft;
    float2* s_twh = reinterpret_cast<float2*>(smem_raw);          // [M]
    float2* s_twf = s_twh + M;                                     // [M + 1] (+1 pad to keep 16B alignment below)
    const int MP = fft_pitch(M), ps = fft_pad_shift(M);
    float2* s_a = s_twf + (M + 2);                                 // [FP][MP] (padded layout of fft.cuh)
    float2* s_b = s_a + kStftFramesPerPass * MP;                   // [FP][MP]
    float* s_win = reinterpret_cast<float*>(s_b + kStftFramesPerPass * MP);  // [n_fft]
    float* s_seg = s_win + n_fft;                                  // [seg_len]

    const int b = blockIdx.y;
    const int lt0 = blockIdx.x * kStftFramesPerBlock;
    const float* xb = x + (long long)b * x_stride;
    const long long seg_start = (long long)(t_begin + lt0) * hop - M;   // clip sample index of s_seg[0]

    for (int i = threadIdx.x; i < M; i += blockDim.x) s_twh[i] = tw_half[i];
    for (int i = threadIdx.x; i <= M; i += blockDim.x) s_twf[i] = tw_full[i];
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) s_win[i] = window[i];
    for (int i = threadIdx.x; i < seg_len; i += blockDim.x) {
        const long long sidx = seg_start + i;
        const long long li = sidx - x_origin;
        s_seg[i] = (sidx >= 0 && sidx < N && li >= 0 && li < x_avail) ? xb[li] : 0.f;
    }
    __syncthreads();

    for (int pass = 0; pass < kStftFramesPerBlock / kStftFramesPerPass; ++pass) {
        const int fbase = pass * kStftFramesPerPass;               // first frame (within block) of this pass
        // windowed frames -> half-length complex sequences z[j] = (xw[2j], xw[2j+1])
        for (int idx = threadIdx.x; idx < kStftFramesPerPass * M; idx += blockDim.x) {
            const int q = idx >> lM, j = idx & (M - 1);
            const float* fr = s_seg + (fbase + q) * hop;
            s_a[q * MP + fft_pad(j, ps)] = make_float2(fr[2 * j] * s_win[2 * j], fr[2 * j + 1] * s_win[2 * j + 1]);
        }
        __syncthreads();
        const float2* Zh = block_fft_forward(s_a, s_b, s_twh, M, kStftFramesPerPass);
        // X[k] = E[k] + e^{-2 pi i k/n} O[k],  E = (Z[k] + conj Z[M-k])/2,  O = (Z[k] - conj Z[M-k])/(2i)
        for (int q = 0; q < kStftFramesPerPass; ++q)
        for (int k = threadIdx.x; k < ldf; k += blockDim.x) {
            const int lt = lt0 + fbase + q;
            if (lt >= t_count) continue;
            float2 X = make_float2(0.f, 0.f);
            float mag = 0.f;
            if (k < F) {
                const float2* zq = Zh + q * MP;
                const float2 za = zq[fft_pad(k & (M - 1), ps)];
                const float2 zb = zq[fft_pad((M - k) & (M - 1), ps)];
                const float2 E = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
                const float2 O = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));
                const float2 t = cmul(s_twf[k], O);
                X = make_float2((E.x + t.x) * scale, (E.y + t.y) * scale);
                if (k == 0 || k == M) X.y = 0.f;                  // exactly real for real input
                mag = hypotf(X.x, X.y);
            }
            const long long o = (long long)b * vz_stride + (long long)lt * ldf + k;
            V[o] = mag;
            Z[o] = X;
        }
        __syncthreads();
    }
}

static size_t stft_smem_bytes(int n_fft, int hop) {
    const int M = n_fft / 2;
    return sizeof(float2) * (size_t)(M + (M + 2) + 2 * kStftFramesPerPass * fft_pitch(M)) +
           sizeof(float) * (size_t)(n_fft + (kStftFramesPerBlock - 1) * hop + n_fft);
}

cudaError_t launch_stft(const float* x, long long x_stride, long long x_origin, long long x_avail, int B,
                        const StftGeom& g,
                        int t_begin, int t_count, const FftTables& tb, float* V, float2* Z,
                        long long vz_stride, cudaStream_t s) {
    if (t_count <= 0 || B <= 0) return cudaSuccess;
    const size_t smem = stft_smem_bytes(g.n_fft, g.hop);
    cudaError_t e = cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(t_count, kStftFramesPerBlock), B);
    AINMF_LAUNCH(stft_kernel, grid, dim3(kThreads), smem, s, x, x_stride, x_origin, x_avail, g.N, g.n_fft, g.hop,
                 t_begin, t_count, g.F, g.ldf, tb.tw_half, tb.tw_full, tb.window, 1.0f / tb.win_sum, V, Z,
                 vz_stride);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// K5.  grid = (ceil(n_count / (kIstftHopsPerBlock*hop)), B).  A block owns output samples
// [s0, s0 + G*hop) of the clip and gathers every frame that overlaps them (no atomics; frames are
// added in ascending order, the order of scipy's loop at _spectral_py.py:1892-1894).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
istft_kernel(const float* __restrict__ V, const float2* __restrict__ Z, long long vz_stride,
             const unsigned char* __restrict__ bad, long long bad_stride, const int* __restrict__ n_bad,
             const float* __restrict__ x, long long x_stride, long long x_origin, long long N, int n_fft,
             int hop, int t_begin, int t_count, int T_total, int F, int ldf,
             const float2* __restrict__ tw_half, const float2* __restrict__ tw_full,
             const float* __restrict__ window, float win_sum, float* __restrict__ y, long long y_stride,
             long long n_begin, long long n_count) {
    AINMF_DYN_SMEM(smem_raw);
    const int M = n_fft >> 1, lM = 31 - __clz(M);
    const int tile = kIstftHopsPerBlock * hop;
    float2* s_twh = reinterpret_cast<float2*>(smem_raw);          // [M]
    float2* s_twf = s_twh + M;                                     // [M + 1]
    const int MP = fft_pitch(M), ps = fft_pad_shift(M);
    float2* s_a = s_twf + (M + 2);                                 // [FP][MP] (padded layout of fft.cuh)
    float2* s_b = s_a + kStftFramesPerPass * MP;                   // [FP][MP]
    float* s_win = reinterpret_cast<float*>(s_b + kStftFramesPerPass * MP);  // [n_fft]
    float* s_out = s_win + n_fft;                                  // [tile]

    const int b = blockIdx.y;
    const long long s0 = n_begin + (long long)blockIdx.x * tile;  // first clip sample of this block
    const long long s_end = (s0 + tile < n_begin + n_count) ? s0 + tile : n_begin + n_count;
    float* yb = y + (long long)b * y_stride;

    if (n_bad[b] == 0) {   // whole block takes this branch: nothing to restore -> return the input
        const float* xb = x + (long long)b * x_stride;
        for (long long n = s0 + threadIdx.x; n < s_end; n += blockDim.x) yb[n - n_begin] = xb[n - x_origin];
        return;
    }

    // frames c with c*hop <= e < c*hop + n_fft for some e = n + M, n in [s0, s_end)
    long long c_lo = (s0 + M - n_fft >= 0) ? (s0 + M - n_fft) / hop + 1 : 0;     // floor((e0-n_fft)/hop)+1
    long long c_hi = (s_end - 1 + M) / hop;
    if (c_hi > T_total - 1) c_hi = T_total - 1;
    if (c_lo < t_begin) c_lo = t_begin;                      // caller guarantees these are not needed
    if (c_hi > (long long)t_begin + t_count - 1) c_hi = (long long)t_begin + t_count - 1;
    const unsigned char* badb = bad + (long long)b * bad_stride;

    // No frame that reaches into this block's samples was modified: every such frame is the transform of x * w, so
    // the windowed overlap-add returns x * sum(w^2) and the normalisation divides it out again (_spectral_py.py:
    // 1892-1910).  Write x (times sum w^2 where scipy leaves the sum un-normalised) and skip the transforms.
    if (x != nullptr) {
        __shared__ int s_any;
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        int any = 0;
        for (long long c = c_lo + threadIdx.x; c <= c_hi; c += blockDim.x) any |= badb[c - t_begin];
        if (any) s_any = 1;
        __syncthreads();
        if (!s_any) {
            const float* xb = x + (long long)b * x_stride;
            for (long long n = s0 + threadIdx.x; n < s_end; n += blockDim.x) {
                const long long e = n + M;
                long long ca = (e - n_fft >= 0) ? (e - n_fft) / hop + 1 : 0;
                long long cb = e / hop;
                if (cb > T_total - 1) cb = T_total - 1;
                float nr = 0.f;
                for (long long c = ca; c <= cb; ++c) { const float w = window[e - c * hop]; nr += w * w; }
                const float xv = xb[n - x_origin];
                yb[n - n_begin] = (nr > 1e-10f) ? xv : xv * nr;
            }
            return;
        }
    }

    for (int i = threadIdx.x; i < M; i += blockDim.x) s_twh[i] = tw_half[i];
    for (int i = threadIdx.x; i <= M; i += blockDim.x) s_twf[i] = tw_full[i];
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) s_win[i] = window[i];
    for (int i = threadIdx.x; i < tile; i += blockDim.x) s_out[i] = 0.f;
    __syncthreads();

    const float inv_M = 1.0f / (float)M;
    for (long long c0 = c_lo; c0 <= c_hi; c0 += kStftFramesPerPass) {
        // half-length spectrum for the inverse: Zk[k] = E[k] + i O[k], conjugated so that the forward
        // routine computes the inverse transform.
        for (int idx = threadIdx.x; idx < kStftFramesPerPass * M; idx += blockDim.x) {
            const int q = idx >> lM, k = idx & (M - 1);
            const long long c = c0 + q;
            float2 out = make_float2(0.f, 0.f);
            if (c <= c_hi) {
                const int lt = (int)(c - t_begin);
                const long long row = (long long)b * vz_stride + (long long)lt * ldf;
                float2 Xa = Z[row + k];
                float2 Xb = Z[row + (M - k)];
                if (badb[lt]) {     // magnitude from the model, phase from the corrupted signal
                    const float ma = hypotf(Xa.x, Xa.y), mb = hypotf(Xb.x, Xb.y);
                    const float va = V[row + k], vb = V[row + (M - k)];
                    Xa = (ma > 0.f) ? make_float2(va * (Xa.x / ma), va * (Xa.y / ma)) : make_float2(va, 0.f);
                    Xb = (mb > 0.f) ? make_float2(vb * (Xb.x / mb), vb * (Xb.y / mb)) : make_float2(vb, 0.f);
                }
                if (k == 0) { Xa.y = 0.f; Xb.y = 0.f; }      // c2r ignores imag of DC and Nyquist
                const float2 E = make_float2(0.5f * (Xa.x + Xb.x), 0.5f * (Xa.y - Xb.y));
                const float2 D = make_float2(0.5f * (Xa.x - Xb.x), 0.5f * (Xa.y + Xb.y));
                const float2 w = s_twf[k];                    // e^{-i th}; need e^{+i th} = conj
                const float2 O = cmul(make_float2(w.x, -w.y), D);
                // Zk = E + i*O ; store conj(Zk)
                out = make_float2(E.x - O.y, -(E.y + O.x));
            }
            s_a[q * MP + fft_pad(k, ps)] = out;
        }
        __syncthreads();
        const float2* zt = block_fft_forward(s_a, s_b, s_twh, M, kStftFramesPerPass);
        // x[2j] = Re conj(zt[j]) / M, x[2j+1] = Im conj(zt[j]) / M ; then * sum(w) * w[j], overlap-add
        for (int q = 0; q < kStftFramesPerPass; ++q) {
            const long long c = c0 + q;
            if (c <= c_hi) {   // uniform across the block
                const int base = (int)(c * hop - M - s0);      // tile offset of frame sample 0
                for (int j = threadIdx.x; j < n_fft; j += blockDim.x) {
                    const int p = base + j;
                    if (p >= 0 && p < tile) {
                        const float2 z = zt[q * MP + fft_pad(j >> 1, ps)];
                        float v = (j & 1) ? -z.y : z.x;
                        v = (v * inv_M) * win_sum;
                        s_out[p] += v * s_win[j];
                    }
                }
            }
            __syncthreads();
        }
    }
    // sum of w^2 over the frames that cover a sample, in the order the frames were added
    for (long long n = s0 + threadIdx.x; n < s_end; n += blockDim.x) {
        const int p = (int)(n - s0);
        const long long e = n + M;
        long long ca = (e - n_fft >= 0) ? (e - n_fft) / hop + 1 : 0;
        long long cb = e / hop;
        if (ca < c_lo) ca = c_lo;
        if (cb > c_hi) cb = c_hi;
        float nr = 0.f;
        for (long long c = ca; c <= cb; ++c) { const float w = s_win[e - c * hop]; nr += w * w; }
        yb[n - n_begin] = s_out[p] / (nr > 1e-10f ? nr : 1.0f);
    }
}

static size_t istft_smem_bytes(int n_fft, int hop);
size_t stft_smem_need(int n_fft, int hop, bool forward, bool inverse) {
    const size_t a = forward ? stft_smem_bytes(n_fft, hop) : 0, b = inverse ? istft_smem_bytes(n_fft, hop) : 0;
    return a > b ? a : b;
}
static size_t istft_smem_bytes(int n_fft, int hop) {
    const int M = n_fft / 2;
    return sizeof(float2) * (size_t)(M + (M + 2) + 2 * kStftFramesPerPass * fft_pitch(M)) +
           sizeof(float) * (size_t)(n_fft + kIstftHopsPerBlock * hop);
}

cudaError_t launch_istft(const float* V, const float2* Z, long long vz_stride, const unsigned char* bad,
                         long long bad_stride, const int* n_bad, const float* x, long long x_stride,
                         long long x_origin, int B, const StftGeom& g, int t_begin, int t_count,
                         const FftTables& tb, float* y, long long y_stride, long long n_begin,
                         long long n_count, int T_total, cudaStream_t s) {
    if (n_count <= 0 || B <= 0) return cudaSuccess;
    const size_t smem = istft_smem_bytes(g.n_fft, g.hop);
    cudaError_t e = cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long tile = (long long)kIstftHopsPerBlock * g.hop;
    dim3 grid((unsigned)ceil_div64(n_count, tile), B);
    AINMF_LAUNCH(istft_kernel, grid, dim3(kThreads), smem, s, V, Z, vz_stride, bad, bad_stride, n_bad, x,
                 x_stride, x_origin, g.N, g.n_fft, g.hop, t_begin, t_count, T_total, g.F, g.ldf, tb.tw_half,
                 tb.tw_full, tb.window, tb.win_sum, y, y_stride, n_begin, n_count);
    return cudaGetLastError();
}

}  // namespace ainmf
