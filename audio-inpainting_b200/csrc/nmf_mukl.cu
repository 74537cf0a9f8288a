// nmf_mukl.cu -- multiplicative update for the generalised Kullback-Leibler divergence, one fused kernel per half-step:
// the product W.H of a tile, the ratio R = X / max(W.H, eps) and the contraction of R with the other factor all stay on
// the SM; neither W.H nor R is ever written to memory (an iteration reads X twice and the factors, nothing else).
//
// sklearn (solver='mu', beta_loss='kullback-leibler'), $SP/sklearn/decomposition/_nmf.py:
//   W update :551-626   numerator (X / WH).H^T with WH < eps -> eps, denominator = row sums of H (0 -> eps)
//   H update :636-721   numerator W^T.(X / WH) with the NEW W, denominator = column sums of W (0 -> 1); :862-864 H < 2.2e-16 -> 0
//   error    :129-154   sum over x > eps of x log(x / max(WH, eps)) - x, plus sum(WH); reported as sqrt(2 res); :867-879
//                       convergence test every 10th iteration
// FFMA throughout: at K = 64 the two contractions are 8 F T K flops per iteration against 8 F T bytes, i.e. compute-bound
// on the FP32 pipe; the tensor-core form of this kernel (R staged through TMEM as the H step stages X) is not built.
#include "kernels.h"

namespace ainmf {

namespace {

// packed fp32x2 FMA (sm_100a FFMA2: two lanes per issue slot); plain fmaf pairs in the CPU test harness
__device__ __forceinline__ float2 kl_fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(AINMF_EMU)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

constexpr float kEps32 = 1.1920929e-07f;      // np.finfo(np.float32).eps: sklearn's EPSILON
constexpr int KL_TR = 64;                     // rows of the updated factor per block
constexpr int KL_TC = 64;                     // columns (rows of the other factor) per chunk

template <int KP> struct KlCfg {
    static constexpr int PT = KL_TR + 4;      // pitch of the k-major / c-major tiles (float4 reads along rows)
    static constexpr int PB = KP + 4;         // pitch of the chunk of the other factor, row-major
    static constexpr int KQ = KP / 16;        // components per thread in the second contraction
    static constexpr size_t smem_bytes = sizeof(float) * ((size_t)KP * PT * 2 + (size_t)KL_TC * PB + (size_t)KL_TC * PT);
};

// MODE 0: H update (rows = frames, columns = bins; X contiguous along columns)
// MODE 1: W update (rows = bins, columns = frames; X contiguous along rows)
// MODE 2: error terms over a tile of frames (first contraction only)
// A: [rows][KP] the factor of the rows, Bm: [cols][KP] the other factor; Xt: [T][ldf].
// grid = (ceil(rows / 64), B), 256 threads: thread (ty, tx) = (tid / 16, tid % 16) owns rows 4ty..4ty+3 and, in the first
// contraction, columns 4tx..4tx+3 of the chunk, in the second components tx*KQ..tx*KQ+KQ-1.
template <int KP, int MODE>
__global__ void __launch_bounds__(kThreads)
kl_update_kernel(float* __restrict__ A, long long a_stride, int n_rows, const float* __restrict__ Bm, long long b_stride,
                 int n_cols, const float* __restrict__ Xt, long long x_stride, int ldf, const float* __restrict__ den /*[B][KP]*/,
                 double* __restrict__ err_partial /*[B][gridDim.x]*/, const ClipState* __restrict__ st) {
    using Cfg = KlCfg<KP>;
    constexpr int PT = Cfg::PT, PB = Cfg::PB, KQ = Cfg::KQ;
    AINMF_DYN_SMEM(smem_raw);
    float* sAT = reinterpret_cast<float*>(smem_raw);      // [KP][PT]: A tile, component-major
    float* sBT = sAT + KP * PT;                           // [KP][PT]: chunk of the other factor, component-major
    float* sB = sBT + KP * PT;                            // [TC][PB]: the same chunk, row-major
    float* sRT = sB + KL_TC * PB;                         // [TC][PT]: ratio tile, column-major
    __shared__ double s_red[32];
    const int b = blockIdx.y;
    if (st && st[b].done) return;                         // st == null: evaluate every clip (final error)
    const int r0 = blockIdx.x * KL_TR;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float* Ab = A + (long long)b * a_stride;
    const float* Bb = Bm + (long long)b * b_stride;
    const float* Xb = Xt + (long long)b * x_stride;
    // A tile: lane = row (consecutive rows -> conflict-free transposed stores)
    for (int i = tid; i < KL_TR * (KP / 4); i += kThreads) {
        const int r = i % KL_TR, k4 = (i / KL_TR) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < n_rows) v = *reinterpret_cast<const float4*>(Ab + (long long)(r0 + r) * KP + k4);
        sAT[(k4 + 0) * PT + r] = v.x; sAT[(k4 + 1) * PT + r] = v.y; sAT[(k4 + 2) * PT + r] = v.z; sAT[(k4 + 3) * PT + r] = v.w;
    }
    float acc[4][KQ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < KQ; ++j) acc[i][j] = 0.f;
    double esum = 0.0;
    for (int c0 = 0; c0 < n_cols; c0 += KL_TC) {
        __syncthreads();                                   // previous chunk consumed (and the A tile written)
        for (int i = tid; i < KL_TC * (KP / 4); i += kThreads) {
            const int c = i % KL_TC, k4 = (i / KL_TC) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + c < n_cols) v = *reinterpret_cast<const float4*>(Bb + (long long)(c0 + c) * KP + k4);
            sBT[(k4 + 0) * PT + c] = v.x; sBT[(k4 + 1) * PT + c] = v.y; sBT[(k4 + 2) * PT + c] = v.z; sBT[(k4 + 3) * PT + c] = v.w;
            *reinterpret_cast<float4*>(sB + c * PB + k4) = v;
        }
        // the X tile of this thread: x[i][j] = X(row 4ty+i, column 4tx+j); zero outside the matrix
        float x[4][4];
        if (MODE == 1) {                                   // rows = bins (contiguous), columns = frames
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = c0 + 4 * tx + j, f = r0 + 4 * ty;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < n_cols && f < ldf) v = *reinterpret_cast<const float4*>(Xb + (long long)t * ldf + f);
                x[0][j] = v.x; x[1][j] = v.y; x[2][j] = v.z; x[3][j] = v.w;
            }
        } else {                                           // rows = frames, columns = bins (contiguous)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = r0 + 4 * ty + i, f = c0 + 4 * tx;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < n_rows && f < ldf) v = *reinterpret_cast<const float4*>(Xb + (long long)t * ldf + f);
                x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
            }
        }
        __syncthreads();
        // first contraction: wh[i][j] = sum_k A[row i][k] * Bm[column j][k]
        float2 wh2[4][2];                                  // [row][column pair]: packed FFMA2
#pragma unroll
        for (int i = 0; i < 4; ++i) { wh2[i][0] = make_float2(0.f, 0.f); wh2[i][1] = make_float2(0.f, 0.f); }
#pragma unroll 4
        for (int k = 0; k < KP; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(sAT + k * PT + 4 * ty);
            const float4 c = *reinterpret_cast<const float4*>(sBT + k * PT + 4 * tx);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float2 c01 = make_float2(c.x, c.y), c23 = make_float2(c.z, c.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 aa = make_float2(av[i], av[i]);
                wh2[i][0] = kl_fma2(aa, c01, wh2[i][0]);
                wh2[i][1] = kl_fma2(aa, c23, wh2[i][1]);
            }
        }
        float wh[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { wh[i][0] = wh2[i][0].x; wh[i][1] = wh2[i][0].y; wh[i][2] = wh2[i][1].x; wh[i][3] = wh2[i][1].y; }
        if (MODE == 2) {
            // x > eps: x log(x / max(wh, eps)) - x; every element of the matrix: + wh (the sum of W.H)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool inside = (r0 + 4 * ty + i < n_rows) && (c0 + 4 * tx + j < n_cols);
                    if (inside) {
                        double term = (double)wh[i][j];
                        if (x[i][j] > kEps32) {
                            const float q = x[i][j] / fmaxf(wh[i][j], kEps32);
                            term += (double)(x[i][j] * logf(q)) - (double)x[i][j];
                        }
                        esum += term;
                    }
                }
            continue;
        }
        // ratio tile, column-major: sRT[column][row]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 rv;
            rv.x = x[0][j] / fmaxf(wh[0][j], kEps32); rv.y = x[1][j] / fmaxf(wh[1][j], kEps32);
            rv.z = x[2][j] / fmaxf(wh[2][j], kEps32); rv.w = x[3][j] / fmaxf(wh[3][j], kEps32);
            *reinterpret_cast<float4*>(sRT + (4 * tx + j) * PT + 4 * ty) = rv;
        }
        __syncthreads();
        // second contraction: acc[i][q] += sum_c R[row i][c] * Bm[c][tx*KQ + q]
#pragma unroll 4
        for (int c = 0; c < KL_TC; ++c) {
            const float4 rr = *reinterpret_cast<const float4*>(sRT + c * PT + 4 * ty);
            const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
            float bv[KQ];
            if constexpr (KQ == 2) {
                const float2 t2 = *reinterpret_cast<const float2*>(sB + c * PB + tx * KQ);
                bv[0] = t2.x; bv[1] = t2.y;
            } else {
#pragma unroll
                for (int q = 0; q < KQ; q += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(sB + c * PB + tx * KQ + q);
                    bv[q] = t4.x; bv[q + 1] = t4.y; bv[q + 2] = t4.z; bv[q + 3] = t4.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 rr2 = make_float2(rv[i], rv[i]);
#pragma unroll
                for (int q = 0; q < KQ; q += 2) {
                    const float2 t2 = kl_fma2(rr2, make_float2(bv[q], bv[q + 1]), make_float2(acc[i][q], acc[i][q + 1]));
                    acc[i][q] = t2.x; acc[i][q + 1] = t2.y;
                }
            }
        }
    }
    if (MODE == 2) {
        const double tot = block_sum_d(esum, s_red);
        if (tid == 0) err_partial[(long long)b * gridDim.x + blockIdx.x] = tot;
        return;
    }
    // A[row][k] *= numerator / denominator; H additionally: values below float64 eps -> 0
    float dv[KQ];
#pragma unroll
    for (int q = 0; q < KQ; ++q) dv[q] = den[(long long)b * KP + tx * KQ + q];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * ty + i;
        if (r >= n_rows) continue;
        float* ar = Ab + (long long)r * KP + tx * KQ;
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
            float v = ar[q] * (acc[i][q] / dv[q]);
            if (MODE == 0 && (double)v < 2.220446049250313e-16) v = 0.f;
            ar[q] = v;
        }
    }
}

// out[b][k] = sum over rows of A[b][row][k]; zeros replaced by `repl` (the denominators of the two updates); one block per clip
__global__ void __launch_bounds__(kThreads)
kl_colsum_kernel(const float* __restrict__ A, long long a_stride, int rows, int KP, float repl, float* __restrict__ out,
                 const ClipState* __restrict__ st) {
    __shared__ float s_part[kThreads];
    const int b = blockIdx.x;
    if (st[b].done) return;
    const int k = threadIdx.x % KP, g = threadIdx.x / KP, G = kThreads / KP;
    float v = 0.f;
    for (int r = g; r < rows; r += G) v += A[(long long)b * a_stride + (long long)r * KP + k];
    s_part[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < KP) {
        float t = 0.f;
        for (int i = 0; i < G; ++i) t += s_part[i * KP + threadIdx.x];       // fixed order
        out[(long long)b * KP + threadIdx.x] = (t == 0.f) ? repl : t;
    }
}

// err = sqrt(2 * sum of the tile terms); `keep` (final evaluation) receives the value as well
__global__ void __launch_bounds__(kThreads)
kl_err_reduce_kernel(ClipState* __restrict__ st, int B, const double* __restrict__ err_partial, int n, double* __restrict__ keep) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    if (st[b].done && !keep) return;
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v += err_partial[(long long)b * n + i];
    v = warp_sum_d(v);
    if (lane == 0) {
        const double e = sqrt(2.0 * fmax(v, 0.0));
        if (keep) keep[b] = e; else st[b].err = (float)e;
    }
}
__global__ void __launch_bounds__(kThreads)
kl_set_err_kernel(ClipState* __restrict__ st, int B, const double* __restrict__ keep) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) st[b].err = (float)keep[b];
}

template <int KP, int MODE>
cudaError_t kl_launch(float* A, long long a_stride, int n_rows, const float* Bm, long long b_stride, int n_cols,
                      const NmfProblem& p, const float* den, double* err_partial, cudaStream_t s, bool all_clips = false) {
    auto kern = kl_update_kernel<KP, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KlCfg<KP>::smem_bytes);
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(kern, dim3(ceil_div(n_rows, KL_TR), p.B), dim3(kThreads), KlCfg<KP>::smem_bytes, s, A, a_stride, n_rows, Bm,
                 b_stride, n_cols, p.Xt, p.x_stride, p.ldf, den, err_partial, all_clips ? nullptr : p.state);
    return cudaGetLastError();
}

template <int KP>
cudaError_t kl_error_impl(const NmfProblem& p, const NmfWork& wk, double* keep, cudaStream_t s) {
    // tiles of frames; the columns are the F bins (pad bins are outside the matrix)
    cudaError_t e = kl_launch<KP, 2>(p.Ht, p.h_stride, p.T, p.W, p.w_stride, p.F, p, nullptr, wk.err_partial, s, keep != nullptr);
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(kl_err_reduce_kernel, dim3(ceil_div(p.B, kThreads / 32)), dim3(kThreads), 0, s, p.state, p.B, wk.err_partial,
                 ceil_div(p.T, KL_TR), keep);
    return cudaGetLastError();
}

template <int KP>
cudaError_t kl_iterate_impl(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    cudaError_t e;
    float* hsum = wk.kl_sums;
    float* wsum = wk.kl_sums + (size_t)p.B * KP;
    // W <- W * ((X / WH) H^T) / rowsum(H)
    AINMF_LAUNCH(kl_colsum_kernel, dim3(p.B), dim3(kThreads), 0, s, p.Ht, p.h_stride, p.T, KP, kEps32, hsum, p.state);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = kl_launch<KP, 1>(p.W, p.w_stride, p.F, p.Ht, p.h_stride, p.T, p, hsum, nullptr, s)) != cudaSuccess) return e;
    // H <- H * (W^T (X / WH)) / colsum(W), with the new W
    AINMF_LAUNCH(kl_colsum_kernel, dim3(p.B), dim3(kThreads), 0, s, p.W, p.w_stride, p.F, KP, 1.0f, wsum, p.state);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return kl_launch<KP, 0>(p.Ht, p.h_stride, p.T, p.W, p.w_stride, p.F, p, wsum, nullptr, s);
}

}  // namespace

cudaError_t nmf_mukl_error(const NmfProblem& p, const NmfWork& wk, bool keep, cudaStream_t s) {
    double* k = keep ? wk.kl_err : nullptr;
    switch (p.KP) {
        case 32: return kl_error_impl<32>(p, wk, k, s);
        case 64: return kl_error_impl<64>(p, wk, k, s);
        case 128: return kl_error_impl<128>(p, wk, k, s);
    }
    return (cudaError_t)1;
}
cudaError_t nmf_mukl_set_err(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    AINMF_LAUNCH(kl_set_err_kernel, dim3(ceil_div(p.B, kThreads)), dim3(kThreads), 0, s, p.state, p.B, wk.kl_err);
    return cudaGetLastError();
}
cudaError_t nmf_mukl_begin(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    cudaError_t e = nmf_mukl_error(p, wk, false, s);     // error of the initial factors (_nmf.py:822)
    return e != cudaSuccess ? e : nmf_mu_stop(p, 0, s);
}
cudaError_t nmf_mukl_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s) {
    cudaError_t e = (cudaError_t)1;
    switch (p.KP) {
        case 32: e = kl_iterate_impl<32>(p, wk, s); break;
        case 64: e = kl_iterate_impl<64>(p, wk, s); break;
        case 128: e = kl_iterate_impl<128>(p, wk, s); break;
    }
    if (e != cudaSuccess) return e;
    if ((e = nmf_mu_tick(p, it, s)) != cudaSuccess) return e;
    if (p.tol > 0.f && it % 10 == 0) {
        if ((e = nmf_mukl_error(p, wk, false, s)) != cudaSuccess) return e;
        return nmf_mu_stop(p, it, s);
    }
    return cudaSuccess;
}

}  // namespace ainmf
