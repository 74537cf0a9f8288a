// nmf_cd.cuh -- device building blocks of the coordinate-descent NMF iteration (FFMA path).
//
// Restates sklearn's solver="cd", beta_loss="frobenius" as the reference uses it
// (NMF(n_components, init='random', random_state, max_iter=200) at main4_NMF_gap.py:62,
// main4_NMF_mask.py:67, main4_NMF.py:83):  $SP/sklearn/decomposition/_nmf.py:369-396
// (_update_coordinate_descent: Gram + X.Ht products), _cdnmf_fast.pyx:8-38 (the sweep) and
// _nmf.py:491-516 (violation stop rule).
#pragma once
#include "common.cuh"

namespace ainmf {

// ---- register fragments of a shared-memory row ------------------------------------------------------
// A 16-thread-wide dimension owns TN = KP/16 values of a KP-wide row; the mapping keeps every warp
// LDS.128 conflict-free: TN=8 -> {4t..4t+3, 64+4t..64+4t+3}, TN=4 -> {4t..4t+3}, TN=2 -> {2t, 2t+1}.
template <int TN> __device__ __forceinline__ int frag_col(int t, int e) {
    if (TN == 8) return (e < 4) ? 4 * t + e : 64 + 4 * t + (e - 4);
    if (TN == 4) return 4 * t + e;
    return 2 * t + e;
}
template <int TN> __device__ __forceinline__ void load_frag(const float* row, int t, float (&f)[TN]) {
    if (TN == 8) {
        const float4 a = *reinterpret_cast<const float4*>(row + 4 * t);
        const float4 b = *reinterpret_cast<const float4*>(row + 64 + 4 * t);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else if (TN == 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + 4 * t);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    } else {
        const float2 a = *reinterpret_cast<const float2*>(row + 2 * t);
        f[0] = a.x; f[1] = a.y;
    }
}

// ---- the sweep ------------------------------------------------------------------------------------------
// One row of A (W or Ht) is owned by L consecutive lanes; lane l holds the contiguous slice
// a[q] = A[row][l*S + q], S = KP/L, and the matching slice of B (XHt or XtW).  The Gram matrix lives in
// shared memory with row pitch KP + 4L: element (t, r) at t*(KP+4L) + (r/S)*(S+4) + r%S, which spreads
// the L slices of a row over disjoint banks.
//
// for t in 0..KP-1 (the reference's order, permutation = arange):
//     grad = -B[row,t] + sum_r G[t,r] * A[row,r]        (uses the already updated A[row,<t])
//     pg   = (A[row,t] == 0) ? min(0, grad) : grad ;  violation += |pg|
//     if G[t,t] != 0:  A[row,t] = max(A[row,t] - grad / G[t,t], 0)
// Returns this lane's share of the violation.
// sInv (optional): sInv[t] = 1 / G[t,t] (0 where the diagonal is 0); when given, the update multiplies by the
// reciprocal instead of dividing (one rounding of difference per update; takes the division off the serial chain).
// pg_row (optional): the owner lane of coordinate t stores |pg| to pg_row[t] (exact-order violation sum, stop_kernel).
template <int KP, int L>
__device__ __forceinline__ float cd_sweep_row(float (&a)[KP / L], const float (&bv)[KP / L],
                                              const float* __restrict__ sG, int l, bool valid,
                                              const float* __restrict__ sInv = nullptr, float* __restrict__ pg_row = nullptr) {
    constexpr int S = KP / L;
    constexpr int PITCH = KP + 4 * L;
    float viol = 0.f;
    const float* gl = sG + l * (S + 4);
    for (int o = 0; o < L; ++o) {          // owner lane of coordinates [o*S, (o+1)*S)
#pragma unroll
        for (int q = 0; q < S; ++q) {
            const int t = o * S + q;
            const float* g = gl + t * PITCH;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
            if (S >= 4) {
#pragma unroll
                for (int r = 0; r < S; r += 4) {
                    const float4 gv = *reinterpret_cast<const float4*>(g + r);
                    d0 = fmaf(gv.x, a[r], d0);
                    d1 = fmaf(gv.y, a[r + 1], d1);
                    d2 = fmaf(gv.z, a[r + 2], d2);
                    d3 = fmaf(gv.w, a[r + 3], d3);
                }
            } else {
#pragma unroll
                for (int r = 0; r < S; ++r) d0 = fmaf(g[r], a[r], d0);
            }
            float dot = (d0 + d1) + (d2 + d3);
#pragma unroll
            for (int m = L / 2; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
            const float aq = a[q];
            const float grad = dot - bv[q];
            const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
            if (sInv) {
                const float inv = sInv[t];
                if (l == o && valid) {
                    viol += fabsf(pg);
                    if (pg_row) pg_row[t] = fabsf(pg);
                    if (inv != 0.f) a[q] = fmaxf(fmaf(-grad, inv, aq), 0.f);
                }
            } else {
                const float hess = sG[t * PITCH + o * (S + 4) + q];
                if (l == o && valid) {
                    viol += fabsf(pg);
                    if (pg_row) pg_row[t] = fabsf(pg);
                    if (hess != 0.f) a[q] = fmaxf(aq - grad / hess, 0.f);
                }
            }
        }
    }
    return viol;
}

// Copy a KP x KP Gram matrix from global memory into the padded shared layout described above.
template <int KP, int L>
__device__ __forceinline__ void load_gram_padded(float* sG, const float* __restrict__ G) {
    constexpr int S = KP / L;
    constexpr int PITCH = KP + 4 * L;
    for (int i = threadIdx.x; i < KP * KP / 4; i += blockDim.x) {
        const int t = (4 * i) / KP, r = (4 * i) % KP;      // r multiple of 4, S multiple of 4 -> same slice
        const float4 v = *reinterpret_cast<const float4*>(G + 4 * i);
        *reinterpret_cast<float4*>(sG + t * PITCH + (r / S) * (S + 4) + (r % S)) = v;
    }
}

}  // namespace ainmf

namespace ainmf {

// ---- the sweep, incremental-gradient form ------------------------------------------------------------------
// Same Gauss-Seidel update as cd_sweep_row, reorganised so that (i) no dot product is recomputed and (ii) every
// Gram element fetched from shared memory is used for R rows (the plain form needs one 4-byte shared load per FMA
// and is bound by the 128 B/clk shared-memory return path, not by the FMA pipe).
//   g = A.G - B is supplied for the OLD A (the tensor-core path produces it in the same accumulator as X^T.W);
//   for t = 0..KP-1:   grad = g[t]  ->  pg / violation / update exactly as in the reference
//                      delta = A_new[t] - A_old[t];   g[:] += delta * G[t,:]      (G symmetric)
// A group of L consecutive lanes owns R rows; lane l holds the slices a[i][q] = A[row_i][l*S+q], g[i][q] likewise.
// delta is broadcast inside the group with one shuffle per row; a warp whose deltas are all zero skips the rank-1
// update (frequent once the factors are sparse).  sG uses the padded layout of load_gram_padded<KP, L>.
template <int KP, int L, int R>
__device__ __forceinline__ float cd_sweep_rows_inc(float (&a)[R][KP / L], float (&g)[R][KP / L],
                                                   const float* __restrict__ sG, const float* __restrict__ sInv, int l,
                                                   const bool (&valid)[R]) {
    constexpr int S = KP / L;
    constexpr int PITCH = KP + 4 * L;
    float viol = 0.f;
    const float* gl = sG + l * (S + 4);
    for (int o = 0; o < L; ++o) {
#pragma unroll
        for (int q = 0; q < S; ++q) {
            const int t = o * S + q;
            const float inv = sInv[t];
            float delta[R];
            bool any = false;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                // branch-free: every lane evaluates the update, only the owner of coordinate t keeps it
                const float aq = a[i][q];
                const float grad = g[i][q];
                const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
                const float an = fmaxf(fmaf(-grad, inv, aq), 0.f);
                const bool own = (l == o) && valid[i];
                const bool upd = own && (inv != 0.f);
                viol += own ? fabsf(pg) : 0.f;
                const float d = upd ? an - aq : 0.f;
                a[i][q] = upd ? an : aq;
                delta[i] = __shfl_sync(0xffffffffu, d, o, L);
                any = any || (delta[i] != 0.f);
            }
            if (__ballot_sync(0xffffffffu, any) == 0u) continue;      // warp-uniform
            const float* gr = gl + t * PITCH;
#pragma unroll
            for (int r = 0; r < S; r += 4) {
                const float4 gv = *reinterpret_cast<const float4*>(gr + r);
#pragma unroll
                for (int i = 0; i < R; ++i) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(AINMF_EMU)
                    // packed fp32x2 FMA (sm_100): two gradient entries per issue slot
                    const float2 dd = make_float2(delta[i], delta[i]);
                    const float2 lo2 = __ffma2_rn(dd, make_float2(gv.x, gv.y), make_float2(g[i][r], g[i][r + 1]));
                    const float2 hi2 = __ffma2_rn(dd, make_float2(gv.z, gv.w), make_float2(g[i][r + 2], g[i][r + 3]));
                    g[i][r] = lo2.x; g[i][r + 1] = lo2.y; g[i][r + 2] = hi2.x; g[i][r + 3] = hi2.y;
#else
                    g[i][r] = fmaf(delta[i], gv.x, g[i][r]);
                    g[i][r + 1] = fmaf(delta[i], gv.y, g[i][r + 1]);
                    g[i][r + 2] = fmaf(delta[i], gv.z, g[i][r + 2]);
                    g[i][r + 3] = fmaf(delta[i], gv.w, g[i][r + 3]);
#endif
                }
            }
        }
    }
    return viol;
}

}  // namespace ainmf
