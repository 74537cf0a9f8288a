// nmf_small.cu -- the whole coordinate-descent fit of a SMALL spectrogram, and the chain of refits main4_NMF.py runs on it
// (main4_NMF.py:83-90: `for i in range(n_iter): W = model.fit_transform(cur); cur[:, cs:ce] = (W @ H)[:, cs:ce]`), as ONE
// launch: one CTA per clip keeps X, W, Ht, the two Grams and the two product matrices in shared memory for every iteration
// of every refit and tests the stop rule itself, so nothing goes back to the host or to another kernel in between.
//
// Why: config 1 of BASELINE.json is a 257 x 19 spectrogram at K = 40 refitted 50 times with early stopping -- 3181
// iterations.  Through the general path that is ~20 000 launches of kernels that each run a few microseconds on one or
// two SMs, plus a host poll of the stop flag every 8 iterations: 225 ms, 3x the CPU.  Here an iteration is ~7 us.
//
// Arithmetic per iteration is the reference's (sklearn _nmf.py:491-516, _cdnmf_fast.pyx:8-38): HHt, X.Ht, W sweep, WtW,
// Xt.W, H sweep, violation, `violation / violation_init <= tol`.  The violation is summed per thread in float32 and
// across threads in double in a fixed order; when the decision is close (and always for iteration 1) it is re-summed in
// sklearn's own order and precision from the per-coordinate |pg| values, exactly as stop_kernel does (nmf_cd.cu).
#include "kernels.h"
#include <stdio.h>
#include <stdlib.h>

namespace ainmf {

constexpr int kSmallThreads = 512;          // one thread per row in the sweeps (rows are taken 512 at a time), 4x4 tiles in the Grams
// Two things decide the speed here, both measured with the per-phase cycle counters (AINMF_SMALL_DEBUG=1):
//  * code size: one CTA runs alone on its SM for thousands of iterations, and a loop body that does not fit the 32 KB
//    instruction cache is re-fetched from L2 line by line on every pass (a fully unrolled 64 x 64 thread-per-row sweep, 380 KB
//    of SASS, ran at 31 cycles per instruction).  The sweep therefore unrolls only 8 coordinates and ROTATES the register
//    copy of the row by 8 after each group, so that the next 8 coordinates are again registers 0..7;
//  * issue slots: with 4 lanes per row and 1024 threads every warp executes the owner lane's update code (a division among
//    it) at every coordinate -- 2 560 warp instructions per coordinate step, 51 000 cycles per W sweep.  One thread per row
//    needs 9 warps for 257 rows.

struct SmallLayout {            // offsets in floats into dynamic shared memory
    int Kr, P, NG, oX, oW, oB1, oHt, oB2, oG1, oG2, oGp, total;
};
// Kr = rank rounded up to 16 (the shared arrays hold Kr coordinates, not the padded KP of the global layout; the kernel is
// instantiated per Kr); rows of the factor matrices are P = Kr + 4 floats apart (16-byte aligned, and float4 reads of one
// row per thread are conflict-free for Kr = 16, 48; 2-way for 32, 64); Grams are Kr x Kr; NG row groups share a Gram's rows
__host__ __device__ inline SmallLayout small_layout(int F, int T, int K) {
    SmallLayout L;
    L.Kr = (K + 15) / 16 * 16;
    L.P = L.Kr + 4;
    const int tiles = (L.Kr / 4) * (L.Kr / 4);
    L.NG = kSmallThreads / tiles < 1 ? 1 : (kSmallThreads / tiles > 4 ? 4 : kSmallThreads / tiles);
    int o = 0;
    auto take = [&](int n) { const int r = o; o += (n + 3) / 4 * 4; return r; };
    L.oX = take(T * F);            // X[t][f]
    L.oW = take(F * L.P);          // W[f][k]
    L.oB1 = take(F * L.P);         // X.Ht [f][k]
    L.oHt = take(T * L.P);         // Ht[t][k]
    L.oB2 = take(T * L.P);         // Xt.W [t][k]
    L.oG1 = take(L.Kr * L.Kr);     // HHt
    L.oG2 = take(L.Kr * L.Kr);     // WtW
    L.oGp = take(L.NG * L.Kr * L.Kr);   // Gram partials of the row groups
    L.total = o;
    return L;
}

// G[i][j] = sum_r A[r][i] * A[r][j] for A [rows][P] in shared memory (Kr columns); 4x4 register tiles, rows split over NG
// thread groups whose partials are added in group order (deterministic).  All threads call it; ends with a barrier.
__device__ __noinline__ void small_gram(const float* __restrict__ A, int rows, int Kr, int P, int NG, float* __restrict__ Gp,
                                           float* __restrict__ G) {
    const int nt = Kr / 4, tiles = nt * nt;
    const int g = threadIdx.x / tiles, tile = threadIdx.x % tiles;
    if (g < NG) {
        const int ti = tile / nt, tj = tile % nt;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const int per = (rows + NG - 1) / NG, r0 = g * per, r1 = min(rows, r0 + per);
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(A + r * P + 4 * ti);
            const float4 b = *reinterpret_cast<const float4*>(A + r * P + 4 * tj);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(Gp + g * Kr * Kr + (4 * ti + i) * Kr + 4 * tj) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Kr * Kr; e += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < NG; ++q) s += Gp[q * Kr * Kr + e];
        G[e] = s;
    }
    __syncthreads();
}

// One coordinate sweep of the rows of A [rows][P], the reference's formulation (_cdnmf_fast.pyx:8-38): for t = 0..K-1:
// grad = -B[row,t] + G[t,:].A[row,:]; pg = A[row,t] == 0 ? min(0, grad) : grad; A[row,t] <- max(A[row,t] - grad / G[t,t], 0).
// Thread = row, the row in registers a[KR].  Coordinates go in groups of 8 (rolled loop); within a group register j IS
// coordinate 8c + j, and register i holds column (i + 8c) mod KR -- after the group the array is rotated by 8, so the
// unrolled code always updates registers 0..7 and only the shared-memory column offset of the Gram rows depends on c.
// |pg| replaces B[row, t] (read just before, not needed again) for the reference-order sum.  Returns the thread's violation.
template <int KR>
__device__ __noinline__ float small_sweep(float* __restrict__ A, float* __restrict__ Bm, const float* __restrict__ G, int rows, int K, int P) {
    float viol = 0.f;
    for (int row = threadIdx.x; row < rows; row += kSmallThreads) {
        float a[KR];
        float* ar = A + row * P;
#pragma unroll
        for (int q = 0; q < KR; q += 4) {
            const float4 v = *reinterpret_cast<const float4*>(ar + q);
            a[q] = v.x; a[q + 1] = v.y; a[q + 2] = v.z; a[q + 3] = v.w;
        }
        float* br = Bm + row * P;
#pragma unroll 1
        for (int c = 0; c < KR / 8; ++c) {
            const int c8 = 8 * c;
            if (c8 < K) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int t = c8 + j;
                    if (t >= K) break;                                      // uniform: pad coordinates have zero gradient
                    const float* gr = G + t * KR;
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int i = 0; i < KR; i += 4) {
                        int col = i + c8;                                   // column held by registers i..i+3
                        col -= (col >= KR) ? KR : 0;
                        const float4 gv = *reinterpret_cast<const float4*>(gr + col);
                        d0 = fmaf(gv.x, a[i], d0); d1 = fmaf(gv.y, a[i + 1], d1);
                        d2 = fmaf(gv.z, a[i + 2], d2); d3 = fmaf(gv.w, a[i + 3], d3);
                    }
                    const float grad = ((d0 + d1) + (d2 + d3)) - br[t];
                    const float hess = gr[t];
                    const float aq = a[j];
                    const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
                    viol += fabsf(pg);
                    br[t] = fabsf(pg);
                    if (hess != 0.f) a[j] = fmaxf(aq - grad / hess, 0.f);
                }
            }
            float t8[8];                                                    // rotate by one group
#pragma unroll
            for (int j = 0; j < 8; ++j) t8[j] = a[j];
#pragma unroll
            for (int i = 0; i + 8 < KR; ++i) a[i] = a[i + 8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[KR - 8 + j] = t8[j];
        }
#pragma unroll
        for (int q = 0; q < KR; q += 4) *reinterpret_cast<float4*>(ar + q) = make_float4(a[q], a[q + 1], a[q + 2], a[q + 3]);
    }
    return viol;
}

// sklearn's own violation (see nmf_cd.cu: violation_in_reference_order): float32 running sums, coordinates outermost, rows
// innermost, from the |pg| values the sweeps left in the two product matrices.  One thread; ~4 cycles per addition.
static __device__ double small_violation_reference_order(const float* __restrict__ pW, int F, const float* __restrict__ pH, int T, int K, int P) {
    float vw = 0.f, vh = 0.f;
    for (int t = 0; t < K; ++t) {
        int i = 0;
        for (; i + 8 <= F; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = pW[(i + u) * P + t];
#pragma unroll
            for (int u = 0; u < 8; ++u) vw += v[u];
        }
        for (; i < F; ++i) vw += pW[i * P + t];
    }
    for (int t = 0; t < K; ++t)
        for (int i = 0; i < T; ++i) vh += pH[i * P + t];
    return (double)vw + (double)vh;
}

template <int KR>
__global__ void __launch_bounds__(kSmallThreads, 1)
nmf_small_refit_kernel(float* __restrict__ Xt, long long x_stride, int ldf, int F, int T, int K, int KP, float* __restrict__ Wg, long long w_stride,
                       float* __restrict__ Htg, long long h_stride, const float* __restrict__ Wn /*[F][K]*/, const float* __restrict__ Hn /*[K][T]*/,
                       const unsigned char* __restrict__ bad, long long bad_stride, ClipState* __restrict__ st, int n_outer, int max_iter,
                       float tol, long long* __restrict__ dbg /*AINMF_SMALL_DEBUG=1: cycles per phase, thread 0 of clip 0*/) {
    AINMF_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    const SmallLayout L = small_layout(F, T, K);
    constexpr int Kr = KR;
    const int P = L.P;
    float *sX = sm + L.oX, *sW = sm + L.oW, *sB1 = sm + L.oB1, *sHt = sm + L.oHt, *sB2 = sm + L.oB2, *sG1 = sm + L.oG1, *sG2 = sm + L.oG2,
          *sGp = sm + L.oGp;
    __shared__ double s_red[32];
    __shared__ double s_v;
    __shared__ int s_stop;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (st[b].status != 0) return;                         // nothing to restore / undefined fill: whole block leaves together
    float* Xb = Xt + (long long)b * x_stride;
    for (int i = tid; i < T * F; i += blockDim.x) sX[i] = Xb[(long long)(i / F) * ldf + (i % F)];
    for (int i = tid; i < F * P; i += blockDim.x) sB1[i] = 0.f;            // pad coordinates stay zero for the whole run
    for (int i = tid; i < T * P; i += blockDim.x) sB2[i] = 0.f;
    const int K4 = Kr / 4;                                                 // groups of 4 coordinates
    __syncthreads();
    int n_iter = 0;
    long long c_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c_t = 0;
    const bool dbg_on = dbg != nullptr && b == 0 && tid == 0;
#define SM_TIC() do { if (dbg_on) c_t = clock64(); } while (0)
#define SM_TOC(k) do { if (dbg_on) { const long long c_ = clock64(); c_ph[k] += c_ - c_t; c_t = c_; } } while (0)
    double viol_init = 0.0, viol_last = 0.0;
    float err = 0.f;
    for (int outer = 0; outer < n_outer; ++outer) {
        // ---- initial factors of this fit: avg = sqrt(mean(X) / K); W = |avg * N(0,1)|, H likewise (_nmf.py:296-307) ----
        double part = 0.0;
        for (int i = tid; i < T * F; i += blockDim.x) part += (double)sX[i];
        const double sum_x = block_sum_d(part, s_red);
        const float mean_x = (float)(sum_x / ((double)F * (double)T));
        const float avg = sqrtf(mean_x / (float)K);
        for (int i = tid; i < F * Kr; i += blockDim.x) {
            const int f = i / Kr, k = i % Kr;
            sW[f * P + k] = (k < K) ? fabsf(avg * Wn[(long long)f * K + k]) : 0.f;
        }
        for (int i = tid; i < T * Kr; i += blockDim.x) {
            const int t = i / Kr, k = i % Kr;
            sHt[t * P + k] = (k < K) ? fabsf(avg * Hn[(long long)k * T + t]) : 0.f;
        }
        __syncthreads();
        n_iter = 0;
        for (int it = 1; it <= max_iter; ++it) {
            // ---- W half: HHt, X.Ht, sweep ----
            SM_TIC();
            small_gram(sHt, T, Kr, P, L.NG, sGp, sG1);
            SM_TOC(0);
            for (int i = tid; i < F * K4; i += blockDim.x) {                 // thread: row f, 4 consecutive k (pads of a group rewritten as 0)
                const int f = i % F, k4 = (i / F) * 4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int t = 0; t < T; ++t) {
                    const float x = sX[t * F + f];
                    const float4 h = *reinterpret_cast<const float4*>(sHt + t * P + k4);
                    acc.x = fmaf(x, h.x, acc.x); acc.y = fmaf(x, h.y, acc.y); acc.z = fmaf(x, h.z, acc.z); acc.w = fmaf(x, h.w, acc.w);
                }
                *reinterpret_cast<float4*>(sB1 + f * P + k4) = acc;
            }
            __syncthreads();
            SM_TOC(1);
            float v = small_sweep<KR>(sW, sB1, sG1, F, K, P);
            __syncthreads();
            SM_TOC(2);
            // ---- H half: WtW, Xt.W, sweep ----
            small_gram(sW, F, Kr, P, L.NG, sGp, sG2);
            SM_TOC(3);
            for (int i = tid; i < T * K4; i += blockDim.x) {                  // thread: frame t, 4 consecutive k
                const int k4 = (i % K4) * 4, t = i / K4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* xr = sX + t * F;
#pragma unroll 8
                for (int f = 0; f < F; ++f) {
                    const float x = xr[f];
                    const float4 w = *reinterpret_cast<const float4*>(sW + f * P + k4);
                    acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
                }
                *reinterpret_cast<float4*>(sB2 + t * P + k4) = acc;
            }
            __syncthreads();
            SM_TOC(4);
            v += small_sweep<KR>(sHt, sB2, sG2, T, K, P);
            SM_TOC(5);
            // ---- stop rule ----
            double vd = block_sum_d((double)v, s_red);
            __syncthreads();                                                  // the |pg| values of both sweeps are in place
            if (tid == 0) {
                bool exact = (it == 1);
                if (!exact && viol_init != 0.0) {
                    const double r = vd / viol_init, band = 2e-3 * (double)tol;
                    exact = (r > (double)tol - band) && (r < (double)tol + band);
                }
                if (exact) vd = small_violation_reference_order(sB1, F, sB2, T, K, P);
                s_v = vd;
            }
            __syncthreads();
            SM_TOC(6);
            vd = s_v;
            n_iter = it;
            if (dbg_on) c_ph[7] += 1;
            if (it == 1) viol_init = vd;
            viol_last = vd;
            if (viol_init == 0.0 || vd / viol_init <= (double)tol) break;
        }
        // ---- objective ||X - W H||_F (all frames), then bad frames <- (W H) frames (_nmf.py:1623; main4_NMF.py:89-90) ----
        double e2 = 0.0;
        for (int i = tid; i < T * F; i += blockDim.x) {
            const int t = i / F, f = i % F;
            float d = 0.f;
            for (int k = 0; k < Kr; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(sW + f * P + k);
                const float4 h = *reinterpret_cast<const float4*>(sHt + t * P + k);
                d = fmaf(w.x, h.x, d); d = fmaf(w.y, h.y, d); d = fmaf(w.z, h.z, d); d = fmaf(w.w, h.w, d);
            }
            const float r = sX[i] - d;
            e2 += (double)r * (double)r;
            if (bad[(long long)b * bad_stride + t]) sX[i] = d;
        }
        e2 = block_sum_d(e2, s_red);
        err = (float)sqrt(e2);
        __syncthreads();
    }
    // ---- results: restored frames, factors, state ----
    for (int i = tid; i < T * F; i += blockDim.x) {
        const int t = i / F, f = i % F;
        if (bad[(long long)b * bad_stride + t]) Xb[(long long)t * ldf + f] = sX[i];
    }
    for (int i = tid; i < F * KP; i += blockDim.x) Wg[(long long)b * w_stride + i] = (i % KP < Kr) ? sW[(i / KP) * P + (i % KP)] : 0.f;
    for (int i = tid; i < T * KP; i += blockDim.x) Htg[(long long)b * h_stride + i] = (i % KP < Kr) ? sHt[(i / KP) * P + (i % KP)] : 0.f;
    if (dbg_on) for (int i = 0; i < 8; ++i) dbg[i] = c_ph[i];
    if (tid == 0) {
        ClipState s = st[b];
        s.n_iter = n_iter;
        s.viol_init = viol_init;
        s.viol_last = viol_last;
        s.err = err;
        s.done = 1;
        st[b] = s;
    }
}

bool nmf_small_eligible(int F, int T, int K) {
    if (K > 64) return false;
    return (size_t)small_layout(F, T, K).total * sizeof(float) <= 220 * 1024;
}

template <int KR>
static cudaError_t small_launch(const NmfProblem& p, int K, const float* Wn, const float* Hn, const unsigned char* bad, long long bad_stride,
                                int n_outer, int max_iter, cudaStream_t s) {
    const size_t smem = (size_t)small_layout(p.F, p.T, K).total * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(nmf_small_refit_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static long long* dbg = nullptr;
    static int dbg_left = -1;
    if (dbg_left < 0) {
        const char* e_ = getenv("AINMF_SMALL_DEBUG");
        dbg_left = (e_ && e_[0] == '1') ? 1 : 0;
        if (dbg_left) cudaMalloc((void**)&dbg, 8 * sizeof(long long));
    }
    AINMF_LAUNCH(nmf_small_refit_kernel<KR>, dim3(p.B), dim3(kSmallThreads), smem, s, p.Xt, p.x_stride, p.ldf, p.F, p.T, K, p.KP, p.W, p.w_stride,
                 p.Ht, p.h_stride, Wn, Hn, bad, bad_stride, p.state, n_outer, max_iter, p.tol, dbg_left > 0 ? dbg : nullptr);
    if (dbg_left > 0) {
        --dbg_left;
        long long hb[8];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, dbg, sizeof hb, cudaMemcpyDeviceToHost);
        const double n = hb[7] > 0 ? (double)hb[7] : 1.0;
        fprintf(stderr, "[small-debug F=%d T=%d K=%d] %lld iterations; cycles per iteration: gram(Ht) %.0f, X.Ht %.0f, W sweep %.0f, gram(W) %.0f, Xt.W %.0f, "
                        "H sweep %.0f, violation + stop %.0f\n", p.F, p.T, K, hb[7], hb[0] / n, hb[1] / n, hb[2] / n, hb[3] / n, hb[4] / n, hb[5] / n, hb[6] / n);
    }
    return cudaGetLastError();
}

cudaError_t nmf_small_refit(const NmfProblem& p, int K, const float* Wn, const float* Hn, const unsigned char* bad, long long bad_stride,
                            int n_outer, int max_iter, cudaStream_t s) {
    switch ((K + 15) / 16) {
        case 1: return small_launch<16>(p, K, Wn, Hn, bad, bad_stride, n_outer, max_iter, s);
        case 2: return small_launch<32>(p, K, Wn, Hn, bad, bad_stride, n_outer, max_iter, s);
        case 3: return small_launch<48>(p, K, Wn, Hn, bad, bad_stride, n_outer, max_iter, s);
        default: return small_launch<64>(p, K, Wn, Hn, bad, bad_stride, n_outer, max_iter, s);
    }
}

}  // namespace ainmf
