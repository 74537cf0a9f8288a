// nmf_cd.cu -- K4: one coordinate-descent NMF iteration as FFMA kernels for sm_100a.
//
//   X  = imputed magnitude spectrogram, frame-major Xt[b][t][f] (ldf)          (main4_NMF_gap.py:56-59)
//   W  [b][f][KP],  Ht [b][t][KP]                                              (KP = padded rank)
// per iteration ($SP/sklearn/decomposition/_nmf.py:491-516):
//   HHt = Ht^T Ht           gram_kernel            (deterministic two-stage reduction)
//   XHt = X Ht              xht_kernel             (split over time, partials reduced in fixed order)
//   W   <- sweep            w_sweep_kernel
//   WtW = W^T W             gram_kernel
//   XtW = X^T W ; Ht <- sweep   h_step_kernel      (fused: the XtW tile never leaves the SM)
//   stop rule               stop_kernel
// The spectrogram is read exactly twice per iteration (xht_kernel, h_step_kernel): 8*F*T bytes.
#include "kernels.h"
#include "nmf_cd.cuh"
#include "nmf_ts.cuh"

#include <stdlib.h>

#include <vector>

namespace ainmf {

unsigned long long g_launch_count = 0;

// ---- profiling --------------------------------------------------------------------------------------------
namespace {
struct ProfRec { int kind; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t g_prof_open[PROF_KINDS];
cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
}  // namespace
void prof_enable(bool on) { g_prof_on = on; }
void prof_begin(int kind, cudaStream_t s) {
    if (!g_prof_on) return;
    g_prof_open[kind] = prof_event();
    cudaEventRecord(g_prof_open[kind], s);
}
void prof_end(int kind, cudaStream_t s) {
    if (!g_prof_on) return;
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, s);
    g_prof_recs.push_back({kind, g_prof_open[kind], e});
}
void prof_collect(double* ms, long long* counts) {
    for (int i = 0; i < PROF_KINDS; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (const ProfRec& r : g_prof_recs) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.kind] += (double)t;
        counts[r.kind] += 1;
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
}

// =====================================================================================================
// gram: G[b] = A[b]^T A[b] for A [rows][KP]; grid = (P, B); partials then last-block reduce.
// =====================================================================================================
template <int KP>
__global__ void __launch_bounds__(kThreads)
gram_kernel(const float* __restrict__ A, long long a_stride, int rows, int rows_per_block,
            float* __restrict__ partial /*[B][P][KP*KP]*/, float* __restrict__ G /*[B][KP*KP]*/,
            unsigned* __restrict__ counters /*[B]*/, const ClipState* __restrict__ st) {
    constexpr int TN = KP / 16;
    constexpr int RB = 32;                               // rows staged per step
    __shared__ __align__(16) float sA[RB][KP];
    __shared__ unsigned s_last;
    const int b = blockIdx.y, p = blockIdx.x, P = gridDim.x;
    if (st[b].done) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* Ab = A + (long long)b * a_stride;
    const int r_begin = p * rows_per_block;
    const int r_end = min(rows, r_begin + rows_per_block);
    float acc[TN][TN];
#pragma unroll
    for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += RB) {
        for (int i = threadIdx.x; i < RB * KP / 4; i += blockDim.x) {
            const int rr = (4 * i) / KP, cc = (4 * i) % KP;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + rr < r_end) v = *reinterpret_cast<const float4*>(Ab + (long long)(r0 + rr) * KP + cc);
            *reinterpret_cast<float4*>(&sA[rr][cc]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < RB; ++kk) {
            float fi[TN], fj[TN];
            load_frag<TN>(sA[kk], ty, fi);
            load_frag<TN>(sA[kk], tx, fj);
#pragma unroll
            for (int i = 0; i < TN; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(fi[i], fj[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = partial + ((long long)b * P + p) * (KP * KP);
#pragma unroll
    for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) out[frag_col<TN>(ty, i) * KP + frag_col<TN>(tx, j)] = acc[i][j];

    // last block of this clip sums the partials in index order (deterministic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&counters[b], 1u) == (unsigned)(P - 1)) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float* pb = partial + (long long)b * P * (KP * KP);
        for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
            float s = 0.f;
            for (int q = 0; q < P; ++q) s += pb[(long long)q * (KP * KP) + e];
            G[(long long)b * (KP * KP) + e] = s;
        }
        if (threadIdx.x == 0) counters[b] = 0u;          // ready for the next launch
    }
}

// =====================================================================================================
// xht: partial[b][s][f][:] = sum_{t in split s} Xt[t][f] * Ht[t][:];  grid = (ceil(F/128), S, B)
// =====================================================================================================
template <int KP>
__global__ void __launch_bounds__(kThreads)
xht_kernel(const float* __restrict__ Xt, long long x_stride, int ldf, int F, int T,
           const float* __restrict__ Ht, long long h_stride, int frames_per_split,
           float* __restrict__ partial /*[B][S][F][KP]*/, const ClipState* __restrict__ st) {
    constexpr int TN = KP / 16;
    constexpr int BM = 128, BK = 16;
    constexpr int A4 = BK * BM / 4 / kThreads;            // float4 per thread for the A tile (=2)
    constexpr int B4 = (BK * KP / 4 + kThreads - 1) / kThreads;
    __shared__ __align__(16) float sA[BK][BM];
    __shared__ __align__(16) float sB[BK][KP];
    const int b = blockIdx.z, split = blockIdx.y, S = gridDim.y;
    if (st[b].done) return;
    const int f0 = blockIdx.x * BM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int warp = threadIdx.x >> 5;
    const bool act0 = (f0 + 8 * warp) < F;                // rows 4ty..4ty+3
    const bool act1 = (f0 + 64 + 8 * warp) < F;           // rows 64+4ty..
    const float* Xb = Xt + (long long)b * x_stride;
    const float* Hb = Ht + (long long)b * h_stride;
    const int t_begin = split * frames_per_split;
    const int t_end = min(T, t_begin + frames_per_split);

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A4], rb[B4];
    auto gload = [&](int t0) {
#pragma unroll
        for (int u = 0; u < A4; ++u) {
            const int i = threadIdx.x + u * kThreads;     // float4 index in the [BK][BM] tile
            const int kk = i / (BM / 4), m = (i % (BM / 4)) * 4;
            ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t0 + kk < t_end && f0 + m < ldf)
                ra[u] = *reinterpret_cast<const float4*>(Xb + (long long)(t0 + kk) * ldf + f0 + m);
        }
#pragma unroll
        for (int u = 0; u < B4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < BK * KP / 4) {
                const int kk = i / (KP / 4), n = (i % (KP / 4)) * 4;
                if (t0 + kk < t_end) rb[u] = *reinterpret_cast<const float4*>(Hb + (long long)(t0 + kk) * KP + n);
            }
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int u = 0; u < A4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            *reinterpret_cast<float4*>(&sA[i / (BM / 4)][(i % (BM / 4)) * 4]) = ra[u];
        }
#pragma unroll
        for (int u = 0; u < B4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            if (i < BK * KP / 4) *reinterpret_cast<float4*>(&sB[i / (KP / 4)][(i % (KP / 4)) * 4]) = rb[u];
        }
    };

    if (t_begin < t_end) gload(t_begin);
    for (int t0 = t_begin; t0 < t_end; t0 += BK) {
        sstore();
        __syncthreads();
        if (t0 + BK < t_end) gload(t0 + BK);               // prefetch the next tile into registers
        if (act0 || act1) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float fb[TN];
                load_frag<TN>(sB[kk], tx, fb);
                if (act0) {
                    const float4 a = *reinterpret_cast<const float4*>(&sA[kk][4 * ty]);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], fb[j], acc[i][j]);
                }
                if (act1) {
                    const float4 a = *reinterpret_cast<const float4*>(&sA[kk][64 + 4 * ty]);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[4 + i][j] = fmaf(av[i], fb[j], acc[4 + i][j]);
                }
            }
        }
        __syncthreads();
    }
    float* out = partial + (((long long)b * S + split) * F) * KP;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int f = f0 + ((i < 4) ? 4 * ty + i : 64 + 4 * ty + (i - 4));
        if (f < F) {
#pragma unroll
            for (int j = 0; j < TN; ++j) out[(long long)f * KP + frag_col<TN>(tx, j)] = acc[i][j];
        }
    }
}

// =====================================================================================================
// h step: XtW tile = Xt[tile] W ; Ht[tile] <- sweep(Ht[tile], WtW, XtW tile); grid = (ceil(T/BM), B)
// =====================================================================================================
template <int KP, int BM> struct HStepCfg {
    static constexpr int L = (BM == 32) ? 8 : 4;
    static constexpr int BK = 32;
    static constexpr int APITCH = BK + 4;
    static constexpr int CPITCH = KP + 4;
    static constexpr int GPITCH = KP + 4 * L;
    static constexpr int gemm_floats = BM * APITCH + BK * KP;
    static constexpr int c_floats = BM * CPITCH;
    static constexpr int work_floats = gemm_floats > c_floats ? gemm_floats : c_floats;
    static constexpr size_t smem_bytes = sizeof(float) * (size_t)(work_floats + KP * GPITCH);
};

template <int KP, int BM>
__global__ void __launch_bounds__(kThreads)
h_step_kernel(const float* __restrict__ Xt, long long x_stride, int ldf, int F, int T,
              const float* __restrict__ W, long long w_stride, const float* __restrict__ G,
              float* __restrict__ Ht, long long h_stride, float* __restrict__ viol /*[B][gridDim.x]*/,
              const ClipState* __restrict__ st, float* __restrict__ xtw_out /*MU solver: store X^T.W, no sweep*/,
              float* __restrict__ pg_out /*[B][T][KP] or null*/) {
    using Cfg = HStepCfg<KP, BM>;
    constexpr int TN = KP / 16, TM = BM / 16, L = Cfg::L, SL = KP / L, BK = Cfg::BK;
    constexpr int APITCH = Cfg::APITCH, CPITCH = Cfg::CPITCH;
    constexpr int A4 = (BM * BK / 4 + kThreads - 1) / kThreads;
    constexpr int B4 = (BK * KP / 4 + kThreads - 1) / kThreads;
    AINMF_DYN_SMEM(smem_raw);
    float* sA = reinterpret_cast<float*>(smem_raw);        // [BM][APITCH]
    float* sB = sA + BM * APITCH;                          // [BK][KP]
    float* sC = sA;                                        // [BM][CPITCH]   (aliases the GEMM tiles)
    float* sG = sA + Cfg::work_floats;                     // [KP][GPITCH]
    __shared__ float s_red[32];
    const int b = blockIdx.y;
    if (st[b].done) return;
    const int m0 = blockIdx.x * BM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* Xb = Xt + (long long)b * x_stride;
    const float* Wb = W + (long long)b * w_stride;
    load_gram_padded<KP, L>(sG, G + (long long)b * KP * KP);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A4], rb[B4];
    auto gload = [&](int k0) {
#pragma unroll
        for (int u = 0; u < A4; ++u) {
            const int i = threadIdx.x + u * kThreads;       // float4 index in the [BM][BK] tile
            ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < BM * BK / 4) {
                const int m = i / (BK / 4), kk = (i % (BK / 4)) * 4;
                if (m0 + m < T && k0 + kk < ldf)
                    ra[u] = *reinterpret_cast<const float4*>(Xb + (long long)(m0 + m) * ldf + k0 + kk);
            }
        }
#pragma unroll
        for (int u = 0; u < B4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < BK * KP / 4) {
                const int kk = i / (KP / 4), n = (i % (KP / 4)) * 4;
                if (k0 + kk < F) rb[u] = *reinterpret_cast<const float4*>(Wb + (long long)(k0 + kk) * KP + n);
            }
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int u = 0; u < A4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            if (i < BM * BK / 4) *reinterpret_cast<float4*>(sA + (i / (BK / 4)) * APITCH + (i % (BK / 4)) * 4) = ra[u];
        }
#pragma unroll
        for (int u = 0; u < B4; ++u) {
            const int i = threadIdx.x + u * kThreads;
            if (i < BK * KP / 4) *reinterpret_cast<float4*>(sB + (i / (KP / 4)) * KP + (i % (KP / 4)) * 4) = rb[u];
        }
    };

    gload(0);
    for (int k0 = 0; k0 < F; k0 += BK) {
        sstore();
        __syncthreads();
        if (k0 + BK < F) gload(k0 + BK);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            float av[TM][4];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(sA + (ty + 16 * i) * APITCH + k4);
                av[i][0] = a.x; av[i][1] = a.y; av[i][2] = a.z; av[i][3] = a.w;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float fb[TN];
                load_frag<TN>(sB + (k4 + c) * KP, tx, fb);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i][c], fb[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    // XtW tile -> shared (row = frame within tile)
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) sC[(ty + 16 * i) * CPITCH + frag_col<TN>(tx, j)] = acc[i][j];
    __syncthreads();
    if (xtw_out) {      // whole block: the multiplicative-update solver wants the plain product
        float* ob = xtw_out + (long long)b * h_stride;
        for (int i = threadIdx.x; i < BM * KP / 4; i += blockDim.x) {
            const int r = (4 * i) / KP, c = (4 * i) % KP;
            if (m0 + r < T) *reinterpret_cast<float4*>(ob + (long long)(m0 + r) * KP + c) = *reinterpret_cast<const float4*>(sC + r * CPITCH + c);
        }
        return;
    }

    constexpr int ROWS = kThreads / L;                      // rows swept per pass
    const int l = threadIdx.x % L;
    float vsum = 0.f;
    float* Hb = Ht + (long long)b * h_stride;
    for (int r0 = 0; r0 < BM; r0 += ROWS) {
        const int r = r0 + threadIdx.x / L;
        const int t = m0 + r;
        const bool valid = (r < BM) && (t < T);
        float a[SL], bv[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) { a[q] = 0.f; bv[q] = 0.f; }
        if (valid) {
            const float* hr = Hb + (long long)t * KP + l * SL;
            const float* cr = sC + r * CPITCH + l * SL;
#pragma unroll
            for (int q = 0; q < SL; q += 4) {
                const float4 v = *reinterpret_cast<const float4*>(hr + q);
                a[q] = v.x; a[q + 1] = v.y; a[q + 2] = v.z; a[q + 3] = v.w;
                const float4 c = *reinterpret_cast<const float4*>(cr + q);
                bv[q] = c.x; bv[q + 1] = c.y; bv[q + 2] = c.z; bv[q + 3] = c.w;
            }
        }
        vsum += cd_sweep_row<KP, L>(a, bv, sG, l, valid, nullptr, (pg_out && valid) ? pg_out + ((long long)b * T + t) * KP : nullptr);
        if (valid) {
            float* hr = Hb + (long long)t * KP + l * SL;
#pragma unroll
            for (int q = 0; q < SL; q += 4) *reinterpret_cast<float4*>(hr + q) = make_float4(a[q], a[q + 1], a[q + 2], a[q + 3]);
        }
    }
    const float tot = block_sum(vsum, s_red);
    if (threadIdx.x == 0) viol[(long long)b * gridDim.x + blockIdx.x] = tot;
}

// =====================================================================================================
// stop rule: one warp per clip.  it is 1-based.
// =====================================================================================================
// sklearn's own violation: a float32 accumulator per half-step, coordinates outermost, rows innermost
// (_cdnmf_fast.pyx:8-38: `violation += fabs(pg)` inside `for s ... for i ...`), the two halves added as Python floats
// (_nmf.py:505-509).  One thread; pad coordinates contribute exact zeros.
static __device__ double violation_in_reference_order(const float* __restrict__ pgW, int F, const float* __restrict__ pgH, int T, int KP) {
    float vw = 0.f, vh = 0.f;
    for (int t = 0; t < KP; ++t) {
        int i = 0;
        for (; i + 8 <= F; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = pgW[(long long)(i + u) * KP + t];
#pragma unroll
            for (int u = 0; u < 8; ++u) vw += v[u];
        }
        for (; i < F; ++i) vw += pgW[(long long)i * KP + t];
    }
    for (int t = 0; t < KP; ++t) {
        int i = 0;
        for (; i + 8 <= T; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = pgH[(long long)(i + u) * KP + t];
#pragma unroll
            for (int u = 0; u < 8; ++u) vh += v[u];
        }
        for (; i < T; ++i) vh += pgH[(long long)i * KP + t];
    }
    return (double)vw + (double)vh;
}

__global__ void __launch_bounds__(kThreads)
stop_kernel(ClipState* __restrict__ st, int B, const float* __restrict__ violW, int nW,
            const float* __restrict__ violH, int nH, const double* __restrict__ extra, const float* __restrict__ extra_pack, int it, float tol,
            const float* __restrict__ pgW, const float* __restrict__ pgH, int F, int T, int KP) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    if (st[b].done) return;
    double v = 0.0;
    for (int i = lane; i < nW; i += 32) v += (double)violW[(long long)b * nW + i];
    for (int i = lane; i < nH; i += 32) v += (double)violH[(long long)b * nH + i];
    v = warp_sum_d(v);
    if (extra_pack) v += (double)extra_pack[4 * b] + (double)extra_pack[4 * b + 1];   // time-sharded mode: the H-side violation summed
    else if (extra) v += extra[b];                                                    // over the ranks (hi + lo floats, or a double)
    if (lane == 0) {
        ClipState s = st[b];
        if (pgW) {
            // Small problems: the reference's float32 running sum carries ~1e-5 of rounding noise, enough to move the
            // iteration at which v / v1 <= tol first holds when the ratio passes close to tol -- and in main4_NMF.py's 50
            // chained refits one moved stop changes everything after it.  v1 always, and any v whose ratio lies within
            // 0.2 % of tol (the worst case of the float32 sum is n.u = 7e-4), is therefore summed in the reference's
            // order and precision; far from tol the decision cannot depend on it.
            bool exact = (it == 1);
            if (!exact && s.viol_init != 0.0) {
                const double r = v / s.viol_init, band = 2e-3 * (double)tol;
                exact = (r > (double)tol - band) && (r < (double)tol + band);
            }
            if (exact) v = violation_in_reference_order(pgW + (long long)b * F * KP, F, pgH + (long long)b * T * KP, T, KP);
        }
        s.n_iter = it;
        if (it == 1) s.viol_init = v;
        s.viol_last = v;
        if (s.viol_init == 0.0) s.done = 1;
        else if (v / s.viol_init <= (double)tol) s.done = 1;
        st[b] = s;
    }
}

// =====================================================================================================
// finalize: err^2 partial = sum (X - W Ht^T)^2 over a tile of frames; bad frames <- (W Ht^T) row
// (a8 + a9: _nmf.py:1623 and main4_NMF_gap.py:65-68).  grid = (ceil(T/kFinalizeFrames), B)
// =====================================================================================================
constexpr int kFinalizeFrames = 16;
template <int KP>
__global__ void __launch_bounds__(kThreads)
finalize_kernel(float* __restrict__ Xt, long long x_stride, int ldf, int F, int T,
                const float* __restrict__ W, long long w_stride, const float* __restrict__ Ht,
                long long h_stride, const unsigned char* __restrict__ bad, long long bad_stride,
                double* __restrict__ err_partial /*[B][gridDim.x]*/) {
    constexpr int FR = kFinalizeFrames;
    __shared__ __align__(16) float sH[FR][KP];
    __shared__ unsigned char s_bad[FR];
    __shared__ double s_red[32];
    const int b = blockIdx.y, t0 = blockIdx.x * FR;
    const float* Hb = Ht + (long long)b * h_stride;
    for (int i = threadIdx.x; i < FR * KP; i += blockDim.x) {
        const int r = i / KP, c = i % KP;
        sH[r][c] = (t0 + r < T) ? Hb[(long long)(t0 + r) * KP + c] : 0.f;
    }
    if (threadIdx.x < FR) s_bad[threadIdx.x] = (t0 + threadIdx.x < T) ? bad[(long long)b * bad_stride + t0 + threadIdx.x] : 0;
    __syncthreads();
    float* Xb = Xt + (long long)b * x_stride;
    const float* Wb = W + (long long)b * w_stride;
    double e2 = 0.0;
    // Two bins per thread share every H load (FMA : LDS.128 = 8 : 1); the bins past the last full round of
    // blockDim.x (one bin for F = 2^m + 1) are spread over (bin, frame) pairs instead of idling all but a few threads.
    const int Fmain = (F / (int)blockDim.x) * (int)blockDim.x;
    for (int fb = 0; fb < Fmain; fb += 2 * blockDim.x) {
        const int f0 = fb + threadIdx.x, f1 = f0 + blockDim.x;
        const bool has1 = f1 < Fmain;                              // uniform across the block
        float d0[FR], d1[FR];
#pragma unroll
        for (int r = 0; r < FR; ++r) { d0[r] = 0.f; d1[r] = 0.f; }
        const float* w0r = Wb + (long long)f0 * KP;
        const float* w1r = Wb + (long long)(has1 ? f1 : f0) * KP;
        // X is read before the products so that its DRAM latency hides behind them (the stores to bad frames
        // below would otherwise order every load after the previous frame's store)
        float x0[FR], x1[FR];
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            const bool ok = t0 + r < T;
            const long long o = (long long)(t0 + r) * ldf + f0;
            x0[r] = ok ? Xb[o] : 0.f;
            x1[r] = (ok && has1) ? Xb[o + blockDim.x] : 0.f;
        }
        for (int k = 0; k < KP; k += 4) {
            const float4 w0 = *reinterpret_cast<const float4*>(w0r + k);
            const float4 w1 = *reinterpret_cast<const float4*>(w1r + k);
#pragma unroll
            for (int r = 0; r < FR; ++r) {
                const float4 h = *reinterpret_cast<const float4*>(&sH[r][k]);
                d0[r] = fmaf(w0.x, h.x, d0[r]);
                d1[r] = fmaf(w1.x, h.x, d1[r]);
                d0[r] = fmaf(w0.y, h.y, d0[r]);
                d1[r] = fmaf(w1.y, h.y, d1[r]);
                d0[r] = fmaf(w0.z, h.z, d0[r]);
                d1[r] = fmaf(w1.z, h.z, d1[r]);
                d0[r] = fmaf(w0.w, h.w, d0[r]);
                d1[r] = fmaf(w1.w, h.w, d1[r]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < FR; ++r) {
            if (t0 + r < T) {
                const long long o = (long long)(t0 + r) * ldf + f0;
                const float d = x0[r] - d0[r];
                s = fmaf(d, d, s);
                if (s_bad[r]) Xb[o] = d0[r];
                if (has1) {
                    const float e = x1[r] - d1[r];
                    s = fmaf(e, e, s);
                    if (s_bad[r]) Xb[o + blockDim.x] = d1[r];
                }
            }
        }
        e2 += (double)s;
    }
    for (int i = threadIdx.x; i < (F - Fmain) * FR; i += blockDim.x) {
        const int f = Fmain + i / FR, r = i % FR;
        if (t0 + r >= T) continue;
        const float* wr = Wb + (long long)f * KP;
        float dot = 0.f;
        for (int k = 0; k < KP; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wr + k);
            const float4 h = *reinterpret_cast<const float4*>(&sH[r][k]);
            dot = fmaf(w.x, h.x, dot);
            dot = fmaf(w.y, h.y, dot);
            dot = fmaf(w.z, h.z, dot);
            dot = fmaf(w.w, h.w, dot);
        }
        const long long o = (long long)(t0 + r) * ldf + f;
        const float d = Xb[o] - dot;
        e2 += (double)(d * d);
        if (s_bad[r]) Xb[o] = dot;
    }
    const double tot = block_sum_d(e2, s_red);
    if (threadIdx.x == 0) err_partial[(long long)b * gridDim.x + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kThreads)
err_reduce_kernel(ClipState* __restrict__ st, int B, const double* __restrict__ err_partial, int n,
                  double* __restrict__ err_sq) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v += err_partial[(long long)b * n + i];
    v = warp_sum_d(v);
    if (lane == 0) {
        if (err_sq) err_sq[b] = v;          // time-sharded mode: the caller all-reduces, then launch_set_err
        else st[b].err = (float)sqrt(v);
    }
}
__global__ void __launch_bounds__(kThreads)
set_err_kernel(ClipState* __restrict__ st, int B, const double* __restrict__ err_sq) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) st[b].err = (float)sqrt(err_sq[b]);
}
cudaError_t launch_set_err(ClipState* st, int B, const double* err_sq, cudaStream_t s) {
    AINMF_LAUNCH(set_err_kernel, dim3(ceil_div(B, kThreads)), dim3(kThreads), 0, s, st, B, err_sq);
    return cudaGetLastError();
}

// =====================================================================================================
// host side
// =====================================================================================================
template <int KP>
static cudaError_t run_gram(const float* A, long long a_stride, int rows, int B, const NmfWork& wk, float* G,
                            const ClipState* st, cudaStream_t s) {
    int P = ceil_div(rows, 256);
    if (P > wk.gram_max_blocks) P = wk.gram_max_blocks;
    if (P < 1) P = 1;
    const int rpb = round_up(ceil_div(rows, P), 32);
    P = ceil_div(rows, rpb);
    AINMF_LAUNCH(gram_kernel<KP>, dim3(P, B), dim3(kThreads), 0, s, A, a_stride, rows, rpb, wk.gram_partial, G,
                 wk.counters, st);
    return cudaGetLastError();
}

// hbad[b][k] = sum over the bad frames t >= t_good[b] of Ht[b][t][k] (good-first frame order).  grid = (PB, B): block p sums
// its contiguous share of the bad rows into part[b][p][:]; hbad_reduce_kernel adds the PB shares in order.
__global__ void __launch_bounds__(kThreads)
hbad_kernel(const float* __restrict__ Ht, long long h_stride, int T, int KP, const int* __restrict__ t_good, float* __restrict__ part,
            const ClipState* __restrict__ st) {
    __shared__ float s_acc[kThreads];
    const int b = blockIdx.y, PB = gridDim.x;
    if (st[b].done) return;
    const int G = kThreads / KP, k = threadIdx.x % KP, gidx = threadIdx.x / KP;
    const int t0 = t_good[b], per = (T - t0 + PB - 1) / PB;
    const int lo = t0 + blockIdx.x * per, hi = min(T, lo + per);
    float acc = 0.f;
    for (int t = lo + gidx; t < hi; t += G) acc += Ht[(long long)b * h_stride + (long long)t * KP + k];
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    if (gidx == 0) {
        float sum = 0.f;
        for (int i = 0; i < G; ++i) sum += s_acc[i * KP + k];
        part[((long long)b * PB + blockIdx.x) * KP + k] = sum;
    }
}
__global__ void __launch_bounds__(128)
hbad_reduce_kernel(const float* __restrict__ part, int PB, int KP, float* __restrict__ hbad, const ClipState* __restrict__ st) {
    const int b = blockIdx.x, k = threadIdx.x;
    if (st[b].done || k >= KP) return;
    float sum = 0.f;
    for (int p = 0; p < PB; ++p) sum += part[((long long)b * PB + p) * KP + k];
    hbad[(long long)b * KP + k] = sum;
}

template <int KP, int BM>
static cudaError_t run_h_step(const NmfProblem& p, const NmfWork& wk, cudaStream_t s, float* xtw_out = nullptr) {
    using Cfg = HStepCfg<KP, BM>;
    cudaError_t e = cudaFuncSetAttribute(h_step_kernel<KP, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    auto kern = h_step_kernel<KP, BM>;
    AINMF_LAUNCH(kern, dim3(ceil_div(p.T, BM), p.B), dim3(kThreads), Cfg::smem_bytes, s, p.Xt,
                 p.x_stride, p.ldf, p.F, p.T, p.W, p.w_stride, wk.WtW, p.Ht, p.h_stride, wk.violH, p.state, xtw_out,
                 (wk.exact_viol && !xtw_out) ? wk.pgH : nullptr);
    return cudaGetLastError();
}

// red[f][k] = sum_s partial[s][f][k] (fixed order); used by the time-sharded mode to build the all-reduce buffer
__global__ void __launch_bounds__(kThreads)
reduce_partials_kernel(const float* __restrict__ partial, int S, long long n4, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(partial + ((long long)s * n4 + i) * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i * 4) = acc;
}
// out[b] = sum_i v[b][i] in double (one warp per clip)
__global__ void __launch_bounds__(kThreads)
viol_sum_kernel(const float* __restrict__ v, int n, int B, double* __restrict__ out, float* __restrict__ pack) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += (double)v[(long long)b * n + i];
    s = warp_sum_d(s);
    if (lane == 0) {
        out[b] = s;
        if (pack) {              // the same sum as two floats (hi + lo) riding at the end of the next float all-reduce
            const float hi = (float)s;
            pack[4 * b] = hi; pack[4 * b + 1] = (float)(s - (double)hi); pack[4 * b + 2] = 0.f; pack[4 * b + 3] = 0.f;
        }
    }
}

template <int KP>
static cudaError_t iterate_impl(const NmfProblem& p, const NmfWork& wk, int it, int phases, cudaStream_t s) {
    cudaError_t e;
    const int fps = round_up(ceil_div(p.T, wk.xht_splits), 16);
    const int S = wk.use_tc ? wk.tc_splits : ceil_div(p.T, fps);
    if ((phases & NMF_PHASE_PARTIALS) && wk.use_tc) {
        // tensor-core path: one kernel yields the X.Ht partials and the Gram of Ht
        // hbad reads Ht only, like the X.Ht kernel: with a side stream its two small launches (latency-bound, a few blocks
        // per clip) share the SMs with the persistent X.Ht kernel (one CTA of 512 threads per SM leaves room for them)
        // instead of adding their 13 us to every iteration
        const bool fork = p.t_good && wk.aux_stream;
        cudaStream_t hs = fork ? wk.aux_stream : s;
        auto launch_hbad = [&]() -> cudaError_t {
            AINMF_LAUNCH(hbad_kernel, dim3(wk.hbad_blocks, p.B), dim3(kThreads), 0, hs, p.Ht, p.h_stride, p.T, KP, p.t_good, wk.tc_hbad_part,
                         p.state);
            AINMF_LAUNCH(hbad_reduce_kernel, dim3(p.B), dim3(128), 0, hs, wk.tc_hbad_part, wk.hbad_blocks, KP, wk.tc_hbad, p.state);
            return cudaGetLastError();
        };
        if (fork) {
            if ((e = cudaEventRecord(wk.ev_fork, s)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(hs, wk.ev_fork, 0)) != cudaSuccess) return e;
        }
        prof_begin(PROF_XHT, s);
        if ((e = nmf_tc_half1(p, wk, s)) != cudaSuccess) return e;
        if (p.t_good && (e = launch_hbad()) != cudaSuccess) return e;
        if (fork) {
            if ((e = cudaEventRecord(wk.ev_join, hs)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(s, wk.ev_join, 0)) != cudaSuccess) return e;
        }
        prof_end(PROF_XHT, s);
        if (wk.xht_reduced) {
            const long long n4 = (long long)p.F * KP / 4;
            AINMF_LAUNCH(reduce_partials_kernel, dim3((unsigned)ceil_div64(n4, kThreads)), dim3(kThreads), 0, s,
                         wk.xht_partial, S, n4, wk.xht_reduced);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    } else if (phases & NMF_PHASE_PARTIALS) {
        // W half-step, local part: Gram of Ht and the X.Ht partial sums
        prof_begin(PROF_GRAM_H, s);
        if ((e = run_gram<KP>(p.Ht, p.h_stride, p.T, p.B, wk, wk.HHt, p.state, s)) != cudaSuccess) return e;
        prof_end(PROF_GRAM_H, s);
        prof_begin(PROF_XHT, s);
        AINMF_LAUNCH(xht_kernel<KP>, dim3(ceil_div(p.F, 128), S, p.B), dim3(kThreads), 0, s, p.Xt, p.x_stride, p.ldf,
                     p.F, p.T, p.Ht, p.h_stride, fps, wk.xht_partial, p.state);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        prof_end(PROF_XHT, s);
        if (wk.xht_reduced) {     // time-sharded mode (B == 1): one contiguous [F][KP] buffer for the all-reduce
            const long long n4 = (long long)p.F * KP / 4;
            AINMF_LAUNCH(reduce_partials_kernel, dim3((unsigned)ceil_div64(n4, kThreads)), dim3(kThreads), 0, s,
                         wk.xht_partial, S, n4, wk.xht_reduced);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    if (phases & NMF_PHASE_UPDATE) {
        prof_begin(PROF_W_SWEEP, s);
        if ((e = launch_w_side(p, wk, S, s)) != cudaSuccess) return e;
        prof_end(PROF_W_SWEEP, s);
        // H half-step (W^T W came out of the W-side kernel)
        prof_begin(PROF_H_STEP, s);
        if (wk.use_tc) e = nmf_tc_hstep(p, wk, s);
        else if (wk.h_bm == 32) e = run_h_step<KP, 32>(p, wk, s);
        else if (wk.h_bm == 64) e = run_h_step<KP, 64>(p, wk, s);
        else e = run_h_step<KP, 128>(p, wk, s);
        if (e != cudaSuccess) return e;
        prof_end(PROF_H_STEP, s);
        if (wk.h_viol_sum) {      // time-sharded mode: local H-side violation as one double for the all-reduce
            AINMF_LAUNCH(viol_sum_kernel, dim3(ceil_div(p.B, kThreads / 32)), dim3(kThreads), 0, s, wk.violH, wk.nH,
                         p.B, wk.h_viol_sum, wk.h_viol_pack);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    if (phases & NMF_PHASE_STOP) {
        prof_begin(PROF_STOP, s);
        AINMF_LAUNCH(stop_kernel, dim3(ceil_div(p.B, kThreads / 32)), dim3(kThreads), 0, s, p.state, p.B, wk.violW,
                     wk.nW, wk.violH, wk.h_viol_sum ? 0 : wk.nH, (const double*)wk.h_viol_sum,
                     (phases & NMF_PHASE_STOP_PACKED) ? (const float*)wk.h_viol_pack : nullptr, it, p.tol,
                     (wk.exact_viol && !wk.h_viol_sum) ? wk.pgW : nullptr, wk.pgH, p.F, p.T, KP);
        e = cudaGetLastError();
        prof_end(PROF_STOP, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t nmf_cd_phase(const NmfProblem& p, const NmfWork& wk, int it, int phases, cudaStream_t s) {
    switch (p.KP) {
        case 32: return iterate_impl<32>(p, wk, it, phases, s);
        case 64: return iterate_impl<64>(p, wk, it, phases, s);
        case 128: return iterate_impl<128>(p, wk, it, phases, s);
    }
    return (cudaError_t)1;
}

cudaError_t nmf_cd_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s) {
    return nmf_cd_phase(p, wk, it, NMF_PHASE_PARTIALS | NMF_PHASE_UPDATE | NMF_PHASE_STOP, s);
}

template <int KP>
static cudaError_t finalize_impl(const NmfProblem& p, const NmfWork& wk, const unsigned char* bad,
                                 long long bad_stride, double* err_sq, cudaStream_t s) {
    const int n = ceil_div(p.T, kFinalizeFrames);
    AINMF_LAUNCH(finalize_kernel<KP>, dim3(n, p.B), dim3(kThreads), 0, s, p.Xt, p.x_stride, p.ldf, p.F, p.T, p.W,
                 p.w_stride, p.Ht, p.h_stride, bad, bad_stride, wk.err_partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(err_reduce_kernel, dim3(ceil_div(p.B, kThreads / 32)), dim3(kThreads), 0, s, p.state, p.B,
                 wk.err_partial, n, err_sq);
    return cudaGetLastError();
}

// =====================================================================================================
// multiplicative update (Frobenius):  A[row,:] *= Num[row,:] / (A[row,:] . G), zero denominators -> float32 eps
// ($SP/sklearn/decomposition/_nmf.py:536-549 for W, :615-624 for H).  Num = sum over `S` split buffers.
// 32 rows per block, 8 lanes per row; grid = (ceil(rows/32), B).
// =====================================================================================================
template <int KP>
__global__ void __launch_bounds__(kThreads)
mu_rows_kernel(float* __restrict__ A, long long a_stride, int rows, const float* __restrict__ G,
               const float* __restrict__ num, int S, long long num_split_stride, long long num_clip_stride,
               const ClipState* __restrict__ st) {
    constexpr int L = 8, SL = KP / L, PITCH = KP + 4 * L;
    AINMF_DYN_SMEM(smem_raw);
    float* sG = reinterpret_cast<float*>(smem_raw);        // [KP][PITCH] padded as in load_gram_padded<KP, 8>
    float* sA = sG + KP * PITCH;                           // [32][KP]
    const int b = blockIdx.y;
    if (st[b].done) return;
    load_gram_padded<KP, L>(sG, G + (long long)b * KP * KP);
    const int l = threadIdx.x % L, lr = threadIdx.x / L;
    const int row = blockIdx.x * 32 + lr;
    const bool valid = row < rows;
    float* ar = A + (long long)b * a_stride + (long long)row * KP;
    for (int i = threadIdx.x; i < 32 * KP / 4; i += blockDim.x) {
        const int r = (4 * i) / KP, c = (4 * i) % KP;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (blockIdx.x * 32 + r < rows) v = *reinterpret_cast<const float4*>(A + (long long)b * a_stride + (long long)(blockIdx.x * 32 + r) * KP + c);
        *reinterpret_cast<float4*>(sA + r * KP + c) = v;
    }
    __syncthreads();
    float den[SL];
#pragma unroll
    for (int q = 0; q < SL; ++q) den[q] = 0.f;
    const float* arow = sA + lr * KP;
    const float* gl = sG + l * (SL + 4);
    for (int r = 0; r < KP; ++r) {
        const float a = arow[r];
        const float* g = gl + r * PITCH;                   // G[r][l*SL .. ] (G symmetric: column slice = row slice)
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(g + q);
            den[q] = fmaf(a, gv.x, den[q]);
            den[q + 1] = fmaf(a, gv.y, den[q + 1]);
            den[q + 2] = fmaf(a, gv.z, den[q + 2]);
            den[q + 3] = fmaf(a, gv.w, den[q + 3]);
        }
    }
    if (valid) {
        float nm[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) nm[q] = 0.f;
        for (int s = 0; s < S; ++s) {
            const float* nr = num + (long long)b * num_clip_stride + (long long)s * num_split_stride + (long long)row * KP + l * SL;
#pragma unroll
            for (int q = 0; q < SL; q += 4) {
                const float4 v = *reinterpret_cast<const float4*>(nr + q);
                nm[q] += v.x; nm[q + 1] += v.y; nm[q + 2] += v.z; nm[q + 3] += v.w;
            }
        }
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = (den[q + e] == 0.f) ? 1.1920929e-07f : den[q + e];
                o[e] = arow[l * SL + q + e] * (nm[q + e] / d);
            }
            *reinterpret_cast<float4*>(ar + l * SL + q) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// it == 0: error of the initial factors; it % 10 == 0: (previous_error - error) / error_at_init < tol ($SP .../_nmf.py:867-879)
__global__ void __launch_bounds__(kThreads)
mu_stop_kernel(ClipState* __restrict__ st, int B, int it, float tol) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ClipState s = st[b];
    if (s.done) return;
    const double e = (double)s.err;
    if (it == 0) { s.err_init = e; s.err_prev = e; }
    else {
        if ((s.err_prev - e) / s.err_init < (double)tol) s.done = 1;
        s.err_prev = e;
    }
    st[b] = s;
}
__global__ void __launch_bounds__(kThreads)
mu_tick_kernel(ClipState* __restrict__ st, int B, int it) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && !st[b].done) st[b].n_iter = it;
}

template <int KP>
static cudaError_t mu_rows(float* A, long long a_stride, int rows, const float* G, const float* num, int S,
                           long long num_split_stride, long long num_clip_stride, const NmfProblem& p, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)KP * (KP + 32) + 32 * KP);
    cudaError_t e = cudaFuncSetAttribute(mu_rows_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(mu_rows_kernel<KP>, dim3(ceil_div(rows, 32), p.B), dim3(kThreads), smem, s, A, a_stride, rows, G, num, S,
                 num_split_stride, num_clip_stride, p.state);
    return cudaGetLastError();
}

template <int KP>
static cudaError_t mu_error(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s) {
    cudaError_t e = finalize_impl<KP>(p, wk, wk.zero_flags, wk.zero_stride, nullptr, s);
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(mu_stop_kernel, dim3(ceil_div(p.B, kThreads)), dim3(kThreads), 0, s, p.state, p.B, it, p.tol);
    return cudaGetLastError();
}

template <int KP>
static cudaError_t mu_iterate_impl(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s) {
    cudaError_t e;
    const int fps = round_up(ceil_div(p.T, wk.xht_splits), 16);
    const int S = ceil_div(p.T, fps);
    // W <- W * (X Ht) / (W (Ht^T Ht))
    if ((e = run_gram<KP>(p.Ht, p.h_stride, p.T, p.B, wk, wk.HHt, p.state, s)) != cudaSuccess) return e;
    AINMF_LAUNCH(xht_kernel<KP>, dim3(ceil_div(p.F, 128), S, p.B), dim3(kThreads), 0, s, p.Xt, p.x_stride, p.ldf, p.F, p.T,
                 p.Ht, p.h_stride, fps, wk.xht_partial, p.state);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = mu_rows<KP>(p.W, p.w_stride, p.F, wk.HHt, wk.xht_partial, S, (long long)p.F * KP, (long long)S * p.F * KP, p, s)) != cudaSuccess) return e;
    // H <- H * (W^T X) / ((W^T W) H)
    if ((e = run_gram<KP>(p.W, p.w_stride, p.F, p.B, wk, wk.WtW, p.state, s)) != cudaSuccess) return e;
    if (wk.h_bm == 32) e = run_h_step<KP, 32>(p, wk, s, wk.xtw);
    else if (wk.h_bm == 64) e = run_h_step<KP, 64>(p, wk, s, wk.xtw);
    else e = run_h_step<KP, 128>(p, wk, s, wk.xtw);
    if (e != cudaSuccess) return e;
    if ((e = mu_rows<KP>(p.Ht, p.h_stride, p.T, wk.WtW, wk.xtw, 1, 0, p.h_stride, p, s)) != cudaSuccess) return e;
    AINMF_LAUNCH(mu_tick_kernel, dim3(ceil_div(p.B, kThreads)), dim3(kThreads), 0, s, p.state, p.B, it);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (p.tol > 0.f && it % 10 == 0) return mu_error<KP>(p, wk, it, s);
    return cudaSuccess;
}

cudaError_t nmf_mu_tick(const NmfProblem& p, int it, cudaStream_t s) {
    AINMF_LAUNCH(mu_tick_kernel, dim3(ceil_div(p.B, kThreads)), dim3(kThreads), 0, s, p.state, p.B, it);
    return cudaGetLastError();
}
cudaError_t nmf_mu_stop(const NmfProblem& p, int it, cudaStream_t s) {
    AINMF_LAUNCH(mu_stop_kernel, dim3(ceil_div(p.B, kThreads)), dim3(kThreads), 0, s, p.state, p.B, it, p.tol);
    return cudaGetLastError();
}
cudaError_t nmf_mu_begin(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    switch (p.KP) {
        case 32: return mu_error<32>(p, wk, 0, s);
        case 64: return mu_error<64>(p, wk, 0, s);
        case 128: return mu_error<128>(p, wk, 0, s);
    }
    return (cudaError_t)1;
}
cudaError_t nmf_mu_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s) {
    switch (p.KP) {
        case 32: return mu_iterate_impl<32>(p, wk, it, s);
        case 64: return mu_iterate_impl<64>(p, wk, it, s);
        case 128: return mu_iterate_impl<128>(p, wk, it, s);
    }
    return (cudaError_t)1;
}

cudaError_t nmf_finalize(const NmfProblem& p, const NmfWork& wk, const unsigned char* bad, long long bad_stride,
                         cudaStream_t s, double* err_sq) {
    switch (p.KP) {
        case 32: return finalize_impl<32>(p, wk, bad, bad_stride, err_sq, s);
        case 64: return finalize_impl<64>(p, wk, bad, bad_stride, err_sq, s);
        case 128: return finalize_impl<128>(p, wk, bad, bad_stride, err_sq, s);
    }
    return (cudaError_t)1;
}

// Workspace layout for the iteration (all sizes in bytes, 256-aligned), see NmfWork in kernels.h.
static size_t al256(size_t n) { return (n + 255) / 256 * 256; }

void nmf_plan(int B, int T, int F, int KP, int n_sm, NmfWork* wk) {
    // h-step tile: the largest BM that still gives every SM two blocks
    const long long want = 2LL * n_sm;
    int bm = 128;
    if ((long long)B * ceil_div(T, 128) < want) bm = 64;
    if ((long long)B * ceil_div(T, 64) < want) bm = 32;
    if (KP == 128 && bm == 128) bm = 64;                    // keeps two blocks per SM resident (smem)
    wk->h_bm = bm;
    wk->nH = ceil_div(T, bm);
    {   // W-side kernel shape: 128 rows per block and as few lanes per row as still give the machine about 8 warps per
        // SM; when even 8 lanes leave most SMs idle (one long signal), 32-row blocks with the most lanes
        wk->w_rows = 128;
        wk->nW = ceil_div(F, 128);
        const long long blocks = (long long)B * wk->nW;
        int lanes = 1;
        while (lanes < 8 && blocks * 4 * lanes < 8LL * n_sm) lanes *= 2;
        wk->w_lanes = lanes;
        if (lanes == 8 && blocks * 2 < n_sm) { wk->w_rows = 32; wk->nW = ceil_div(F, 32); }
        if (lanes == 1 && KP <= 64) {                       // two rows per thread (nmf_wside.cu); a remainder of at most
            wk->w_rows = 256;                               // kWSideTail rows (F = 2^m + 1) rides with the clip's last block
            const int rem = F % 256;
            wk->nW = (F > 256 && rem > 0 && rem <= kWSideTail) ? F / 256 : ceil_div(F, 256);
        }
    }
    {   // w_finish_kernel: groups of 8 Gram rows per block -- as many as still leave two blocks per SM
        int gpb = 1;
        while (gpb * 2 <= KP / 8 && (long long)B * (KP / 8 / (gpb * 2)) >= want) gpb *= 2;
        wk->finish_gpb = gpb;
    }
    const int f_tiles = ceil_div(F, 128);
    long long splits = want / ((long long)B * f_tiles);
    if (splits < 1) splits = 1;
    if (splits > ceil_div(T, 16)) splits = ceil_div(T, 16);
    wk->xht_splits = (int)splits;
    long long gb = want / B;
    if (gb < 1) gb = 1;
    if (gb > want) gb = want;
    wk->gram_max_blocks = (int)gb;
    wk->exact_viol = 0;
#ifndef AINMF_EMU
    // tensor-core path for the V-sized contractions (disable with AINMF_DISABLE_TC=1, e.g. to test the FFMA kernels)
    const char* off = getenv("AINMF_DISABLE_TC");
    wk->use_tc = (KP >= 64 && T >= 128 && F >= 128 && !(off && off[0] == '1')) ? 1 : 0;
    if (wk->use_tc) {
        const int nxs = ceil_div(F, 32);
        wk->tc_mtiles = ceil_div(32 * nxs + KP, 128);
        long long sp = (long long)n_sm / ((long long)B * wk->tc_mtiles);
        if (sp < 1) sp = 1;
        if (sp > ceil_div(T, 32)) sp = ceil_div(T, 32);
        wk->tc_fps = round_up(ceil_div(T, (int)sp), 32);
        wk->tc_splits = ceil_div(T, wk->tc_fps);
        wk->nH = ceil_div(T, 128);
        long long hb = 2LL * n_sm / B;
        wk->hbad_blocks = (int)(hb < 1 ? 1 : (hb > 64 ? 64 : hb));
    }
#endif
    // small FFMA-path problems keep |pg| per (row, coordinate) for the reference-order violation sum (stop_kernel)
    wk->exact_viol = (!wk->use_tc && ((long long)B * ((long long)F + T) * KP <= (1LL << 20))) ? 1 : 0;
}

size_t nmf_work_bytes(int B, int T, int F, int KP, const NmfWork& wk) {
    size_t n = 0;
    n += al256(sizeof(float) * (size_t)B * KP * KP) * 2;                          // HHt, WtW
    int gp = wk.gram_max_blocks > wk.tc_splits ? wk.gram_max_blocks : wk.tc_splits;
    if (wk.nW > gp) gp = wk.nW;
    const int xs = wk.xht_splits > wk.tc_splits ? wk.xht_splits : wk.tc_splits;
    n += al256(sizeof(float) * (size_t)B * gp * KP * KP);                          // gram partials
    n += al256(sizeof(unsigned) * (size_t)B);                                      // counters
    n += al256(sizeof(float) * (size_t)B * (xs + 1) * F * KP);                      // xht partials
    if (wk.use_tc) n += 2 * al256(sizeof(float) * (size_t)B * KP * round_up(F, 4)) + al256(sizeof(float) * (size_t)B * KP * KP) + al256(sizeof(float) * (size_t)B * (KP / 8) * (16 * KP)) + al256(sizeof(float) * (size_t)B * (KP / 8) * kSweepScalars) // Wt, Wt_lo, G_lo, sweep blobs + scalars
                       + al256(sizeof(float) * (size_t)B * wk.nW * KP) + 2 * al256(sizeof(float) * (size_t)B * KP) + al256(sizeof(float) * (size_t)B * wk.hbad_blocks * KP); // fill^T.W partials, bad-frame sums of Ht, fill^T.W, hbad shares
    n += al256(sizeof(float) * (size_t)B * wk.nW) + al256(sizeof(float) * (size_t)B * wk.nH);
    n += al256(sizeof(double) * (size_t)B * ceil_div(T, 16));
    n += al256(4096);                                                             // coop_scratch
    if (wk.want_mu == 1) n += al256(sizeof(float) * (size_t)B * T * KP) + al256((size_t)B * round_up(T, 16));
    if (wk.want_mu == 2) n += al256(sizeof(float) * (size_t)B * 2 * KP) + al256(sizeof(double) * (size_t)B);
    if (wk.exact_viol) n += al256(sizeof(float) * (size_t)B * F * KP) + al256(sizeof(float) * (size_t)B * T * KP);
    return n;
}

void nmf_carve(void* base, int B, int T, int F, int KP, NmfWork* wk) {
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) { char* r = p; p += al256(bytes); return r; };
    wk->HHt = (float*)take(sizeof(float) * (size_t)B * KP * KP);
    wk->WtW = (float*)take(sizeof(float) * (size_t)B * KP * KP);
    int gp = wk->gram_max_blocks > wk->tc_splits ? wk->gram_max_blocks : wk->tc_splits;
    if (wk->nW > gp) gp = wk->nW;
    const int xs = wk->xht_splits > wk->tc_splits ? wk->xht_splits : wk->tc_splits;
    wk->gram_partial = (float*)take(sizeof(float) * (size_t)B * gp * KP * KP);
    wk->counters = (unsigned*)take(sizeof(unsigned) * (size_t)B);
    wk->xht_partial = (float*)take(sizeof(float) * (size_t)B * (xs + 1) * F * KP);
    if (wk->use_tc) {
        wk->tc_Wt = (float*)take(sizeof(float) * (size_t)B * KP * round_up(F, 4));
        wk->tc_WtLo = (float*)take(sizeof(float) * (size_t)B * KP * round_up(F, 4));
        wk->tc_GLo = (float*)take(sizeof(float) * (size_t)B * KP * KP);
        wk->tc_blobs = (float*)take(sizeof(float) * (size_t)B * (KP / 8) * (16 * KP));
        wk->tc_scal = (float*)take(sizeof(float) * (size_t)B * (KP / 8) * kSweepScalars);
        wk->tc_vpartial = (float*)take(sizeof(float) * (size_t)B * wk->nW * KP);
        wk->tc_hbad = (float*)take(sizeof(float) * (size_t)B * KP);
        wk->tc_vfill = (float*)take(sizeof(float) * (size_t)B * KP);
        wk->tc_hbad_part = (float*)take(sizeof(float) * (size_t)B * wk->hbad_blocks * KP);
    }
    wk->violW = (float*)take(sizeof(float) * (size_t)B * wk->nW);
    wk->violH = (float*)take(sizeof(float) * (size_t)B * wk->nH);
    wk->err_partial = (double*)take(sizeof(double) * (size_t)B * ceil_div(T, 16));
    wk->coop_scratch = (char*)take(4096);
    if (wk->want_mu == 2) {
        wk->kl_sums = (float*)take(sizeof(float) * (size_t)B * 2 * KP);
        wk->kl_err = (double*)take(sizeof(double) * (size_t)B);
    }
    if (wk->want_mu == 1) {
        wk->xtw = (float*)take(sizeof(float) * (size_t)B * T * KP);
        wk->zero_stride = round_up(T, 16);
        wk->zero_flags = (unsigned char*)take((size_t)B * wk->zero_stride);      // the caller zeroes it once
    }
    if (wk->exact_viol) {
        wk->pgW = (float*)take(sizeof(float) * (size_t)B * F * KP);
        wk->pgH = (float*)take(sizeof(float) * (size_t)B * T * KP);
    }
}

#ifdef AINMF_EMU
int nmf_tc_setup(const NmfProblem&, NmfWork* wk, TcMaps*) { wk->use_tc = 0; return 0; }
cudaError_t nmf_tc_half1(const NmfProblem&, const NmfWork&, cudaStream_t) { return (cudaError_t)1; }
cudaError_t nmf_tc_hstep(const NmfProblem&, const NmfWork&, cudaStream_t) { return (cudaError_t)1; }
#endif

}  // namespace ainmf
