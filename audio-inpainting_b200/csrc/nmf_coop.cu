// nmf_coop.cu -- the whole coordinate-descent fit of ONE spectrogram (BASELINE configs[1], [2]: a 10 s clip) as one
// cooperative persistent launch: every SM keeps a few rows of W and a few frames of Ht -- and the matching slices of the
// spectrogram -- in shared memory for all iterations, the iteration's four phases are separated by grid barriers, the stop
// rule is evaluated on the device.
//
// Why: through the general kernels one clip is six launches per iteration that each run 5-26 us on a handful of SMs (the
// H step has 7 tiles for 148 SMs): 70 us per iteration, 14 ms per clip.  Here the 2 x F x T x K contractions of an iteration
// are spread over all SMs as FFMA, and the sweeps run on the rows a CTA owns.
//
// Arithmetic per iteration is the reference's (sklearn _nmf.py:491-516, _cdnmf_fast.pyx:8-38):
//   A  every CTA streams Ht (all frames) through shared memory once: X.Ht for its rows of W, and its 1/128 of the elements
//      of Ht^T.Ht (complete sums: no reduction over CTAs)
//   B  W sweep on the CTA's rows (8 lanes per row); rows -> global
//   C  every CTA streams W: Xt.W for its frames, its share of W^T.W
//   D  H sweep on the CTA's frames; frames -> global; the CTA's share of the violation
//   then every CTA adds the shares in CTA order (identical everywhere) and tests `violation / violation_init <= tol`.
#include "kernels.h"
#include <stdio.h>
#include <stdlib.h>

namespace ainmf {

#ifdef AINMF_EMU
// the CPU test harness runs blocks one after the other: no grid barrier, the general kernels serve these problems there
bool nmf_coop_eligible(const NmfProblem&, const NmfWork&, int) { return false; }
cudaError_t nmf_coop_fit(const NmfProblem&, const NmfWork&, int, int, cudaStream_t) { return (cudaError_t)1; }
#else

namespace {

constexpr int kCoopThreads = 256;
constexpr int kCoopGramCtas = 128;           // CTAs that each own KP*KP/128 elements of a Gram
constexpr int kCoopMaxRows = 16;             // rows of W / frames of Ht a CTA can own (accumulators of the products)

template <int KP> struct CoopCfg {
    static constexpr int L = 8, SL = KP / L, GP = KP + 4 * L, AP = KP + 4;
    static constexpr int CH = 4096 / KP;     // rows of a factor per staged chunk (16 KB)
    static constexpr int STAGES = (KP == 128) ? 2 : 3, STAGE_FLOATS = CH * KP;   // K = 128: G alone is 82 KB
    static constexpr int KQ = KP / 4;        // products: lanes along k (one float4 each) ...
    static constexpr int NJ = kCoopThreads / KQ;   // ... times groups that split the streamed rows
    static constexpr int WG = (KQ < 32) ? 32 / KQ : 1;            // groups that share a warp (added by shuffles)
    static constexpr int NJR = NJ / WG;                           // groups left after the in-warp reduction
    static constexpr int GE = KP * KP / kCoopGramCtas;            // Gram elements per CTA; 256 / GE row groups
    // G | 1/diag | own rows of W | own frames of Ht | Gram partials | ring of staged chunks (after a stream: the groups' sums)
    static constexpr size_t fixed_floats = (size_t)KP * GP + KP + 2 * (size_t)kCoopMaxRows * AP + kCoopThreads + (size_t)STAGES * STAGE_FLOATS;
};
__host__ __device__ inline int coop_pad4(int n) { return (n + 3) / 4 * 4; }
__host__ __device__ inline size_t coop_x_floats(int F, int T, int rows_w, int rows_h) {
    return (size_t)T * coop_pad4(rows_w) + (size_t)F * coop_pad4(rows_h);
}

struct CoopParams {
    const float* X; int ldf, F, T;
    float* W; float* Ht;                     // [F][KP], [T][KP]
    ClipState* st;
    float* gram;                             // [2][KP*KP]: Ht^T.Ht, W^T.W
    double* viol_part;                       // [gridDim.x]
    unsigned* bar;                           // zeroed before the launch
    int max_iter; float tol;
    int rows_w, rows_h;                      // rows of W / frames of Ht per CTA
    long long* dbg;                          // AINMF_COOP_DEBUG=1: cycles per phase of CTA 0
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");     // .cg: L2 only, coherent
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One pass over Other [n][KP] (global, written by other CTAs), staged through the ring in chunks of CH rows with STAGES
// chunks in flight.  Per chunk, from shared memory:
//   products  acc[r][0..3] += x(r, j) * Other[j][4kq..4kq+3] for the CTA's rows r (x(r, j) = sX[j*rpp + r], zero-padded to
//             rpp = a multiple of 4 rows); thread = (kq, jg): the NJ groups split the rows j of a chunk;
//   Gram      this CTA's GE elements of Other^T.Other, thread = (element, row group).
// Afterwards the groups' sums are added in group order: on return the ring holds, at [r][KP], the products B[r][:] of the
// CTA's rows, and (CTAs < 128) gram_out[cta*GE ..] the finished Gram elements.  All threads call.
template <int KP, int RQ>                      // RQ = rpp / 4: quads of rows (compile time: no branches among the FMAs)
__device__ __forceinline__ void stream_products(const float* __restrict__ Other, int n, const float* __restrict__ sX,
                                                float* __restrict__ ring, float* __restrict__ sGp, float* __restrict__ gram_out, long long* tw = nullptr) {
    constexpr int rpp = 4 * RQ;
    long long qa = 0;
    if (tw) qa = clock64();
    using Cfg = CoopCfg<KP>;
    constexpr int CH = Cfg::CH, SF = Cfg::STAGE_FLOATS, KQ = Cfg::KQ, NJ = Cfg::NJ, WG = Cfg::WG, NJR = Cfg::NJR, GE = Cfg::GE,
                  RG = kCoopThreads / GE;
    const int tid = threadIdx.x;
    const int kq = tid % KQ, jg = tid / KQ;
    const bool gram_cta = (int)blockIdx.x < kCoopGramCtas;
    const int ge = tid % GE, rg = tid / GE;
    const int gidx = (int)blockIdx.x * GE + ge, gi = gram_cta ? gidx / KP : 0, gj = gram_cta ? gidx % KP : 0;
    float2 acc[rpp][2];                                                // [row][k pair]: packed FFMA2
#pragma unroll
    for (int r = 0; r < rpp; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
    float gacc = 0.f;
    const int nch = (n + CH - 1) / CH;
    // Every CTA reads the same matrix: started together they would all ask the same L2 slices for the same lines at the same
    // moment (measured: 2 300 cycles per 16 KB chunk, 1.9 TB/s aggregate).  Each CTA therefore starts at its own chunk and
    // wraps around; a CTA's order is fixed, so its sums are reproducible.
    const int rot = (int)(((long long)blockIdx.x * nch) / gridDim.x);
    auto phys = [&](int c) { const int q = c + rot; return q >= nch ? q - nch : q; };
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring) + 16u * tid;      // this thread's 16 bytes of every 4 KB
    auto issue = [&](int cl) {
        const int c = phys(cl);
        const int n16 = min(CH, n - c * CH) * (KP / 4);                                 // 16-byte pieces of the chunk
        const unsigned sa = ring_sa + (unsigned)(cl % Cfg::STAGES) * (SF * 4);
        const float* src = Other + (long long)c * CH * KP + 4 * tid;
#pragma unroll
        for (int k = 0; k < CH * (KP / 4) / kCoopThreads; ++k)
            if (tid + k * kCoopThreads < n16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + k * (16u * kCoopThreads)), "l"(src + k * (4 * kCoopThreads)) : "memory");
    };
    if (tw) { const long long qb = clock64(); tw[6] += qb - qa; qa = qb; }
#pragma unroll
    for (int c = 0; c < Cfg::STAGES - 1; ++c) {
        if (c < nch) issue(c);
        cp_async_commit();
    }
    if (tw) { const long long qb = clock64(); tw[7] += qb - qa; }
    for (int c = 0; c < nch; ++c) {
        long long q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        if (tw) q0 = clock64();
        if (c + Cfg::STAGES - 1 < nch) issue(c + Cfg::STAGES - 1);
        cp_async_commit();
        if (tw) q1 = clock64();
        cp_async_wait<Cfg::STAGES - 1>();
        if (tw) q2 = clock64();
        __syncthreads();
        if (tw) q3 = clock64();
        const float* st = ring + (c % Cfg::STAGES) * SF;
        const int cp_ = phys(c);
        const int rows = min(CH, n - cp_ * CH);
        const float* xb = sX + (long long)cp_ * CH * rpp;
#pragma unroll 4
        for (int jj = jg; jj < rows; jj += NJ) {
            const float4 h = *reinterpret_cast<const float4*>(st + jj * KP + 4 * kq);
            const float* xr = xb + jj * rpp;
            float4 x[RQ];
#pragma unroll
            for (int q = 0; q < RQ; ++q) x[q] = *reinterpret_cast<const float4*>(xr + 4 * q);
            const float2 h01 = make_float2(h.x, h.y), h23 = make_float2(h.z, h.w);
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const float xv[4] = {x[q].x, x[q].y, x[q].z, x[q].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 xx = make_float2(xv[i], xv[i]);
                    acc[4 * q + i][0] = __ffma2_rn(xx, h01, acc[4 * q + i][0]);
                    acc[4 * q + i][1] = __ffma2_rn(xx, h23, acc[4 * q + i][1]);
                }
            }
        }
        if (gram_cta) {
#pragma unroll 8
            for (int rr = rg; rr < rows; rr += RG) gacc = fmaf(st[rr * KP + gi], st[rr * KP + gj], gacc);
        }
        long long q4 = 0;
        if (tw) q4 = clock64();
        __syncthreads();
        if (tw) { const long long q5 = clock64(); tw[0] += q1 - q0; tw[1] += q2 - q1; tw[2] += q3 - q2; tw[3] += q4 - q3; tw[4] += q5 - q4; }
    }
    cp_async_wait<0>();
    long long qe = 0;
    if (tw) qe = clock64();
    // groups that share a warp first (fixed order), then through the ring
    if constexpr (WG > 1) {
#pragma unroll
        for (int off = KQ; off < 32; off <<= 1)
#pragma unroll
            for (int r = 0; r < rpp; ++r)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    acc[r][i].x += __shfl_xor_sync(0xffffffffu, acc[r][i].x, off);
                    acc[r][i].y += __shfl_xor_sync(0xffffffffu, acc[r][i].y, off);
                }
    }
    float* part = ring;                                                // [NJR][rpp][KP]
    if (jg % WG == 0) {
        const int jr = jg / WG;
#pragma unroll
        for (int r = 0; r < rpp; ++r)
            *reinterpret_cast<float4*>(part + ((long long)jr * rpp + r) * KP + 4 * kq) = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
    }
    sGp[rg * GE + ge] = gacc;
    __syncthreads();
    for (int i = tid; i < rpp * KQ; i += kCoopThreads) {               // i = (row, kq)
        const int r = i / KQ, q = i % KQ;
        float4 s = *reinterpret_cast<const float4*>(part + (long long)r * KP + 4 * q);
        for (int g = 1; g < NJR; ++g) {
            const float4 v = *reinterpret_cast<const float4*>(part + ((long long)g * rpp + r) * KP + 4 * q);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(part + (long long)r * KP + 4 * q) = s;       // group 0's slot: read above by this thread only
    }
    if (gram_cta && tid < GE) {
        float s = 0.f;
        for (int g = 0; g < RG; ++g) s += sGp[g * GE + tid];
        gram_out[(int)blockIdx.x * GE + tid] = s;
    }
    __syncthreads();
    if (tw) tw[5] += clock64() - qe;
}

// G (global, finished by the CTAs of the previous phase) -> shared memory in the 8-lane layout (element (t, o*SL + q) at
// t*GP + o*(SL+4) + q), and the reciprocals of its diagonal
template <int KP>
__device__ __forceinline__ void gram_load(const float* __restrict__ G, float* __restrict__ sG, float* __restrict__ sInv) {
    constexpr int SL = CoopCfg<KP>::SL, GP = CoopCfg<KP>::GP;
    for (int e4 = threadIdx.x; e4 < KP * KP / 4; e4 += kCoopThreads) {
        const int t = (4 * e4) / KP, c = (4 * e4) % KP;
        *reinterpret_cast<float4*>(sG + t * GP + (c / SL) * (SL + 4) + (c % SL)) = ldcg4(G + 4 * e4);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP; t += kCoopThreads) {
        const float d = sG[t * GP + (t / SL) * (SL + 4) + (t % SL)];
        sInv[t] = (d != 0.f) ? 1.0f / d : 0.f;
    }
    __syncthreads();
}

// Coordinate sweep of the CTA's rows (shared-memory tile sA [rows][AP]), 8 lanes per row; sB [rows][KP] holds the products
// B (X.Ht or Xt.W rows).  The row's gradient g = A.G - B lives in the lanes' registers (lane l: coordinates l*SL..).  The
// coordinates are visited in the reference's order in blocks of SL: the owner lane's slice of g is broadcast, all eight
// lanes redo the block's SL sequential updates on identical data (no shuffle, no divergence on the critical path), then
// every lane applies the block's SL deltas to its own slice.  In exact arithmetic this is the sequential sweep.
template <int KP>
__device__ __forceinline__ float sweep_rows(float* __restrict__ sA, const float* __restrict__ sB, const float* __restrict__ sG,
                                         const float* __restrict__ sInv, int nrows) {
    constexpr int SL = KP / 8, GP = CoopCfg<KP>::GP, AP = CoopCfg<KP>::AP;
    const int l = threadIdx.x & 7, slot = threadIdx.x >> 3;
    const int slot_end = (nrows + 3) & ~3;                 // whole warps take part (the shuffles name every lane)
    float viol = 0.f;
    if (slot >= slot_end) return 0.f;
    const bool valid = slot < nrows;
    float* ar = sA + (valid ? slot : 0) * AP;
    const float* gl = sG + l * (SL + 4);
    float g[SL];
#pragma unroll
    for (int q = 0; q < SL; ++q) g[q] = valid ? -sB[slot * KP + l * SL + q] : 0.f;
#pragma unroll 4
    for (int t = 0; t < KP; ++t) {                         // g = A.G - B
        const float c = valid ? ar[t] : 0.f;
        const float* gr = gl + t * GP;
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(gr + q);
            g[q] = fmaf(c, gv.x, g[q]); g[q + 1] = fmaf(c, gv.y, g[q + 1]);
            g[q + 2] = fmaf(c, gv.z, g[q + 2]); g[q + 3] = fmaf(c, gv.w, g[q + 3]);
        }
    }
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
        float gt[SL], d[SL], anv[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) gt[q] = __shfl_sync(0xffffffffu, g[q], o, 8);
        const float* go = sG + o * (SL + 4);               // the owner's columns of G (all lanes read the same words)
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            const int t = o * SL + q;
            const float inv = sInv[t];
            const float aq = valid ? ar[t] : 0.f;
            const float grad = gt[q];
            const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
            const float an = (inv != 0.f) ? fmaxf(fmaf(-grad, inv, aq), 0.f) : aq;
            d[q] = an - aq;
            anv[q] = an;
            if (l == o && valid) viol += fabsf(pg);
            const float* gr = go + t * GP;                 // corrections inside the block: coordinates still to come
#pragma unroll
            for (int q2 = q + 1; q2 < SL; ++q2) gt[q2] = fmaf(d[q], gr[q2], gt[q2]);
        }
        __syncwarp();                                      // every lane has read the block's old coordinates
        if (l == o && valid) {
#pragma unroll
            for (int q = 0; q < SL; ++q) ar[o * SL + q] = anv[q];
        }
#pragma unroll
        for (int q = 0; q < SL; ++q) {                     // the block's deltas on this lane's slice
            const float* gr = gl + (o * SL + q) * GP;
#pragma unroll
            for (int q2 = 0; q2 < SL; q2 += 4) {
                const float4 gv = *reinterpret_cast<const float4*>(gr + q2);
                g[q2] = fmaf(d[q], gv.x, g[q2]); g[q2 + 1] = fmaf(d[q], gv.y, g[q2 + 1]);
                g[q2 + 2] = fmaf(d[q], gv.z, g[q2 + 2]); g[q2 + 3] = fmaf(d[q], gv.w, g[q2 + 3]);
            }
        }
    }
    return viol;
}

template <int KP>
__global__ void __launch_bounds__(kCoopThreads, 1)
nmf_coop_kernel(CoopParams p) {
    using Cfg = CoopCfg<KP>;
    constexpr int GP = Cfg::GP, AP = Cfg::AP;
    AINMF_DYN_SMEM(smem_raw);
    float* sG = reinterpret_cast<float*>(smem_raw);        // [KP][GP]
    float* sInv = sG + KP * GP;                            // [KP]
    float* sW = sInv + KP;                                 // [16][AP]: the CTA's rows of W
    float* sH = sW + kCoopMaxRows * AP;                    // [16][AP]: the CTA's frames of Ht
    float* sGp = sH + kCoopMaxRows * AP;                   // [256]: Gram partials of the row groups
    float* ring = sGp + kCoopThreads;                      // [STAGES][CH][KP]: staged chunks of the other factor
    const int rpw = coop_pad4(p.rows_w), rph = coop_pad4(p.rows_h);
    float* sXw = ring + Cfg::STAGES * Cfg::STAGE_FLOATS;   // [T][rpw]: X of the CTA's bins, every frame
    float* sXh = sXw + (long long)p.T * rpw;               // [F][rph]: X of the CTA's frames, every bin
    __shared__ double s_red[32];
    __shared__ double s_tot;
    const int cta = blockIdx.x, nb_ctas = gridDim.x, tid = threadIdx.x;
    if (p.st[0].status != 0) return;                       // uniform over the grid: nothing to fit
    const int f0 = cta * p.rows_w, nf = max(0, min(p.rows_w, p.F - f0));
    const int t0 = cta * p.rows_h, nt = max(0, min(p.rows_h, p.T - t0));
    // own rows and the CTA's slices of the spectrogram into shared memory, once
    for (int i = tid; i < kCoopMaxRows * (KP / 4); i += kCoopThreads) {
        const int r = i / (KP / 4), c = (i % (KP / 4)) * 4;
        *reinterpret_cast<float4*>(sW + r * AP + c) = (r < nf) ? ldcg4(p.W + (long long)(f0 + r) * KP + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sH + r * AP + c) = (r < nt) ? ldcg4(p.Ht + (long long)(t0 + r) * KP + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < p.T * rpw; i += kCoopThreads) {
        const int j = i / rpw, r = i % rpw;
        sXw[i] = (r < nf) ? p.X[(long long)j * p.ldf + f0 + r] : 0.f;
    }
    for (int i = tid; i < p.F * rph; i += kCoopThreads) {
        const int r = i / p.F, j = i % p.F;                // consecutive threads: consecutive bins of one frame
        sXh[j * rph + r] = (r < nt) ? p.X[(long long)(t0 + r) * p.ldf + j] : 0.f;
    }
    __syncthreads();
    unsigned bar_target = 0;
    double viol_init = 0.0, viol_last = 0.0;
    int n_iter = 0, done = 0;
    float* gramH = p.gram;
    float* gramW = p.gram + KP * KP;
    long long c_ph[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, c_t = 0;
    const bool dbg_on = p.dbg != nullptr && tid == 0 && cta == 0;
#define CO_TIC() do { if (dbg_on) c_t = clock64(); } while (0)
#define CO_TOC(k) do { if (dbg_on) { const long long c_ = clock64(); c_ph[k] += c_ - c_t; c_t = c_; } } while (0)
    for (int it = 1; it <= p.max_iter; ++it) {
        float v = 0.f;
        // Two halves with the same code (one copy in the instruction cache): half 0 = A + B (stream Ht: X.Ht for the own rows
        // of W and this CTA's elements of Ht^T.Ht; then the W sweep), half 1 = C + D (stream W: Xt.W for the own frames and the
        // elements of W^T.W; then the H sweep).
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const float* other = half ? p.W : p.Ht;
            const int n_other = half ? p.F : p.T;
            const float* sX = half ? sXh : sXw;
            const int rp = half ? rph : rpw;
            float* gram = half ? gramW : gramH;
            float* sOwn = half ? sH : sW;
            const int n_own = half ? nt : nf;
            float* gOwn = half ? p.Ht + (long long)t0 * KP : p.W + (long long)f0 * KP;
            long long* tw = (dbg_on && half == 0) ? c_ph + 10 : nullptr;
            CO_TIC();
            switch (rp) {
                case 4: stream_products<KP, 1>(other, n_other, sX, ring, sGp, gram, tw); break;
                case 8: stream_products<KP, 2>(other, n_other, sX, ring, sGp, gram, tw); break;
                case 12: stream_products<KP, 3>(other, n_other, sX, ring, sGp, gram, tw); break;
                default: stream_products<KP, 4>(other, n_other, sX, ring, sGp, gram, tw); break;
            }
            CO_TOC(4 * half);
            grid_barrier(p.bar, nb_ctas, bar_target);
            CO_TOC(4 * half + 1);
            gram_load<KP>(gram, sG, sInv);
            v += sweep_rows<KP>(sOwn, ring, sG, sInv, n_own);
            __syncthreads();
            for (int i = tid; i < n_own * (KP / 4); i += kCoopThreads) {
                const int r = i / (KP / 4), c = (i % (KP / 4)) * 4;
                *reinterpret_cast<float4*>(gOwn + (long long)r * KP + c) = *reinterpret_cast<const float4*>(sOwn + r * AP + c);
            }
            if (half == 1) {
                const double vd = block_sum_d((double)v, s_red);
                if (tid == 0) p.viol_part[cta] = vd;
            }
            CO_TOC(4 * half + 2);
            grid_barrier(p.bar, nb_ctas, bar_target);
            CO_TOC(4 * half + 3);
        }
        // ---- stop rule: the same ordered sum in every CTA ----
        if (tid < 32) {
            double s = 0.0;
            for (int c = tid; c < nb_ctas; c += 32) s += __ldcg(p.viol_part + c);
            s = warp_sum_d(s);
            if (tid == 0) s_tot = s;
        }
        __syncthreads();
        const double tot = s_tot;
        CO_TOC(8);
        n_iter = it;
        if (it == 1) viol_init = tot;
        viol_last = tot;
        if (viol_init == 0.0 || tot / viol_init <= (double)p.tol) { done = 1; break; }
        // (the next write of viol_part is three barriers away)
    }
    if (dbg_on) for (int i = 0; i < 20; ++i) p.dbg[i] = (i == 9) ? n_iter : c_ph[i];
    if (cta == 0 && tid == 0) {
        ClipState s = p.st[0];
        s.n_iter = n_iter; s.viol_init = viol_init; s.viol_last = viol_last; s.done = done;
        p.st[0] = s;
    }
}

template <int KP> size_t coop_smem_bytes(int F, int T, int n_sm) {
    return sizeof(float) * (CoopCfg<KP>::fixed_floats + coop_x_floats(F, T, ceil_div(F, n_sm), ceil_div(T, n_sm)));
}
template <int KP> bool coop_fits(int F, int T, int n_sm) {
    const int rw = coop_pad4(ceil_div(F, n_sm)), rh = coop_pad4(ceil_div(T, n_sm));
    const int rmax = rw > rh ? rw : rh;
    if (rmax > kCoopMaxRows) return false;
    if ((size_t)CoopCfg<KP>::NJR * rmax * KP > (size_t)CoopCfg<KP>::STAGES * CoopCfg<KP>::STAGE_FLOATS) return false;   // the groups' sums fit the ring
    return coop_smem_bytes<KP>(F, T, n_sm) <= 225 * 1024;
}

template <int KP>
cudaError_t coop_launch(const NmfProblem& p, const NmfWork& wk, int max_iter, int n_sm, cudaStream_t s) {
    auto kern = nmf_coop_kernel<KP>;
    const size_t smem = coop_smem_bytes<KP>(p.F, p.T, n_sm);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    CoopParams cp;
    cp.X = p.Xt; cp.ldf = p.ldf; cp.F = p.F; cp.T = p.T; cp.W = p.W; cp.Ht = p.Ht; cp.st = p.state;
    cp.gram = wk.gram_partial;
    cp.viol_part = reinterpret_cast<double*>(wk.coop_scratch + 256);
    cp.bar = reinterpret_cast<unsigned*>(wk.coop_scratch);
    cp.max_iter = max_iter; cp.tol = p.tol;
    cp.rows_w = ceil_div(p.F, n_sm); cp.rows_h = ceil_div(p.T, n_sm);
    if ((e = cudaMemsetAsync(wk.coop_scratch, 0, 256, s)) != cudaSuccess) return e;
    static int dbg_left = -1;
    if (dbg_left < 0) { const char* ev = getenv("AINMF_COOP_DEBUG"); dbg_left = (ev && ev[0] == '1') ? 2 : 0; }
    cp.dbg = dbg_left > 0 ? reinterpret_cast<long long*>(wk.coop_scratch + 3584) : nullptr;
    void* args[] = {&cp};
    ++g_launch_count;
    e = cudaLaunchCooperativeKernel((const void*)kern, dim3(n_sm), dim3(kCoopThreads), args, smem, s);
    if (dbg_left > 0 && e == cudaSuccess) {
        --dbg_left;
        long long hb[20];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, cp.dbg, sizeof hb, cudaMemcpyDeviceToHost);
        const char* nm[9] = {"A stream Ht", "barrier 1", "B G+sweep+store", "barrier 2", "C stream W", "barrier 3", "D G+sweep+store+viol", "barrier 4", "stop sum"};
        const double n = hb[9] > 0 ? (double)hb[9] : 1.0;
        fprintf(stderr, "[coop-debug F=%d T=%d KP=%d rows %d/%d smem %zu] CTA 0, %lld iterations, cycles per iteration:", p.F, p.T, KP, cp.rows_w, cp.rows_h, smem, hb[9]);
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.0f;", nm[i], hb[i] / n);
        fprintf(stderr, " [A per iteration: issue %.0f, cp.async wait %.0f, barrier %.0f, arithmetic %.0f, barrier %.0f, sums after the stream %.0f, setup %.0f, first issues %.0f]\n",
                hb[10] / n, hb[11] / n, hb[12] / n, hb[13] / n, hb[14] / n, hb[15] / n, hb[16] / n, hb[17] / n);
    }
    return e;
}

}  // namespace

bool nmf_coop_eligible(const NmfProblem& p, const NmfWork& wk, int n_sm) {
    if (p.B != 1 || p.solver != 0 || p.t_good || wk.exact_viol || !wk.coop_scratch || wk.h_viol_sum) return false;
    if (n_sm < kCoopGramCtas || n_sm > 400) return false;                     // 128 CTAs share a Gram; viol_part holds 400 shares
    if (wk.gram_max_blocks < 2) return false;                                 // room in gram_partial for the two Grams
    switch (p.KP) {
        case 32: return coop_fits<32>(p.F, p.T, n_sm);
        case 64: return coop_fits<64>(p.F, p.T, n_sm);
        case 128: return coop_fits<128>(p.F, p.T, n_sm);
    }
    return false;
}

cudaError_t nmf_coop_fit(const NmfProblem& p, const NmfWork& wk, int max_iter, int n_sm, cudaStream_t s) {
    switch (p.KP) {
        case 32: return coop_launch<32>(p, wk, max_iter, n_sm, s);
        case 64: return coop_launch<64>(p, wk, max_iter, n_sm, s);
        case 128: return coop_launch<128>(p, wk, max_iter, n_sm, s);
    }
    return (cudaError_t)1;
}
#endif

}  // namespace ainmf
