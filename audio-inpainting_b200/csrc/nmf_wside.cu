// nmf_wside.cu -- the W half-step of the coordinate-descent iteration: ordered sum of the X.Ht partials, W sweep, stores
// of W / W^T / the bf16 cross operand, W^T W and the H step's sweep operands.
// $SP/sklearn/decomposition/_nmf.py:379-396 (X.Ht, HHt -> W sweep, _cdnmf_fast.pyx:8-38), then :379-380 for the H half (W^T W).
#include "kernels.h"
#include "nmf_cd.cuh"
#include "nmf_ts.cuh"

namespace ainmf {

// =====================================================================================================
// W side of the iteration in one kernel; grid = (ceil(F/ROWS), B), 128 threads, a group of L lanes per row of W:
//   B = sum_s partial[s] (fixed order);  g = W.G - B accumulated as rank-1 updates (rows of zeros are skipped);
//   W <- sweep(W, G = HHt, B) in the incremental-gradient form: the whole gradient lives in registers, a coordinate
//     step costs KP FMAs per row instead of a dot product, shuffles and the update logic on every lane;
//   the new rows stay in shared memory and feed: the store of W; (tensor-core path) W^T and its bf16 cross operand
//     for the H step's contraction (see tc::cross_pack8); the block's share of W^T W (register-tiled FFMA).
//   w_finish_kernel then sums the Gram partials in block order (deterministic) and (tensor-core path) derives the
//   H step's sweep operands from them (g_prep_block).
// $SP/sklearn/decomposition/_nmf.py:379-396 (X.Ht, HHt -> W sweep), then :379-380 for the H half (W^T W).
// =====================================================================================================
// packed fp32x2 FMA (sm_100a FFMA2: two lanes per issue slot); plain fmaf pairs elsewhere
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(AINMF_EMU)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
template <int KP, int L_, int ROWS_> struct WSideCfg {
    static constexpr int L = L_;                           // lanes per row: 1 for big batches (fewest instructions), up to 8 when
    static constexpr int SL = KP / L;                      // few rows must fill the machine (shorter serial chain per lane)
    static constexpr int ROWS = ROWS_;                     // rows per block: 256 (L = 1), 128, or 32 (with the most lanes) for a single clip
    static constexpr int RPT = (L == 1) ? 2 : 1;           // rows per thread: with one lane per row a thread sweeps two rows, so
    static constexpr int THREADS = ROWS * L / RPT;         // that every Gram value fetched from shared memory feeds two FFMA2
    static_assert(SL >= 4 && SL <= 64 && THREADS / 16 <= KP && THREADS >= 128, "lanes per row");
    // F = 2^m + 1 would leave every clip a block with a single row, which holds an SM slot for the latency of a whole
    // sweep: with one lane per row the last block of a clip takes up to TAIL rows beyond its ROWS, swept by one more warp
    // (one row per lane) that runs next to the block's other warps
    static constexpr int TAIL = (L == 1) ? kWSideTail : 0;
    static constexpr int LAUNCH_THREADS = THREADS + TAIL;
    static constexpr int TROWS = ROWS + TAIL;              // rows of the shared-memory tile (one more, all zero, when TAIL > 0:
                                                           // the unused second row of the extra warp's lanes)
    static constexpr int GP = KP + 4 * L;                  // Gram row pitch (load_gram_padded<KP, L>)
    static constexpr int AP = KP + 4;                      // row pitch of the W tile (16-byte aligned rows)
    static constexpr size_t smem_bytes = sizeof(float) * ((size_t)KP * GP + KP + (size_t)(TROWS + (TAIL ? 1 : 0)) * AP);
};
struct WSideTc {                      // outputs for the tensor-core H step; all null on the FFMA path
    float* Wt; float* WtX; long long wt_stride; int ldw;
    float* GX; float* blobs; float* scal;
    // good-first frame order: the bad frames' share of X.Ht is fill (x) hbad, and the H step needs v = fill^T.W
    const float* fill; long long fill_stride; const float* hbad; float* vpartial; float* vfill;
};
template <int KP, int LANES, int NROWS>
__global__ void __launch_bounds__(WSideCfg<KP, LANES, NROWS>::LAUNCH_THREADS, (LANES == 1) ? 2 : 1)
w_side_kernel(float* __restrict__ W, long long w_stride, int F, const float* __restrict__ G,
              const float* __restrict__ partial, int S, float* __restrict__ viol /*[B][gridDim.x]*/,
              float* __restrict__ gram_partial /*[B][gridDim.x][KP*KP]*/, WSideTc tc_out, const ClipState* __restrict__ st,
              float* __restrict__ pg_out /*[B][F][KP] or null*/) {
    using Cfg = WSideCfg<KP, LANES, NROWS>;
    constexpr int L = Cfg::L, SL = Cfg::SL, ROWS = Cfg::ROWS, GP = Cfg::GP, AP = Cfg::AP;
    AINMF_DYN_SMEM(smem_raw);
    float* sG = reinterpret_cast<float*>(smem_raw);       // [KP][GP]
    float* sInv = sG + KP * GP;                           // [KP]
    float* sA = sInv + KP;                                // [ROWS][AP]
    __shared__ float s_red[32];
    const int b = blockIdx.y, P = gridDim.x;
    if (st[b].done) return;
    const float* Gb = G + (long long)b * KP * KP;
    load_gram_padded<KP, L>(sG, Gb);
    for (int t = threadIdx.x; t < KP; t += blockDim.x) { const float d = Gb[t * KP + t]; sInv[t] = (d != 0.f) ? 1.0f / d : 0.f; }
    const int f0 = blockIdx.x * ROWS;
    // rows of this block: ROWS, except that the clip's last block takes what is left (at most TROWS, see nmf_plan)
    const int rows_here = ((int)blockIdx.x == P - 1) ? F - f0 : ROWS;
    float vsum = 0.f;
    if constexpr (L == 1) {
        // Big batches: the tile holds the rows of W, a thread sweeps rows r and r + ROWS/2.  The sweep is the reference's
        // (_cdnmf_fast.pyx:8-38) in blocks of 8 coordinates: the 8 gradients of a block are dot products of the row with
        // 8 rows of G (symmetric; shared memory, broadcast LDS.128 shared by the thread's two rows, packed FFMA2, 16
        // independent chains), started from -B, which is read from the partial sums (ordered over the splits) while the
        // dot products run; inside the block the coordinates are visited in order and each delta corrects the gradients
        // still to come (delta * G[t][t'], t' > t in the block).  In exact arithmetic this is the sequential sweep; K^2
        // FMAs per row.  Loops stay rolled (the row lives in shared memory, not in registers): ~400 instructions of code
        // instead of an unrolled 100 KB that no instruction cache holds.
        constexpr int HR = ROWS / 2;
        {
            const float4* Wg = reinterpret_cast<const float4*>(W + (long long)b * w_stride + (long long)f0 * KP);
            for (int i = threadIdx.x; i < (Cfg::TROWS + 1) * KP / 4; i += blockDim.x) {
                const int rr = (4 * i) / KP, cc = (4 * i) % KP;
                *reinterpret_cast<float4*>(sA + rr * AP + cc) = (rr < rows_here) ? Wg[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        // threads 0..THREADS-1: rows r and r + ROWS/2; the extra warp: row ROWS + lane (second row unused)
        const bool tailw = (int)threadIdx.x >= Cfg::THREADS;
        const int r = tailw ? ROWS + ((int)threadIdx.x - Cfg::THREADS) : (int)threadIdx.x;
        const int r1 = tailw ? Cfg::TROWS : r + HR;            // the extra warp's second row: the zero row, stays zero
        const bool valid[2] = {r < rows_here, !tailw && r1 < rows_here};
        float* ar[2] = {sA + r * AP, sA + r1 * AP};
        float fl[2] = {0.f, 0.f};
        if (tc_out.hbad) {
            if (valid[0]) fl[0] = tc_out.fill[(long long)b * tc_out.fill_stride + f0 + r];
            if (valid[1]) fl[1] = tc_out.fill[(long long)b * tc_out.fill_stride + f0 + r1];
        }
        // F = 2^m + 1: warps without a valid row skip the sweep
        if (__ballot_sync(0xffffffffu, valid[0]) != 0u) {
        // -B (the X.Ht sums) of coordinate block c: split 0 is fetched one block ahead, raw, so that no warp waits for
        // it (two warps per scheduler do not hide a global load); further splits (FFMA path only) are added in order
        const float* pbase[2] = {partial + (((long long)b * S) * F + f0 + r) * KP, partial + (((long long)b * S) * F + f0 + r1) * KP};
        float4 pn[2][2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            pn[rr][0] = pn[rr][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[rr]) { pn[rr][0] = *reinterpret_cast<const float4*>(pbase[rr]); pn[rr][1] = *reinterpret_cast<const float4*>(pbase[rr] + 4); }
        }
#pragma unroll 1
        for (int c = 0; c < KP; c += 8) {
            float4 pc[2][2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                pc[rr][0] = pn[rr][0]; pc[rr][1] = pn[rr][1];
                if (valid[rr] && c + 8 < KP) {
                    pn[rr][0] = *reinterpret_cast<const float4*>(pbase[rr] + c + 8);
                    pn[rr][1] = *reinterpret_cast<const float4*>(pbase[rr] + c + 12);
                }
            }
            // gradient i of the block = row c+i of G . a: x accumulates the even, y the odd columns
            float2 acc[2][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[0][i] = make_float2(0.f, 0.f); acc[1][i] = make_float2(0.f, 0.f); }
            const float* gc = sG + c * GP;
#pragma unroll 2
            for (int j = 0; j < KP; j += 4) {
                const float4 a0 = *reinterpret_cast<const float4*>(ar[0] + j), a1 = *reinterpret_cast<const float4*>(ar[1] + j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 u = *reinterpret_cast<const float4*>(gc + i * GP + j);
                    acc[0][i] = fma2(make_float2(a0.x, a0.y), make_float2(u.x, u.y), acc[0][i]);
                    acc[1][i] = fma2(make_float2(a1.x, a1.y), make_float2(u.x, u.y), acc[1][i]);
                    acc[0][i] = fma2(make_float2(a0.z, a0.w), make_float2(u.z, u.w), acc[0][i]);
                    acc[1][i] = fma2(make_float2(a1.z, a1.w), make_float2(u.z, u.w), acc[1][i]);
                }
            }
            float nb[2][8];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                nb[rr][0] = 0.f - pc[rr][0].x; nb[rr][1] = 0.f - pc[rr][0].y; nb[rr][2] = 0.f - pc[rr][0].z; nb[rr][3] = 0.f - pc[rr][0].w;
                nb[rr][4] = 0.f - pc[rr][1].x; nb[rr][5] = 0.f - pc[rr][1].y; nb[rr][6] = 0.f - pc[rr][1].z; nb[rr][7] = 0.f - pc[rr][1].w;
                if (valid[rr]) {
                    for (int sp = 1; sp < S; ++sp) {           // fixed order -> deterministic
                        const float* pr = pbase[rr] + (long long)sp * F * KP + c;
                        const float4 v0 = *reinterpret_cast<const float4*>(pr), v1 = *reinterpret_cast<const float4*>(pr + 4);
                        nb[rr][0] -= v0.x; nb[rr][1] -= v0.y; nb[rr][2] -= v0.z; nb[rr][3] -= v0.w;
                        nb[rr][4] -= v1.x; nb[rr][5] -= v1.y; nb[rr][6] -= v1.z; nb[rr][7] -= v1.w;
                    }
                }
            }
            if (tc_out.hbad) {                                 // bad frames: X.Ht += fill[f] * (sum of their rows of Ht)
                const float4 h0 = *reinterpret_cast<const float4*>(tc_out.hbad + (long long)b * KP + c);
                const float4 h1 = *reinterpret_cast<const float4*>(tc_out.hbad + (long long)b * KP + c + 4);
                const float hb[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int i = 0; i < 8; ++i) nb[rr][i] = fmaf(-fl[rr], hb[i], nb[rr][i]);
            }
            float gr[2][8], aq[2][8];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float4 q0 = *reinterpret_cast<const float4*>(ar[rr] + c), q1 = *reinterpret_cast<const float4*>(ar[rr] + c + 4);
                aq[rr][0] = q0.x; aq[rr][1] = q0.y; aq[rr][2] = q0.z; aq[rr][3] = q0.w;
                aq[rr][4] = q1.x; aq[rr][5] = q1.y; aq[rr][6] = q1.z; aq[rr][7] = q1.w;
#pragma unroll
                for (int i = 0; i < 8; ++i) gr[rr][i] = (acc[rr][i].x + acc[rr][i].y) + nb[rr][i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = c + i;
                const float inv = sInv[t];
                const float4 u0 = *reinterpret_cast<const float4*>(gc + i * GP + c), u1 = *reinterpret_cast<const float4*>(gc + i * GP + c + 4);
                const float gd[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const float grad = gr[rr][i];
                    const float a_ = aq[rr][i];
                    const float pg = (a_ == 0.f) ? fminf(0.f, grad) : grad;
                    vsum += fabsf(pg);                         // rows past the clip's last one: a = 0, B = 0 -> pg = 0
                    if (pg_out && valid[rr]) pg_out[((long long)b * F + f0 + (rr ? r1 : r)) * KP + t] = fabsf(pg);
                    const float an = (inv != 0.f) ? fmaxf(fmaf(-grad, inv, a_), 0.f) : a_;
                    const float d = an - a_;
                    aq[rr][i] = an;
#pragma unroll
                    for (int i2 = i + 1; i2 < 8; ++i2) gr[rr][i2] = fmaf(d, gd[i2], gr[rr][i2]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                *reinterpret_cast<float4*>(ar[rr] + c) = make_float4(aq[rr][0], aq[rr][1], aq[rr][2], aq[rr][3]);
                *reinterpret_cast<float4*>(ar[rr] + c + 4) = make_float4(aq[rr][4], aq[rr][5], aq[rr][6], aq[rr][7]);
            }
        }
        }
    } else {
    const int l = threadIdx.x % L, r = threadIdx.x / L;
    const int f = f0 + r;
    const bool valid = f < F;
    float* ar = sA + r * AP;
    float g[SL];
#pragma unroll
    for (int q = 0; q < SL; ++q) g[q] = 0.f;
    if (valid) {
        const float* wr = W + (long long)b * w_stride + (long long)f * KP + l * SL;
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            const float4 v = *reinterpret_cast<const float4*>(wr + q);
            ar[l * SL + q] = v.x; ar[l * SL + q + 1] = v.y; ar[l * SL + q + 2] = v.z; ar[l * SL + q + 3] = v.w;
        }
        for (int s = 0; s < S; ++s) {                      // fixed order -> deterministic
            const float* pr = partial + ((((long long)b * S + s) * F) + f) * KP + l * SL;
#pragma unroll
            for (int q = 0; q < SL; q += 4) {
                const float4 v = *reinterpret_cast<const float4*>(pr + q);
                g[q] -= v.x; g[q + 1] -= v.y; g[q + 2] -= v.z; g[q + 3] -= v.w;
            }
        }
        if (tc_out.hbad) {                                 // bad frames: X.Ht += fill[f] * (sum of their rows of Ht)
            const float fl = tc_out.fill[(long long)b * tc_out.fill_stride + f];
            const float* hb = tc_out.hbad + (long long)b * KP + l * SL;
#pragma unroll
            for (int q = 0; q < SL; ++q) g[q] = fmaf(-fl, hb[q], g[q]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < SL; ++q) ar[l * SL + q] = 0.f;
    }
    __syncthreads();
    const float* gl = sG + l * (SL + 4);
    auto rank1 = [&](float c, int t) {                     // g[:] += c * G[t][this lane's slice]
        const float* gr = gl + t * GP;
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(gr + q);
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(AINMF_EMU)
            const float2 cc = make_float2(c, c);
            const float2 lo2 = __ffma2_rn(cc, make_float2(gv.x, gv.y), make_float2(g[q], g[q + 1]));
            const float2 hi2 = __ffma2_rn(cc, make_float2(gv.z, gv.w), make_float2(g[q + 2], g[q + 3]));
            g[q] = lo2.x; g[q + 1] = lo2.y; g[q + 2] = hi2.x; g[q + 3] = hi2.y;
#else
            g[q] = fmaf(c, gv.x, g[q]); g[q + 1] = fmaf(c, gv.y, g[q + 1]);
            g[q + 2] = fmaf(c, gv.z, g[q + 2]); g[q + 3] = fmaf(c, gv.w, g[q + 3]);
#endif
        }
    };
    // gradient at the old W:  g = W.G - B  (G symmetric: row t of G scaled by W[row][t])
#pragma unroll 4
    for (int t = 0; t < KP; ++t) {
        const float c = ar[t];
        if (__ballot_sync(0xffffffffu, c != 0.f) == 0u) continue;       // warp-uniform
        rank1(c, t);
    }
    // the sweep (reference order t = 0..KP-1); the lane that owns coordinate t decides, everybody applies the delta
#pragma unroll
    for (int t = 0; t < KP; ++t) {
        const int o = t / SL, q = t % SL;
        const float inv = sInv[t];
        const float aq = ar[t];
        const float grad = g[q];
        const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
        const float an = fmaxf(fmaf(-grad, inv, aq), 0.f);
        const bool own = (l == o) && valid;
        const bool upd = own && (inv != 0.f);
        vsum += own ? fabsf(pg) : 0.f;
        if (pg_out && own) pg_out[((long long)b * F + f) * KP + t] = fabsf(pg);
        float d = upd ? an - aq : 0.f;
        if (upd) ar[t] = an;
        if (L > 1) d = __shfl_sync(0xffffffffu, d, o, L);
        if (__ballot_sync(0xffffffffu, d != 0.f) == 0u) continue;       // warp-uniform
        rank1(d, t);
    }
    }
    __syncthreads();

    // ---- the new rows: W ------------------------------------------------------------------------------------------
    float* Wb = W + (long long)b * w_stride + (long long)f0 * KP;
    for (int i = threadIdx.x; i < rows_here * KP / 4; i += blockDim.x)
        *reinterpret_cast<float4*>(Wb + 4 * i) = *reinterpret_cast<const float4*>(sA + ((4 * i) / KP) * AP + ((4 * i) % KP));
#ifndef AINMF_EMU
    if (tc_out.Wt) {
        // Wt[k][f] = W[f][k] (tf32 main-term operand) and WtX, the bf16 cross operand with the same footprint: per group of
        // 8 consecutive f, words 0-3 = pairs of bf16(w), words 4-7 = pairs of bf16(w - trunc_tf32(w))
        float* Wt = tc_out.Wt + (long long)b * tc_out.wt_stride;
        float* WtX = tc_out.WtX + (long long)b * tc_out.wt_stride;
        // one item = (component k, group of 8 rows): consecutive lanes take consecutive k (conflict-free reads of the
        // tile's columns) and write 32 contiguous bytes of row k of each operand
        for (int i = threadIdx.x; i < KP * (Cfg::TROWS / 8); i += blockDim.x) {
            const int k = i % KP, g8 = (i / KP) * 8, ff = f0 + g8;
            if (g8 >= rows_here || ff >= tc_out.ldw) continue;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = sA[(g8 + e) * AP + k];
            float4* wt = reinterpret_cast<float4*>(Wt + (long long)k * tc_out.ldw + ff);
            float4* wx = reinterpret_cast<float4*>(WtX + (long long)k * tc_out.ldw + ff);
            wt[0] = make_float4(v[0], v[1], v[2], v[3]);
            wx[0] = make_float4(__uint_as_float(tc::pack_bf16x2(v[0], v[1])), __uint_as_float(tc::pack_bf16x2(v[2], v[3])),
                                __uint_as_float(tc::pack_bf16x2(v[4], v[5])), __uint_as_float(tc::pack_bf16x2(v[6], v[7])));
            if (ff + 4 < tc_out.ldw) {
                float h, lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) tc::split_tf32(v[e], h, lo[e]);
                wt[1] = make_float4(v[4], v[5], v[6], v[7]);
                wx[1] = make_float4(__uint_as_float(tc::pack_bf16x2(lo[0], lo[1])), __uint_as_float(tc::pack_bf16x2(lo[2], lo[3])),
                                    __uint_as_float(tc::pack_bf16x2(lo[4], lo[5])), __uint_as_float(tc::pack_bf16x2(lo[6], lo[7])));
            }
        }
    }
#endif
    if (tc_out.vpartial) {
        for (int k = threadIdx.x; k < KP; k += blockDim.x) {
            float acc = 0.f;
            for (int rr = 0; rr < rows_here; ++rr) acc = fmaf(tc_out.fill[(long long)b * tc_out.fill_stride + f0 + rr], sA[rr * AP + k], acc);
            tc_out.vpartial[((long long)b * P + blockIdx.x) * KP + k] = acc;
        }
    }
    // ---- this block's share of W^T W: thread (ty, tx) of a (THREADS/16) x 16 grid owns rows ty*TI.., columns pass*16*TJ + tx*TJ.. ----
    if ((int)threadIdx.x < Cfg::THREADS) {
        constexpr int TY = Cfg::THREADS / 16, TI = KP / TY, TJ = (KP >= 64) ? 4 : 2, PASSES = KP / (16 * TJ);
        const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
        float* out = gram_partial + ((long long)b * P + blockIdx.x) * (KP * KP);
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
            float acc[TI][TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;
            const int c0 = pass * 16 * TJ + tx * TJ;
#pragma unroll 2
            for (int rr = 0; rr < rows_here; ++rr) {        // rows past the clip's last one are zero
                const float* row = sA + rr * AP;
                float fi[TI], fj[TJ];
                if constexpr (TI % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < TI; i += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(row + ty * TI + i);
                        fi[i] = v.x; fi[i + 1] = v.y; fi[i + 2] = v.z; fi[i + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TI; ++i) fi[i] = row[ty * TI + i];
                }
                if constexpr (TJ == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(row + c0);
                    fj[0] = v.x; fj[1] = v.y; fj[2] = v.z; fj[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < TJ; ++j) fj[j] = row[c0 + j];
                }
#pragma unroll
                for (int i = 0; i < TI; ++i) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(AINMF_EMU)
#pragma unroll
                    for (int j = 0; j < TJ; j += 2) {
                        const float2 r2 = __ffma2_rn(make_float2(fi[i], fi[i]), make_float2(fj[j], fj[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
                        acc[i][j] = r2.x; acc[i][j + 1] = r2.y;
                    }
#else
#pragma unroll
                    for (int j = 0; j < TJ; ++j) acc[i][j] = fmaf(fi[i], fj[j], acc[i][j]);
#endif
                }
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) out[(ty * TI + i) * KP + c0 + j] = acc[i][j];
        }
    }
    const float tot = block_sum(vsum, s_red);
    if (threadIdx.x == 0) viol[(long long)b * P + blockIdx.x] = tot;

}

// W^T W = sum of the W-side kernel's partials in block order (deterministic), 8 * gpb rows per block, and (tensor-core path)
// the H step's operands that derive from those rows (g_prep_block per group of 8); grid = (KP / (8 * gpb), B).  gpb > 1 for
// big batches: one block per clip instead of KP/8 (the kernel is a chain of dependent round trips; fewer, fatter blocks
// finish in one wave).
__global__ void __launch_bounds__(kThreads)
w_finish_kernel(const float* __restrict__ gram_partial /*[B][P][KP*KP]*/, int P, int KP, int gpb, float* __restrict__ WtW, WSideTc tc_out,
                const ClipState* __restrict__ st) {
    const int b = blockIdx.y, blk0 = blockIdx.x * gpb, n_el = 8 * gpb * KP;
    if (st[b].done) return;
    const float* pb = gram_partial + (long long)b * P * (KP * KP) + 8 * blk0 * KP;
    float* Go = WtW + (long long)b * (KP * KP);
    // block-ordered sums (deterministic); the loads of a group of partials are issued together -- with one long signal
    // this kernel is KP/8 blocks and its duration is the latency of P dependent round trips otherwise
    for (int e0 = 0; e0 < n_el; e0 += 4 * blockDim.x) {
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q0 = 0; q0 < P; q0 += 8) {
            float v[8][4];
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int e = e0 + m * blockDim.x + threadIdx.x;
                    v[q][m] = (q0 + q < P && e < n_el) ? pb[(long long)(q0 + q) * (KP * KP) + e] : 0.f;
                }
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int m = 0; m < 4; ++m) sum[m] += v[q][m];
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int e = e0 + m * blockDim.x + threadIdx.x;
            if (e < n_el) Go[8 * blk0 * KP + e] = sum[m];
        }
    }
#ifndef AINMF_EMU
    if (tc_out.blobs) {
        __syncthreads();                                 // the rows just written are read back by other threads of the block
        for (int sub = 0; sub < gpb; ++sub) {
            const int blk = blk0 + sub;
            g_prep_block(Go, tc_out.GX + (long long)b * KP * KP, tc_out.blobs + ((long long)b * (KP / 8) + blk) * (16 * KP),
                         tc_out.scal + ((long long)b * (KP / 8) + blk) * TS_SC, KP, blk, threadIdx.x, blockDim.x);
        }
        if (tc_out.vpartial) {                           // v[k] = fill^T.W: the X^T.W row of every bad frame
            for (int k = threadIdx.x; k < 8 * gpb; k += blockDim.x) {
                float v = 0.f;
                for (int q = 0; q < P; ++q) v += tc_out.vpartial[((long long)b * P + q) * KP + 8 * blk0 + k];
                tc_out.vfill[(long long)b * KP + 8 * blk0 + k] = v;
            }
        }
    }
#endif
}


template <int KP>
static cudaError_t launch_w_side_impl(const NmfProblem& p, const NmfWork& wk, int S, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    WSideTc tco{nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
    if (wk.use_tc) tco = WSideTc{wk.tc_Wt, wk.tc_WtLo, (long long)KP * p.ldf, p.ldf, wk.tc_GLo, wk.tc_blobs, wk.tc_scal,
                                 p.t_good ? p.fill : nullptr, p.fill_stride, p.t_good ? wk.tc_hbad : nullptr,
                                 p.t_good ? wk.tc_vpartial : nullptr, wk.tc_vfill};
    auto launch = [&](auto cfg) -> cudaError_t {
        using WC = decltype(cfg);
        auto kern = w_side_kernel<KP, WC::L, WC::ROWS>;
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WC::smem_bytes);
        if (e2 != cudaSuccess) return e2;
        AINMF_LAUNCH(kern, dim3(wk.nW, p.B), dim3(WC::LAUNCH_THREADS), WC::smem_bytes, s, p.W, p.w_stride, p.F, wk.HHt,
                     wk.xht_reduced ? wk.xht_reduced : wk.xht_partial, wk.xht_reduced ? 1 : S, wk.violW, wk.gram_partial,
                     tco, p.state, wk.exact_viol ? wk.pgW : nullptr);
        return cudaGetLastError();
    };
    constexpr int LMIN = (KP == 128) ? 2 : 1, LMAX = (KP == 32) ? 4 : 8;
    int lanes = wk.w_lanes < LMIN ? LMIN : (wk.w_lanes > LMAX ? LMAX : wk.w_lanes);
    if (wk.w_rows == 32) e = launch(WSideCfg<KP, LMAX, 32>{});
    else if (lanes >= 8) { if constexpr (LMAX >= 8) e = launch(WSideCfg<KP, 8, 128>{}); }
    else if (lanes >= 4) e = launch(WSideCfg<KP, 4, 128>{});
    else if (lanes >= 2) e = launch(WSideCfg<KP, 2, 128>{});
    else { if constexpr (LMIN <= 1) e = launch(WSideCfg<KP, 1, 256>{}); }
    if (e != cudaSuccess) return e;
    const int gpb = (wk.finish_gpb >= 1 && (KP / 8) % wk.finish_gpb == 0) ? wk.finish_gpb : 1;
    AINMF_LAUNCH(w_finish_kernel, dim3(KP / 8 / gpb, p.B), dim3(kThreads), 0, s, wk.gram_partial, wk.nW, KP, gpb, wk.WtW, tco, p.state);
    return cudaGetLastError();
}

cudaError_t launch_w_side(const NmfProblem& p, const NmfWork& wk, int S, cudaStream_t s) {
    switch (p.KP) {
#ifndef AINMF_WSIDE_DEV64              // development switch: compile the K = 64 instantiations only
        case 32: return launch_w_side_impl<32>(p, wk, S, s);
        case 128: return launch_w_side_impl<128>(p, wk, S, s);
#endif
        case 64: return launch_w_side_impl<64>(p, wk, S, s);
    }
    return (cudaError_t)1;
}

}  // namespace ainmf
