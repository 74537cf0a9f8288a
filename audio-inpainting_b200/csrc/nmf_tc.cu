// nmf_tc.cu -- the two V-sized contractions of the CD-NMF iteration on the 5th-generation tensor cores (sm_100a):
//   xht_tc_kernel   : [X | Ht]^T Ht  -> X.Ht partials (F x KP) and the Gram Ht^T Ht (KP x KP), split over time
//   h_step_tc_kernel: X^T W tile (128 frames x KP) in TMEM -> shared -> the row-parallel CD sweep of nmf_cd.cuh
// Operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), multiplied with tcgen05.mma kind::tf32 and
// accumulated in fp32 in TMEM.  A single TF32 pass is not accurate enough for the reference's tolerances
// (objective drifts 1e-3), so every product is error-compensated:
//       a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   a_hi = a with the low 13 mantissa bits cleared, a_lo = a - a_hi
// The tensor core itself ignores the low 13 bits of a 32-bit operand (measured: tests/test_gpu_tc.py), so the raw
// fp32 tile serves as a_hi; converter warps write a_lo beside it (generic proxy -> fence.proxy.async -> MMA).
// ncu on the FFMA kernels (profiles/r01_summary.md) shows the path compute-bound at K >= 64, which is the condition
// north_star sets for using tcgen05 here.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 4-7 = converters and TMEM->register epilogue, then all 8 warps run the sweep.
#include "kernels.h"
#include "nmf_cd.cuh"
#include <stdlib.h>

#include "tc.cuh"

#ifndef AINMF_EMU
namespace ainmf {
using namespace tc;

constexpr int TC_BK = 32;          // contraction elements per stage (one 128-byte row)
constexpr int TC_M = 128;          // MMA M
constexpr int TC_CONV_THREADS = 192;   // warps 2..7

// NSTAGE: h step KP=64 uses 2 stages (96 KB) so that two blocks share an SM and one block's sweep overlaps the
// other's GEMM; KP=128 needs 64 KB per stage and runs one block per SM with 3 stages.
template <int KP, int NSTAGE_> struct TcCfg {
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int A_BYTES = TC_M * TC_BK * 4;               // 16 KB
    static constexpr int B_BYTES = KP * TC_BK * 4;                 // 8 / 16 KB
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // raw + lo of both operands
    static constexpr int L = 8, R = 4;                             // sweep: 8 lanes x 4 frames per group, one pass
    static constexpr int CPITCH = KP + 4;
    static constexpr int GPITCH = KP + 4 * L;
    static constexpr int EPI_SHFL_BYTES = (TC_M * CPITCH + KP * GPITCH + KP) * 4;
    static constexpr int EPI_BLK_BYTES = (2 * KP * KP + 2 * TC_M * 8 + KP) * 4;     // G, G_lo (swizzled), delta hi/lo, 1/diag
    static constexpr int EPI_BYTES = EPI_SHFL_BYTES > EPI_BLK_BYTES ? EPI_SHFL_BYTES : EPI_BLK_BYTES;
    static constexpr int PIPE_BYTES = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_BYTES = (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) + 1024;   // + alignment slack
};
template <int KP> struct HStepStages { static constexpr int value = (KP == 64) ? 2 : 3; };
template <int KP> struct XhtStages { static constexpr int value = (KP == 64) ? 4 : 3; };

struct TcBarriers {
    uint64_t full[4], conv[4], empty[4], accum;
    uint64_t gload, dready, ddone;      // blocked sweep: Gram tiles landed / delta tile written / rank-8 update done
};

// a_lo = a - trunc_tf32(a) for `n4` float4 of a tile; the raw tile is left in place as a_hi.  Loads are issued in
// batches of 4 before the dependent arithmetic so that the shared-memory latency is paid once per batch.
__device__ __forceinline__ void write_lo(const float* __restrict__ raw, float* __restrict__ lo, int n4, int tid, int nthreads) {
    const float4* src = reinterpret_cast<const float4*>(raw);
    float4* dst = reinterpret_cast<float4*>(lo);
    int i = tid;
    for (; i + 3 * nthreads < n4; i += 4 * nthreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = src[i + u * nthreads];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 l;
            float h;
            split_tf32(v[u].x, h, l.x); split_tf32(v[u].y, h, l.y); split_tf32(v[u].z, h, l.z); split_tf32(v[u].w, h, l.w);
            dst[i + u * nthreads] = l;
        }
    }
    for (; i < n4; i += nthreads) {
        const float4 v = src[i];
        float4 l;
        float h;
        split_tf32(v.x, h, l.x); split_tf32(v.y, h, l.y); split_tf32(v.z, h, l.z); split_tf32(v.w, h, l.w);
        dst[i] = l;
    }
}

// =====================================================================================================
// h step: grid = (ceil(T/128), B).  The accumulator collects  D = X_tile.W  -  Ht_tile.(W^T W)  = -(gradient of the
// H half-step at the old Ht), so the sweep starts from the gradient and never recomputes a dot product:
//   stages 0 .. nkX-1      : A = X chunk (K-major; raw + lo by the converters), B = Wt, Wt_lo chunks (precomputed)
//   stages nkX .. nkX+KP/32: A = Ht chunk (K-major), B = W^T W chunk; both lo tiles by the converters; A negated
// Stage layout: [A raw][A lo][B raw][B lo].
// =====================================================================================================
template <int KP, bool BLK>
__global__ void __launch_bounds__(kThreads, (KP == 64) ? 2 : 1)
h_step_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWt,
                 const __grid_constant__ CUtensorMap mapWtLo, const __grid_constant__ CUtensorMap mapHk,
                 const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapGlo, int F, int T,
                 const float* __restrict__ G,
                 float* __restrict__ Ht, long long h_stride, float* __restrict__ viol, const ClipState* __restrict__ st,
                 long long* __restrict__ dbg) {
    using Cfg = TcCfg<KP, HStepStages<KP>::value>;
    constexpr int L = Cfg::L, R = Cfg::R, SL = KP / L, CPITCH = Cfg::CPITCH;
    const bool dbg_on = dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0;
    const long long dbg_t0 = clock64();
#define DBG(slot) do { if (dbg_on) dbg[slot] = clock64() - dbg_t0; } while (0)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) TcBarriers bars;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_red[32];
    const int b = blockIdx.y;
    if (st[b].done) return;
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_M;
    const int nkX = (F + TC_BK - 1) / TC_BK;
    const int nk = nkX + KP / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.conv[s], TC_CONV_THREADS); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.accum, 1);
        mbar_init(&bars.gload, 1); mbar_init(&bars.dready, 128); mbar_init(&bars.ddone, 1);
        mbar_fence_init();
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWt); tma_prefetch_desc(&mapWtLo); tma_prefetch_desc(&mapHk); tma_prefetch_desc(&mapG);
        tma_prefetch_desc(&mapGlo);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, KP);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    // Role loops.  Whole warps iterate (only lane 0 of warps 0/1 acts) so that the other lanes wait at the warp
    // barrier instead of polling an mbarrier and stealing issue slots from the converters.
    if (warp == 0) {
        for (int i = 0; i < nk; ++i) {
            if (lane == 0) {
                const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
                mbar_wait(&bars.empty[s], ph ^ 1);
                DBG(8 + 6 * i + 5);
                unsigned char* stg = smem + (size_t)s * Cfg::STAGE_BYTES;
                if (i < nkX) {
                    mbar_arrive_expect_tx(&bars.full[s], Cfg::A_BYTES + 2 * Cfg::B_BYTES);
                    tma_load_3d(stg, &mapX, &bars.full[s], i * TC_BK, m0, b);
                    tma_load_3d(stg + 2 * Cfg::A_BYTES, &mapWt, &bars.full[s], i * TC_BK, 0, b);
                    tma_load_3d(stg + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapWtLo, &bars.full[s], i * TC_BK, 0, b);
                } else {
                    mbar_arrive_expect_tx(&bars.full[s], Cfg::A_BYTES + 2 * Cfg::B_BYTES);
                    tma_load_3d(stg, &mapHk, &bars.full[s], (i - nkX) * TC_BK, m0, b);
                    tma_load_3d(stg + 2 * Cfg::A_BYTES, &mapG, &bars.full[s], (i - nkX) * TC_BK, 0, b);
                    tma_load_3d(stg + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapGlo, &bars.full[s], (i - nkX) * TC_BK, 0, b);
                }
                DBG(8 + 6 * i + 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(TC_M, KP, 0, 0);
        const uint32_t idesc_neg = idesc | (1u << 13);              // negate A: subtracts Ht.(W^T W)
        uint32_t acc = 0;
        for (int i = 0; i < nk; ++i) {
            if (lane == 0) {
                const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
                mbar_wait(&bars.conv[s], ph);
                DBG(8 + 6 * i + 4);
                tcgen05_fence_after();
                const uint32_t a_raw = smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
                const uint32_t id = (i < nkX) ? idesc : idesc_neg;
                // descriptors differ only in the 14-bit start-address field: build one, add byte offsets >> 4
                const uint64_t d_ar = make_smem_desc(a_raw, 16, 1024);
                const uint64_t d_al = d_ar + (uint64_t)(Cfg::A_BYTES >> 4);
                const uint64_t d_br = d_ar + (uint64_t)((2 * Cfg::A_BYTES) >> 4);
                const uint64_t d_bl = d_br + (uint64_t)(Cfg::B_BYTES >> 4);
#pragma unroll
                for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
                    const uint64_t o = (uint64_t)(k8 * 32 >> 4);
                    mma_tf32_ss(tmem, d_ar + o, d_br + o, id, acc);
                    acc = 1;
                    mma_tf32_ss(tmem, d_ar + o, d_bl + o, id, 1);
                    mma_tf32_ss(tmem, d_al + o, d_br + o, id, 1);
                }
                mma_commit(&bars.empty[s]);
                if (i == nk - 1) mma_commit(&bars.accum);
                DBG(8 + 6 * i + 3);
            }
            __syncwarp();
        }
    } else {
        const int ct = threadIdx.x - 64;
        for (int i = 0; i < nk; ++i) {
            const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
            if (lane == 0) mbar_wait(&bars.full[s], ph);     // one poller per warp; the other lanes sleep at the warp barrier
            __syncwarp();
            if (ct == 0) DBG(8 + 6 * i + 1);
            float* raw = reinterpret_cast<float*>(smem + (size_t)s * Cfg::STAGE_BYTES);
            write_lo(raw, raw + Cfg::A_BYTES / 4, Cfg::A_BYTES / 16, ct, TC_CONV_THREADS);
            fence_proxy_async_smem();
            mbar_arrive(&bars.conv[s]);
            if (ct == 0) DBG(8 + 6 * i + 2);
        }
    }
    if (threadIdx.x == 0) DBG(0);
    if constexpr (BLK) {
        // ---- blocked Gauss-Seidel sweep on the accumulator ------------------------------------------------------
        // The accumulator holds D = -(gradient).  Coordinates are processed in blocks of 8: inside a block one thread
        // per frame updates sequentially (corrections from the block's own deltas in registers); the effect of the
        // block on all later coordinates, D -= delta[128x8] . G[8 x KP], is one (error-compensated) K = 8 tensor-core
        // MMA.  No shuffles, no shared-memory traffic proportional to K^2, and D never leaves TMEM.
        unsigned char* sGr = smem;                                       // G as K-major SWIZZLE_128B chunks [KP/32][KP][128 B]
        unsigned char* sGl = smem + (size_t)KP * KP * 4;                 // its TF32 residual, same layout
        float* sDh = reinterpret_cast<float*>(smem + (size_t)2 * KP * KP * 4);   // delta tile 128 x 8, core-matrix interleaved
        float* sDl = sDh + TC_M * 8;
        float* sInv = sDl + TC_M * 8;
        __syncthreads();                                               // idle lanes park here
        if (lane == 0) mbar_wait(&bars.accum, 0);                      // every MMA has finished reading the stage buffers
        __syncwarp();
        tcgen05_fence_after();
        if (threadIdx.x == 0) DBG(1);
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bars.gload, 2 * KP * KP * 4);
            for (int c = 0; c < KP / 32; ++c) {
                tma_load_3d(sGr + (size_t)c * KP * 128, &mapG, &bars.gload, c * 32, 0, b);
                tma_load_3d(sGl + (size_t)c * KP * 128, &mapGlo, &bars.gload, c * 32, 0, b);
            }
        }
        if (lane == 0) mbar_wait(&bars.gload, 0);
        __syncwarp();
        auto g_at = [&](int n, int c) -> float {                       // G[n][c] from the swizzled copy
            const int cc = c & 31;
            return *reinterpret_cast<const float*>(sGr + (size_t)(c >> 5) * KP * 128 + n * 128 + ((((cc >> 2) ^ (n & 7))) << 4) + ((cc & 3) << 2));
        };
        for (int t = threadIdx.x; t < KP; t += blockDim.x) { const float d = g_at(t, t); sInv[t] = (d != 0.f) ? 1.0f / d : 0.f; }
        __syncthreads();
        if (threadIdx.x == 0) DBG(2);
        float vsum = 0.f;
        constexpr int NBLK = KP / 8;
        if (warp >= 4) {
            const int q = warp & 3, row = q * 32 + lane;
            const int t = m0 + row;
            const bool valid = t < T;
            float* hrow = Ht + (long long)b * h_stride + (long long)t * KP;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
            const int doff = (row >> 3) * 64 + (row & 7) * 4;           // floats: 8-row core matrices of 16 B, K-adjacent cores 128 B apart
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                float d8[8], a8[8], dl[8];
                tmem_ld_32x8(taddr + blk * 8, d8);
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (valid) { v0 = *reinterpret_cast<const float4*>(hrow + 8 * blk); v1 = *reinterpret_cast<const float4*>(hrow + 8 * blk + 4); }
                a8[0] = v0.x; a8[1] = v0.y; a8[2] = v0.z; a8[3] = v0.w; a8[4] = v1.x; a8[5] = v1.y; a8[6] = v1.z; a8[7] = v1.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int tt = 8 * blk + j;
                    float grad = -d8[j];
#pragma unroll
                    for (int i = 0; i < j; ++i) grad = fmaf(g_at(tt, 8 * blk + i), dl[i], grad);
                    const float inv = sInv[tt];
                    const float aq = a8[j];
                    const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
                    const float an = fmaxf(fmaf(-grad, inv, aq), 0.f);
                    const bool upd = valid && (inv != 0.f);
                    vsum += valid ? fabsf(pg) : 0.f;
                    dl[j] = upd ? an - aq : 0.f;
                    a8[j] = upd ? an : aq;
                }
                if (valid) {
                    *reinterpret_cast<float4*>(hrow + 8 * blk) = make_float4(a8[0], a8[1], a8[2], a8[3]);
                    *reinterpret_cast<float4*>(hrow + 8 * blk + 4) = make_float4(a8[4], a8[5], a8[6], a8[7]);
                }
                if (blk + 1 < NBLK) {
                    float hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split_tf32(dl[j], hi[j], lo[j]);
                    *reinterpret_cast<float4*>(sDh + doff) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(sDh + doff + 32) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<float4*>(sDl + doff) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(sDl + doff + 32) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                    fence_proxy_async_smem();
                    mbar_arrive(&bars.dready);
                    if (lane == 0) mbar_wait(&bars.ddone, blk & 1);
                    __syncwarp();
                    tcgen05_fence_after();
                }
            }
        } else if (warp == 1) {
            const uint32_t idn = make_idesc_tf32(TC_M, KP, 0, 0) | (1u << 13);      // D -= delta . G
            const uint64_t dAh = make_smem_desc(smem_u32(sDh), 128, 256, 0);          // K-major, no swizzle
            const uint64_t dAl = make_smem_desc(smem_u32(sDl), 128, 256, 0);
#pragma unroll 1
            for (int blk = 0; blk + 1 < NBLK; ++blk) {
                if (lane == 0) {
                    mbar_wait(&bars.dready, blk & 1);
                    tcgen05_fence_after();
                    const uint32_t boff = (uint32_t)(blk >> 2) * KP * 128 + (uint32_t)(blk & 3) * 32;
                    const uint64_t dBr = make_smem_desc(smem_u32(sGr) + boff, 16, 1024);
                    const uint64_t dBl = make_smem_desc(smem_u32(sGl) + boff, 16, 1024);
                    mma_tf32_ss(tmem, dAh, dBr, idn, 1);
                    mma_tf32_ss(tmem, dAh, dBl, idn, 1);
                    mma_tf32_ss(tmem, dAl, dBr, idn, 1);
                    mma_commit(&bars.ddone);
                }
                __syncwarp();
            }
        }
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, KP);
        const float tot = block_sum(vsum, s_red);
        if (threadIdx.x == 0) viol[(long long)b * gridDim.x + blockIdx.x] = tot;
        if (threadIdx.x == 0) DBG(3);
        return;
    }
    // ---- epilogue: accumulator -> shared, Gram -> shared, sweep --------------------------------------------
    float* sC = reinterpret_cast<float*>(smem);                    // [128][CPITCH] (aliases the pipeline buffers)
    float* sG = sC + TC_M * CPITCH;                                // [KP][GPITCH]
    float* sInv = sG + KP * Cfg::GPITCH;                           // [KP] reciprocal of the Gram diagonal
    __syncthreads();                                               // idle lanes park here (hardware barrier, no polling)
    if (lane == 0) mbar_wait(&bars.accum, 0);                      // every MMA has finished reading the stage buffers
    __syncwarp();
    tcgen05_fence_after();
    if (threadIdx.x == 0) DBG(1);
    if (warp >= 4) {
        const int q = warp & 3;                                    // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(sC + row * CPITCH + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    } else {
        // warps 0-3 stage the Gram matrix meanwhile (disjoint region)
        constexpr int S = KP / L, PITCH = Cfg::GPITCH;
        const float* Gb = G + (long long)b * KP * KP;
        for (int i = threadIdx.x; i < KP * KP / 4; i += 128) {
            const int t = (4 * i) / KP, r = (4 * i) % KP;
            *reinterpret_cast<float4*>(sG + t * PITCH + (r / S) * (S + 4) + (r % S)) = *reinterpret_cast<const float4*>(Gb + 4 * i);
        }
        for (int t = threadIdx.x; t < KP; t += 128) {
            const float d = Gb[t * KP + t];
            sInv[t] = (d != 0.f) ? 1.0f / d : 0.f;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, KP);
    if (threadIdx.x == 0) DBG(2);

    // sweep: a group of L lanes owns R consecutive frames; 256 threads = 32 groups x 4 frames = the 128-frame tile
    const int l = threadIdx.x % L, grp = threadIdx.x / L;
    float* Hb = Ht + (long long)b * h_stride;
    float a[R][SL], g[R][SL];
    bool valid[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int r = grp * R + i;
        const int t = m0 + r;
        valid[i] = t < T;
        const float* cr = sC + r * CPITCH + l * SL;
#pragma unroll
        for (int q = 0; q < SL; q += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[i]) v = *reinterpret_cast<const float4*>(Hb + (long long)t * KP + l * SL + q);
            a[i][q] = v.x; a[i][q + 1] = v.y; a[i][q + 2] = v.z; a[i][q + 3] = v.w;
            const float4 c = *reinterpret_cast<const float4*>(cr + q);
            g[i][q] = -c.x; g[i][q + 1] = -c.y; g[i][q + 2] = -c.z; g[i][q + 3] = -c.w;      // gradient = -(accumulator)
        }
    }
    const float vsum = cd_sweep_rows_inc<KP, L, R>(a, g, sG, sInv, l, valid);
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (valid[i]) {
            float* hr = Hb + (long long)(m0 + grp * R + i) * KP + l * SL;
#pragma unroll
            for (int q = 0; q < SL; q += 4) *reinterpret_cast<float4*>(hr + q) = make_float4(a[i][q], a[i][q + 1], a[i][q + 2], a[i][q + 3]);
        }
    }
    const float tot = block_sum(vsum, s_red);
    if (threadIdx.x == 0) viol[(long long)b * gridDim.x + blockIdx.x] = tot;
    if (threadIdx.x == 0) DBG(3);
#undef DBG
}

// =====================================================================================================
// xht + gram: grid = (m_tiles, S, B).  Both operands are MN-major (time is the slow dimension of Xt and Ht), which
// for tf32 requires the 32-byte-atom 128B swizzle.  The A operand of tile `mt` is the list of 32-column slabs
// v = 4*mt + j of the virtual matrix [ X (ceil(F/32) slabs) | Ht (KP/32 slabs) ]; slabs past the end load zeros.
// Stage layout: [A raw (4 slabs)][A lo][B raw (KP/32 slabs)][B lo]; both lo tiles come from the converters.
// =====================================================================================================
template <int KP>
__global__ void __launch_bounds__(kThreads, 1)
xht_tc_kernel(const __grid_constant__ CUtensorMap mapXmn, const __grid_constant__ CUtensorMap mapHmn, int F, int T,
              int frames_per_split, float* __restrict__ xht_partial /*[B][S][F][KP]*/,
              float* __restrict__ gram_partial /*[B][S][KP][KP]*/, const ClipState* __restrict__ st) {
    using Cfg = TcCfg<KP, XhtStages<KP>::value>;
    constexpr int SLAB = 32 * TC_BK * 4;                           // 4 KB: 32 columns x 32 frames
    constexpr int NB = KP / 32;                                    // B slabs
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) TcBarriers bars;
    __shared__ uint32_t tmem_slot;
    const int b = blockIdx.z, split = blockIdx.y, S = gridDim.y, mt = blockIdx.x;
    if (st[b].done) return;
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nxs = (F + 31) / 32;                                 // X slabs of the virtual matrix
    const int t_begin = split * frames_per_split;
    const int t_end = min(T, t_begin + frames_per_split);
    const int nk = (t_end - t_begin + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.conv[s], TC_CONV_THREADS); mbar_init(&bars.empty[s], 1); }
        mbar_init(&bars.accum, 1);
        mbar_fence_init();
        tma_prefetch_desc(&mapXmn); tma_prefetch_desc(&mapHmn);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, KP);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        for (int i = 0; i < nk; ++i) {
            if (lane == 0) {
                const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
                mbar_wait(&bars.empty[s], ph ^ 1);
                unsigned char* stg = smem + (size_t)s * Cfg::STAGE_BYTES;
                const int t0 = t_begin + i * TC_BK;
                // rows past t_end of this split must not contribute: the split boundary is a multiple of 32 except
                // at T, where TMA zero-fills (d1 = T)
                mbar_arrive_expect_tx(&bars.full[s], Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int v = 4 * mt + j;
                    if (v < nxs) tma_load_3d(stg + j * SLAB, &mapXmn, &bars.full[s], 32 * v, t0, b);
                    else tma_load_3d(stg + j * SLAB, &mapHmn, &bars.full[s], 32 * (v - nxs), t0, b);   // >= KP: zero fill
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) tma_load_3d(stg + 2 * Cfg::A_BYTES + j * SLAB, &mapHmn, &bars.full[s], 32 * j, t0, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(TC_M, KP, 1, 1);
        uint32_t acc = 0;
        for (int i = 0; i < nk; ++i) {
            if (lane == 0) {
                const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
                mbar_wait(&bars.conv[s], ph);
                tcgen05_fence_after();
                const uint32_t a_raw = smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
                const uint32_t a_lo = a_raw + Cfg::A_BYTES, b_raw = a_raw + 2 * Cfg::A_BYTES, b_lo = b_raw + Cfg::B_BYTES;
#pragma unroll
                for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
                    const uint32_t o = k8 * 1024;                  // 8 frames x 128 B
                    const uint64_t dar = make_smem_desc(a_raw + o, SLAB, 512, kLayoutSw128Base32);
                    const uint64_t dal = make_smem_desc(a_lo + o, SLAB, 512, kLayoutSw128Base32);
                    const uint64_t dbr = make_smem_desc(b_raw + o, SLAB, 512, kLayoutSw128Base32);
                    const uint64_t dbl = make_smem_desc(b_lo + o, SLAB, 512, kLayoutSw128Base32);
                    mma_tf32_ss(tmem, dar, dbr, idesc, acc);
                    acc = 1;
                    mma_tf32_ss(tmem, dar, dbl, idesc, 1);
                    mma_tf32_ss(tmem, dal, dbr, idesc, 1);
                }
                mma_commit(&bars.empty[s]);
                if (i == nk - 1) mma_commit(&bars.accum);
            }
            __syncwarp();
        }
    } else {
        const int ct = threadIdx.x - 64;
        for (int i = 0; i < nk; ++i) {
            const int s = i % Cfg::NSTAGE, ph = (i / Cfg::NSTAGE) & 1;
            if (lane == 0) mbar_wait(&bars.full[s], ph);
            __syncwarp();
            float* raw = reinterpret_cast<float*>(smem + (size_t)s * Cfg::STAGE_BYTES);
            write_lo(raw, raw + Cfg::A_BYTES / 4, Cfg::A_BYTES / 16, ct, TC_CONV_THREADS);
            write_lo(raw + 2 * Cfg::A_BYTES / 4, raw + (2 * Cfg::A_BYTES + Cfg::B_BYTES) / 4, Cfg::B_BYTES / 16, ct, TC_CONV_THREADS);
            fence_proxy_async_smem();
            mbar_arrive(&bars.conv[s]);
        }
    }
    // ---- epilogue: rows of the accumulator straight to the partial buffers -------------------------------
    __syncthreads();                                               // idle lanes park here
    if (warp >= 4) {
        if (nk > 0 && lane == 0) mbar_wait(&bars.accum, 0);
        __syncwarp();
        tcgen05_fence_after();
        const int q = warp & 3;
        const int vcol = 128 * mt + q * 32 + lane;                 // virtual column = output row
        float* dst = nullptr;
        if (vcol < 32 * nxs) {
            if (vcol < F) dst = xht_partial + ((((long long)b * S + split) * F) + vcol) * KP;
        } else if (vcol - 32 * nxs < KP) {
            dst = gram_partial + ((((long long)b * S + split) * KP) + (vcol - 32 * nxs)) * KP;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
            float v[32];
            if (nk > 0) {
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (dst) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, KP);
}

// Wt[b][k][f] = W[b][f][k] and Wt_lo = Wt - trunc_tf32(Wt); 32x32 tiles; grid = (ceil(F/32), KP/32, B)
__global__ void __launch_bounds__(kThreads)
wt_split_kernel(const float* __restrict__ W, long long w_stride, int F, int KP, int ldw, float* __restrict__ Wt,
                float* __restrict__ WtLo, long long wt_stride, const ClipState* __restrict__ st) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    if (st[b].done) return;
    const int f0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int r = ly; r < 32; r += 8) {
        const int f = f0 + r;
        tile[r][lx] = (f < F) ? W[(long long)b * w_stride + (long long)f * KP + k0 + lx] : 0.f;
    }
    __syncthreads();
    for (int r = ly; r < 32; r += 8) {
        const int k = k0 + r, f = f0 + lx;
        if (f < ldw) {
            const float v = tile[lx][r];
            float h, l;
            split_tf32(v, h, l);
            Wt[(long long)b * wt_stride + (long long)k * ldw + f] = v;
            WtLo[(long long)b * wt_stride + (long long)k * ldw + f] = l;
        }
    }
}

// Glo = G - trunc_tf32(G) for the KP x KP Gram matrices; grid = (ceil(KP*KP/256), B)
__global__ void __launch_bounds__(kThreads)
g_split_kernel(const float* __restrict__ G, float* __restrict__ Glo, int n, const ClipState* __restrict__ st) {
    const int b = blockIdx.y;
    if (st[b].done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float h, l; split_tf32(G[(long long)b * n + i], h, l); Glo[(long long)b * n + i] = l; }
}

// out[b][e] = sum_s partial[b][s][e]  (fixed order); grid = (ceil(n4/256), B)
__global__ void __launch_bounds__(kThreads)
reduce_splits_kernel(const float* __restrict__ partial, int S, long long n4, float* __restrict__ out, long long out_stride,
                     const ClipState* __restrict__ st) {
    const int b = blockIdx.y;
    if (st[b].done) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(partial + (((long long)b * S + s) * n4 + i) * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (long long)b * out_stride + i * 4) = acc;
}

// ---- host ----------------------------------------------------------------------------------------------------
static_assert(sizeof(CUtensorMap) == sizeof(TcMapBlob), "CUtensorMap is 128 bytes");
static inline const CUtensorMap& as_map(const TcMapBlob& b) { return *reinterpret_cast<const CUtensorMap*>(&b); }
static inline CUtensorMap* as_map_ptr(TcMapBlob* b) { return reinterpret_cast<CUtensorMap*>(b); }
template <int KP>
static cudaError_t tc_half1_impl(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    using Cfg = TcCfg<KP, XhtStages<KP>::value>;
    cudaError_t e = cudaFuncSetAttribute(xht_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int S = wk.tc_splits;
    AINMF_LAUNCH(xht_tc_kernel<KP>, dim3(wk.tc_mtiles, S, p.B), dim3(kThreads), Cfg::SMEM_BYTES, s, as_map(wk.tc->mapXmn),
                 as_map(wk.tc->mapHmn), p.F, p.T, wk.tc_fps, wk.xht_partial, wk.gram_partial, p.state);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const long long n4 = (long long)KP * KP / 4;
    AINMF_LAUNCH(reduce_splits_kernel, dim3((unsigned)ceil_div64(n4, kThreads), p.B), dim3(kThreads), 0, s, wk.gram_partial, S,
                 n4, wk.HHt, (long long)KP * KP, p.state);
    return cudaGetLastError();
}
cudaError_t nmf_tc_half1(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    return p.KP == 64 ? tc_half1_impl<64>(p, wk, s) : tc_half1_impl<128>(p, wk, s);
}

template <int KP>
static cudaError_t tc_hstep_impl(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    using Cfg = TcCfg<KP, HStepStages<KP>::value>;
    const int ldw = p.ldf;
    AINMF_LAUNCH(wt_split_kernel, dim3(ceil_div(ldw, 32), KP / 32, p.B), dim3(kThreads), 0, s, p.W, p.w_stride, p.F, KP, ldw,
                 wk.tc_Wt, wk.tc_WtLo, (long long)KP * ldw, p.state);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    static int ts_mode = -1;       // AINMF_TC_MODE=ss selects the shared-memory-operand kernel below; default: nmf_ts.cu
    if (ts_mode < 0) { const char* m_ = getenv("AINMF_TC_MODE"); ts_mode = (m_ && m_[0] == 's') ? 0 : 1; }
    if (ts_mode) return nmf_ts_hstep(p, wk, s);
    AINMF_LAUNCH(g_split_kernel, dim3(ceil_div(KP * KP, kThreads), p.B), dim3(kThreads), 0, s, wk.WtW, wk.tc_GLo, KP * KP, p.state);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    static int blk_mode = -1;
    if (blk_mode < 0) { const char* m_ = getenv("AINMF_TC_SWEEP"); blk_mode = (m_ && m_[0] == 's') ? 0 : 1; }   // "shfl" selects the shuffle sweep
    auto kern = blk_mode ? h_step_tc_kernel<KP, true> : h_step_tc_kernel<KP, false>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)) != cudaSuccess) return e;
    static long long* dbg = nullptr;
    static int dbg_left = -1;
    if (dbg_left < 0) {
        const char* e_ = getenv("AINMF_TC_DEBUG");
        dbg_left = (e_ && e_[0] == '1') ? 3 : 0;
        if (dbg_left) { cudaMalloc((void**)&dbg, 4096 * sizeof(long long)); cudaMemset(dbg, 0, 4096 * sizeof(long long)); }
    }
    AINMF_LAUNCH(kern, dim3(ceil_div(p.T, TC_M), p.B), dim3(kThreads), Cfg::SMEM_BYTES, s, as_map(wk.tc->mapX),
                 as_map(wk.tc->mapWt), as_map(wk.tc->mapWtLo), as_map(wk.tc->mapHk), as_map(wk.tc->mapG), as_map(wk.tc->mapGlo), p.F, p.T, wk.WtW,
                 p.Ht, p.h_stride, wk.violH, p.state,
                 dbg_left > 0 ? dbg : nullptr);
    if (dbg_left > 0) {
        --dbg_left;
        long long hbuf[8 + 6 * 40];
        cudaStreamSynchronize(s);
        cudaMemcpy(hbuf, dbg, sizeof hbuf, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tc-debug h_step KP=%d] roles done %lld, accum %lld, epilogue %lld, end %lld cycles\n", KP, hbuf[0], hbuf[1], hbuf[2], hbuf[3]);
        const int nk = (p.F + 31) / 32 + KP / 32;
        for (int i = 0; i < nk && i < 40; ++i)
            fprintf(stderr, "  stage %2d: slot free %7lld  tma issued %7lld  full seen %7lld  converted(t0) %7lld  conv barrier %7lld  mma committed %7lld\n", i,
                    hbuf[8 + 6 * i + 5], hbuf[8 + 6 * i], hbuf[8 + 6 * i + 1], hbuf[8 + 6 * i + 2], hbuf[8 + 6 * i + 4], hbuf[8 + 6 * i + 3]);
    }
    return cudaGetLastError();
}
cudaError_t nmf_tc_hstep(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    return p.KP == 64 ? tc_hstep_impl<64>(p, wk, s) : tc_hstep_impl<128>(p, wk, s);
}

int nmf_tc_setup(const NmfProblem& p, NmfWork* wk, TcMaps* m) {
    if (!wk->use_tc) return 0;
    const uint64_t B = p.B, T = p.T, F = p.F, ldf = p.ldf, KP = p.KP;
    int rc = make_tensor_map_3d(as_map_ptr(&m->mapX), p.Xt, F, T, B, ldf, (uint64_t)p.x_stride, TC_BK, TC_M, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapWt), wk->tc_Wt, F, KP, B, ldf, KP * ldf, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapWtLo), wk->tc_WtLo, F, KP, B, ldf, KP * ldf, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapXmn), p.Xt, F, T, B, ldf, (uint64_t)p.x_stride, 32, TC_BK, 1);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapHmn), p.Ht, KP, T, B, KP, (uint64_t)p.h_stride, 32, TC_BK, 1);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapHk), p.Ht, KP, T, B, KP, (uint64_t)p.h_stride, TC_BK, TC_M, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapG), wk->WtW, KP, KP, B, KP, KP * KP, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapGlo), wk->tc_GLo, KP, KP, B, KP, KP * KP, TC_BK, (uint32_t)KP, 0);
    wk->tc = m;
    return rc;
}

}  // namespace ainmf
#endif  // AINMF_EMU
