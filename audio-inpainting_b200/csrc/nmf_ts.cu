// nmf_ts.cu -- H half-step of the CD-NMF iteration as ONE persistent, warp-specialised sm_100a kernel in which the
// spectrogram operand never goes back to shared memory:
//
//   TMA (X chunk, W^T chunk: tf32 tile + bf16 cross tile) -> shared -> converter warps read their frame's row once,
//   build the two A operands in registers and store them with tcgen05.st into TMEM -> tcgen05.mma with the A operand
//   taken FROM TMEM (B = W^T chunk from shared) accumulates  D = X_tile.W - Ht_tile.(W^T W) = -(gradient)  in TMEM.
//
// Error compensation (a single TF32 pass breaks the reference's tolerances): with v_hi = v truncated to
// tf32 and v_lo = v - v_hi,   a*b ~= a_hi*b_hi  +  (a_lo*b_hi + a_hi*b_lo).   A tcgen05.mma costs ~128 cycles whatever
// its N <= 256 (measured, tests/mma_bench.py), so the instruction count is what matters: the main term is one
// kind::tf32 instruction per 8 contraction elements, and BOTH cross terms are one kind::f16 (bf16, K = 16)
// instruction whose K dimension is the concatenation [a_lo | a_hi].[b_hi ; b_lo]; the cross terms are 2^-11 of the
// result, so bf16 factors keep the sum accurate to ~2^-20 (tests/test_gpu_tc.py).
//
// Compared with the shared-memory-operand kernel (nmf_tc.cu) this removes the lo-tile write and the three operand
// reads of the X tile from shared memory (136 KB -> 72 KB of shared-memory traffic per 16 KB of X at K = 64), which
// is what bounded that kernel (profiles/r01_summary.md).  The accumulator is double-buffered, so the coordinate
// sweep of tile i (sweep warps + their own MMA issuer) overlaps the contraction of tile i+1.
//
// Sweep = blocked Gauss-Seidel, identical in exact arithmetic to the reference's sequential sweep
// ($SP/sklearn/decomposition/_cdnmf_fast.pyx:8-38): coordinates go in blocks of 8; inside a block one thread per frame
// updates sequentially (corrections from the block's own deltas), and the block's effect on all later coordinates,
// D -= delta[128x8].G[8xKP], is an error-compensated K = 8 tensor-core MMA whose A operand (delta) the sweep threads
// store straight into TMEM.
//
// Warp roles (K = 64: 512 threads; K = 128: one converter group, 384 threads; one CTA per SM, grid = min(#SM, #tiles), tiles round-robin):
//   warp 0: TMA producer        warp 1: MMA issuer of the contraction (+ TMEM owner)
//   warp 2: MMA issuer of the sweep updates + loader of the per-block Gram operands      warp 3: idle
//   warps 4-7 and 8-11: two converter groups that take alternate chunks (thread = frame = TMEM lane): a converter's
//       chunk costs two mbarrier waits (>= 90 cycles each, even when already complete) + the split + tcgen05.st, ~800
//       cycles, against the ~670 cycles the chunk's 16 KB of X take at full HBM rate -- one group could not keep up
//   warps 12-15: sweep (thread = frame = TMEM lane)
#include "kernels.h"
#include <stdlib.h>

#include "tc.cuh"
#include "nmf_ts.cuh"

#ifndef AINMF_EMU
namespace ainmf {
using namespace tc;

constexpr int TS_BK = 32;            // contraction elements per stage (one 128-byte row)
constexpr int TS_M = 128;            // frames per tile = MMA M = TMEM lanes
// converter groups (4 warps each, alternate chunks): two at K = 64; at K = 128 the second group's gain in the contraction is
// lost again in the sweep and the epilogue, which get 128 instead of 168 registers per thread in a 512-thread CTA (measured)
template <int KP> struct TsShape {
    static constexpr int CONV_GROUPS = (KP == 64) ? 2 : 1;
    static constexpr int SWEEP_WARP0 = 4 + 4 * CONV_GROUPS;      // first sweep / epilogue warp
    static constexpr int THREADS = 32 * (SWEEP_WARP0 + 4);
};

template <int KP> struct TsCfg {
    static constexpr int NSS = (KP == 64) ? 6 : 4;               // shared-memory stages (X chunk + W^T hi/lo chunk)
    static constexpr int NAS = (KP == 64) ? 5 : 3;               // TMEM stages of the converted X chunk (hi | lo)
    static constexpr int X_BYTES = TS_M * TS_BK * 4;             // 16 KB
    static constexpr int B_BYTES = KP * TS_BK * 4;               // 8 / 16 KB
    static constexpr int STAGE_BYTES = X_BYTES + 2 * B_BYTES;
    static constexpr int NBLK = KP / 8;
    static constexpr int NG = (KP == 64) ? 4 : 2;                // ring of per-block update operands
    static constexpr int BLOB_FLOATS = 16 * KP;                  // G rows of the block as a K-major [KP][8] operand: hi, lo
    static constexpr int BLOB_BYTES = BLOB_FLOATS * 4;
    static constexpr int SC_BYTES = NBLK * TS_SC * 4;            // sweep scalars of one clip
    static constexpr int SMEM_BYTES = NSS * STAGE_BYTES + NG * BLOB_BYTES + 2 * SC_BYTES + 1024;
    // TMEM columns: two accumulators, two slots of the sweep's delta operand (hi 8 | lo 8), the A stages (hi 32 | lo 32)
    static constexpr int COL_D = 0, COL_DELTA = 2 * KP, COL_A = 2 * KP + 32;
    static_assert(COL_A + NAS * 64 <= 512, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct TsBarriers {
    uint64_t full[6], empty[6];      // shared-memory stage: TMA landed / MMA finished reading W^T chunk
    uint64_t conv[5], aempty[5];     // TMEM A stage: converters done / MMA finished reading it
    uint64_t dfull[2], dempty[2];    // accumulator: contraction complete / sweep has read its last block
    uint64_t dready[2], ddone[2];    // sweep, per delta slot (block & 1): delta block stored in TMEM / rank-8 update complete
    uint64_t gfull[4], gempty[4];    // update operand ring: blob landed / update MMA finished reading it
    uint64_t sfull[2], sempty[2];    // sweep scalars of the tile's clip: landed / tile swept
};

// Tiles of clips whose stop rule fired are skipped; every role walks the same sequence.
struct TsTiles {
    int tile, step, n, nH;
    const ClipState* st;
    __device__ __forceinline__ bool next(int& b, int& mt) {
        while (tile < n) {
            b = tile / nH;
            mt = tile - b * nH;
            tile += step;
            if (!st[b].done) return true;
        }
        return false;
    }
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sweep_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int KP>
__global__ void __launch_bounds__(TsShape<KP>::THREADS, 1)
h_step_ts_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWt,
                 const __grid_constant__ CUtensorMap mapWtLo, const __grid_constant__ CUtensorMap mapHk,
                 const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapGlo, int F, int T, int B,
                 const float* __restrict__ blobs /*[B][NBLK][16 KP]*/, const float* __restrict__ scal /*[B][NBLK][TS_SC]*/,
                 float* __restrict__ Ht, long long h_stride,
                 float* __restrict__ viol /*[B][nH]*/, const ClipState* __restrict__ st, long long* __restrict__ dbg, int exp_flags,
                 const int* __restrict__ t_good /*good-first frame order: frames >= t_good[b] are the identical bad ones; or null*/,
                 const float* __restrict__ vfill /*[B][KP]: fill^T.W*/) {
    using Cfg = TsCfg<KP>;
    constexpr int NSS = Cfg::NSS, NAS = Cfg::NAS, NBLK = Cfg::NBLK, NG = Cfg::NG;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) TsBarriers bars;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_v[4];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sblob = smem + (size_t)NSS * Cfg::STAGE_BYTES;
    unsigned char* sscal = sblob + (size_t)NG * Cfg::BLOB_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nH = (T + TS_M - 1) / TS_M;
    const int nkX = (F + TS_BK - 1) / TS_BK;
    const int nk = nkX + KP / TS_BK;
    const bool dbg_on = dbg != nullptr && blockIdx.x == 0;
    const long long dbg_t0 = clock64();
    // a tile that holds only bad frames needs no X chunks: its X^T.W rows are all fill^T.W, added by the sweep threads
    auto first_chunk = [&](int bb, int mm) { return (t_good && mm * TS_M >= t_good[bb]) ? nkX : 0; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSS; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int a = 0; a < NAS; ++a) { mbar_init(&bars.conv[a], 4); mbar_init(&bars.aempty[a], 1); }
        for (int j = 0; j < 2; ++j) { mbar_init(&bars.dfull[j], 1); mbar_init(&bars.dempty[j], 4); }
        for (int j = 0; j < 2; ++j) { mbar_init(&bars.dready[j], 4); mbar_init(&bars.ddone[j], 1); }
        for (int j = 0; j < NG; ++j) { mbar_init(&bars.gfull[j], 1); mbar_init(&bars.gempty[j], 1); }
        for (int j = 0; j < 2; ++j) { mbar_init(&bars.sfull[j], 1); mbar_init(&bars.sempty[j], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWt); tma_prefetch_desc(&mapWtLo);
        tma_prefetch_desc(&mapHk); tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapGlo);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    TsTiles tiles{(int)blockIdx.x, (int)gridDim.x, B * nH, nH, st};
    int b, mt;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            uint32_t it = 0;
            while (tiles.next(b, mt)) {
                const int m0 = mt * TS_M;
                for (int i = first_chunk(b, mt); i < nk; ++i, ++it) {
                    const uint32_t s = it % NSS, ph = (it / NSS) & 1;
                    mbar_wait(&bars.empty[s], ph ^ 1);
                    unsigned char* stg = smem + (size_t)s * Cfg::STAGE_BYTES;
                    if (exp_flags && i < nkX) {       // timing experiments only (wrong results): 1 = no Wlo, 2 = no W, 4 = no X
                        const uint32_t bytes = ((exp_flags & 4) ? 0 : Cfg::X_BYTES) + ((exp_flags & 2) ? 0 : Cfg::B_BYTES) + ((exp_flags & 3) ? 0 : Cfg::B_BYTES);
                        if (bytes) mbar_arrive_expect_tx(&bars.full[s], bytes); else mbar_arrive(&bars.full[s]);
                        if (!(exp_flags & 4)) tma_load_3d(stg, &mapX, &bars.full[s], i * TS_BK, m0, b);
                        if (!(exp_flags & 2)) tma_load_3d(stg + Cfg::X_BYTES, &mapWt, &bars.full[s], i * TS_BK, 0, b);
                        if (!(exp_flags & 3)) tma_load_3d(stg + Cfg::X_BYTES + Cfg::B_BYTES, &mapWtLo, &bars.full[s], i * TS_BK, 0, b);
                        continue;
                    }
                    mbar_arrive_expect_tx(&bars.full[s], Cfg::X_BYTES + 2 * Cfg::B_BYTES);
                    if (i < nkX) {
                        tma_load_3d(stg, &mapX, &bars.full[s], i * TS_BK, m0, b);
                        tma_load_3d(stg + Cfg::X_BYTES, &mapWt, &bars.full[s], i * TS_BK, 0, b);
                        tma_load_3d(stg + Cfg::X_BYTES + Cfg::B_BYTES, &mapWtLo, &bars.full[s], i * TS_BK, 0, b);
                    } else {
                        tma_load_3d(stg, &mapHk, &bars.full[s], (i - nkX) * TS_BK, m0, b);
                        tma_load_3d(stg + Cfg::X_BYTES, &mapG, &bars.full[s], (i - nkX) * TS_BK, 0, b);
                        tma_load_3d(stg + Cfg::X_BYTES + Cfg::B_BYTES, &mapGlo, &bars.full[s], (i - nkX) * TS_BK, 0, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: D[buf] = X_tile.W - Ht_tile.G ----------------
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(TS_M, KP, 0, 0), idesc16 = make_idesc_bf16(TS_M, KP, 0, 0);
            const uint32_t neg = 1u << 13;                            // negate A
            uint32_t it = 0, tl = 0;
            long long q_ae = 0, q_full = 0, q_conv = 0, q_iss = 0, q_t = 0;
            while (tiles.next(b, mt)) {
                const uint32_t buf = tl & 1;
                mbar_wait(&bars.dempty[buf], ((tl >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                if (dbg_on && tl < 64) dbg[8 * tl + 0] = clock64() - dbg_t0;
                const uint32_t dcol = tmem + Cfg::COL_D + buf * KP;
                const int i0 = first_chunk(b, mt);
                for (int i = i0; i < nk; ++i, ++it) {
                    const uint32_t s = it % NSS, a = it % NAS;
                    // conv[a] first: the converters waited for full[s] themselves, so the second wait finds its phase
                    // complete.  At most NAS chunks are in the tensor pipe (an A stage is recycled when its MMAs finish).
                    if (dbg_on) q_t = clock64();
                    mbar_wait(&bars.conv[a], (it / NAS) & 1);
                    if (dbg_on) { const long long c = clock64(); q_conv += c - q_t; q_t = c; }
                    mbar_wait(&bars.full[s], (it / NSS) & 1);
                    if (dbg_on) { const long long c = clock64(); q_full += c - q_t; q_t = c; }
                    tcgen05_fence_after();
                    const uint32_t ng = (i < nkX) ? 0u : neg;
                    const uint64_t d_bh = make_smem_desc(smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES + Cfg::X_BYTES), 16, 1024);
                    const uint64_t d_bl = d_bh + (uint64_t)(Cfg::B_BYTES >> 4);
                    const uint32_t acol = tmem + Cfg::COL_A + a * 64;
                    // (the last X chunk may hold fewer than 32 bins: the TMA zero-fills the rest of both operands, and four
                    //  straight-line instruction pairs issue faster than a loop with a data-dependent exit)
#pragma unroll
                    for (int k8 = 0; k8 < TS_BK / 8; ++k8) {
                        const uint64_t o = (uint64_t)(k8 * 32 >> 4);
                        mma_tf32_ts(dcol, acol + k8 * 8, d_bh + o, idesc | ng, (i > i0 || k8 > 0) ? 1u : 0u);
                        mma_bf16_ts(dcol, acol + 32 + k8 * 8, d_bl + o, idesc16 | ng, 1);
                    }
                    mma_commit(&bars.empty[s]);
                    mma_commit(&bars.aempty[a]);
                    if (dbg_on) q_iss += clock64() - q_t;
                }
                mma_commit(&bars.dfull[buf]);
                if (dbg_on && tl < 64) dbg[8 * tl + 1] = clock64() - dbg_t0;
                ++tl;
            }
            if (dbg_on) { dbg[8 * 65 + 0] = q_ae; dbg[8 * 65 + 1] = q_full; dbg[8 * 65 + 2] = q_conv; dbg[8 * 65 + 3] = q_iss; dbg[8 * 65 + 4] = it; }
        }
    } else if (warp == 2) {
        // ---------------- sweep: loader of scalars / update operands + MMA issuer of D -= delta.G ----------------
        // Block blk of a tile needs an update MMA only if blk <= NBLK-3: the sweep threads themselves carry a block's
        // deltas into the NEXT block's 8 coordinates (look-ahead), so an update has one whole block of slack.
        if (elect_one()) {
            const uint32_t idn = make_idesc_tf32(TS_M, KP, 0, 0) | (1u << 13), idn16 = make_idesc_bf16(TS_M, KP, 0, 0) | (1u << 13);
            TsTiles cur{(int)blockIdx.x, (int)gridDim.x, B * nH, nH, st};     // operand load cursor
            int cb = 0, cmt = 0, cblk = 0;
            bool cmore = cur.next(cb, cmt);
            TsTiles scur{(int)blockIdx.x, (int)gridDim.x, B * nH, nH, st};    // scalar load cursor (one tile ahead)
            uint32_t sloaded = 0;
            uint32_t loaded = 0, g = 0, tl = 0;
            auto refill = [&]() {
                while (cmore && loaded < g + NG) {
                    const uint32_t slot = loaded % NG;
                    mbar_wait(&bars.gempty[slot], ((loaded / NG) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars.gfull[slot], Cfg::BLOB_BYTES);
                    bulk_load_1d(sblob + (size_t)slot * Cfg::BLOB_BYTES, blobs + ((size_t)cb * NBLK + cblk) * Cfg::BLOB_FLOATS,
                                 Cfg::BLOB_BYTES, &bars.gfull[slot]);
                    ++loaded;
                    if (++cblk == NBLK - 2) { cblk = 0; cmore = cur.next(cb, cmt); }
                }
            };
            auto load_scalars = [&]() {
                int sb, smt;
                if (!scur.next(sb, smt)) return;
                const uint32_t j = sloaded & 1;
                mbar_wait(&bars.sempty[j], ((sloaded >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars.sfull[j], Cfg::SC_BYTES);
                bulk_load_1d(sscal + (size_t)j * Cfg::SC_BYTES, scal + (size_t)sb * NBLK * TS_SC, Cfg::SC_BYTES, &bars.sfull[j]);
                ++sloaded;
            };
            load_scalars();
            refill();
            while (tiles.next(b, mt)) {
                load_scalars();                                       // next tile's
                const uint32_t dcol = tmem + Cfg::COL_D + (tl & 1) * KP;
                for (int blk = 0; blk + 2 < NBLK; ++blk, ++g) {
                    refill();
                    const uint32_t slot = g % NG;
                    mbar_wait(&bars.gfull[slot], (g / NG) & 1);
                    mbar_wait(&bars.dready[g & 1], (g >> 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t bh = smem_u32(sblob + (size_t)slot * Cfg::BLOB_BYTES);
                    // K-major, no swizzle: 8-row core matrices 128 B apart along N (SBO), K-adjacent cores KP*16 B apart (LBO)
                    const uint64_t d_gh = make_smem_desc(bh, KP * 16, 128, 0);
                    const uint64_t d_gl = make_smem_desc(bh + 8 * KP * 4, KP * 16, 128, 0);
                    mma_tf32_ts(dcol, tmem + Cfg::COL_DELTA + 16 * (g & 1), d_gh, idn, 1);
                    mma_bf16_ts(dcol, tmem + Cfg::COL_DELTA + 16 * (g & 1) + 8, d_gl, idn16, 1);
                    mma_commit(&bars.ddone[g & 1]);
                    mma_commit(&bars.gempty[slot]);
                }
                ++tl;
            }
        }
    } else if (warp >= 4 && warp < TsShape<KP>::SWEEP_WARP0) {
        // ---------------- converters: shared X chunk -> (hi, lo) -> TMEM A stage; group 0 takes the even chunks ----------------
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t grp = (warp - 4) >> 2;
        constexpr uint32_t gmask = TsShape<KP>::CONV_GROUPS - 1;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + Cfg::COL_A;
        const int sw = row & 7;
        uint32_t it = 0;
        long long k_full = 0, k_ae = 0, k_rd = 0, k_st = 0, k_t = 0;
        const bool kd = dbg_on && row == 0;
        while (tiles.next(b, mt)) {
            for (int i = first_chunk(b, mt); i < nk; ++i, ++it) {
                if ((it & gmask) != grp) continue;
                const uint32_t s = it % NSS, a = it % NAS;
                if (kd) k_t = clock64();
                if (lane == 0) mbar_wait(&bars.full[s], (it / NSS) & 1);
                if (kd) { const long long c = clock64(); k_full += c - k_t; k_t = c; }
                if (lane == 0) mbar_wait(&bars.aempty[a], ((it / NAS) & 1) ^ 1);
                __syncwarp();
                tcgen05_fence_after();
                if (kd) { const long long c = clock64(); k_ae += c - k_t; k_t = c; }
                const unsigned char* xr = smem + (size_t)s * Cfg::STAGE_BYTES + row * 128;
                float hi[32], lo[32];                                 // lo: the bf16 cross operand, [lo(8) | hi(8)] per 8 elements
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(xr + ((j ^ sw) << 4));
                    hi[4 * j] = v.x; hi[4 * j + 1] = v.y; hi[4 * j + 2] = v.z; hi[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) cross_pack8(hi + 8 * j, lo + 8 * j, true);
                if (kd) { const long long c = clock64(); k_rd += c - k_t; k_t = c; }
                tmem_st_32x32(tlane + a * 64, hi);                    // raw bits: the tensor core ignores the low 13
                tmem_st_32x32(tlane + a * 64 + 32, lo);
                tmem_wait_st();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.conv[a]);
                if (kd) { const long long c = clock64(); k_st += c - k_t; k_t = c; }
            }
        }
        if (kd) { dbg[8 * 64 + 0] = k_full; dbg[8 * 64 + 1] = k_ae; dbg[8 * 64 + 2] = k_rd; dbg[8 * 64 + 3] = k_st; dbg[8 * 64 + 4] = it; }
    } else if (warp >= TsShape<KP>::SWEEP_WARP0) {
        // ---------------- sweep: thread = frame ----------------
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t tl = 0, ug = 0, wc = 0;      // tiles, update requests, update waits
        // The old values of this thread's Ht row are fetched PD blocks ahead of their use, across tile boundaries: under
        // the contraction's TMA stream a global load takes thousands of cycles, and the sweep is a serial chain.
        constexpr int PD = 4;
        float4 r0[PD], r1[PD];
        TsTiles pre{(int)blockIdx.x, (int)gridDim.x, B * nH, nH, st};
        int pblk = 0;
        const float* prow = nullptr;
        {
            int pb, pmt;
            if (pre.next(pb, pmt)) { const int pt = pmt * TS_M + row; prow = (pt < T) ? Ht + (long long)pb * h_stride + (long long)pt * KP : nullptr; }
        }
        auto prefetch = [&](float4& o0, float4& o1) {            // next (tile, block) of this thread's stream
            o0 = make_float4(0.f, 0.f, 0.f, 0.f); o1 = o0;
            if (prow) { o0 = *reinterpret_cast<const float4*>(prow + 8 * pblk); o1 = *reinterpret_cast<const float4*>(prow + 8 * pblk + 4); }
            if (++pblk == NBLK) {
                pblk = 0;
                prow = nullptr;
                int pb, pmt;
                if (pre.next(pb, pmt)) { const int pt = pmt * TS_M + row; prow = (pt < T) ? Ht + (long long)pb * h_stride + (long long)pt * KP : nullptr; }
                else pre.tile = pre.n;
            }
        };
#pragma unroll
        for (int i = 0; i < PD; ++i) prefetch(r0[i], r1[i]);
        while (tiles.next(b, mt)) {
            const uint32_t buf = tl & 1;
            const int t = mt * TS_M + row;
            const bool valid = t < T;
            const bool badrow = t_good && valid && t >= t_good[b];           // its X row is zero in the permuted copy: X^T.W row = fill^T.W
            float* hrow = Ht + (long long)b * h_stride + (long long)t * KP;
            if (lane == 0) {
                mbar_wait(&bars.sfull[buf], (tl >> 1) & 1);
                mbar_wait(&bars.dfull[buf], (tl >> 1) & 1);
            }
            __syncwarp();
            tcgen05_fence_after();
            if (dbg_on && row == 0 && tl < 64) dbg[8 * tl + 2] = clock64() - dbg_t0;
            const float* sc = reinterpret_cast<const float*>(sscal + (size_t)buf * Cfg::SC_BYTES);
            float vsum = 0.f;
            float d8[8];
            long long c_ld = 0, c_st = 0, c_dd = 0, c_t;
#define TS_TIC() do { if (dbg_on) c_t = clock64(); } while (0)
#define TS_TOC(acc) do { if (dbg_on) acc += clock64() - c_t; } while (0)
            TS_TIC();
            tmem_ld_32x8(tlane + Cfg::COL_D + buf * KP, d8);
            const float* vrow = vfill + (long long)b * KP;
            if (badrow) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(vrow)), v1 = __ldg(reinterpret_cast<const float4*>(vrow + 4));
                d8[0] += v0.x; d8[1] += v0.y; d8[2] += v0.z; d8[3] += v0.w; d8[4] += v1.x; d8[5] += v1.y; d8[6] += v1.z; d8[7] += v1.w;
            }
            TS_TOC(c_ld);
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                const float* gb = sc + blk * TS_SC;
                float a8[8], dl[8];
                float4 vn0 = make_float4(0.f, 0.f, 0.f, 0.f), vn1 = vn0;     // fill^T.W of the next block, fetched a block ahead
                if (badrow && blk + 1 < NBLK) {
                    vn0 = __ldg(reinterpret_cast<const float4*>(vrow + 8 * (blk + 1)));
                    vn1 = __ldg(reinterpret_cast<const float4*>(vrow + 8 * (blk + 1) + 4));
                }
                const bool fd = dbg_on && row == 0 && tl == 2 && blk < 16;
#define TS_STAMP(k) do { if (fd) dbg[8 * 66 + blk * 8 + (k)] = clock64() - dbg_t0; } while (0)
                TS_STAMP(0);
                a8[0] = r0[0].x; a8[1] = r0[0].y; a8[2] = r0[0].z; a8[3] = r0[0].w;
                a8[4] = r1[0].x; a8[5] = r1[0].y; a8[6] = r1[0].z; a8[7] = r1[0].w;
#pragma unroll
                for (int i = 0; i + 1 < PD; ++i) { r0[i] = r0[i + 1]; r1[i] = r1[i + 1]; }
                prefetch(r0[PD - 1], r1[PD - 1]);
                const float4 i0 = *reinterpret_cast<const float4*>(gb + 128);
                const float4 i1 = *reinterpret_cast<const float4*>(gb + 132);
                const float inv8[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // gradient at the current point: -(accumulator) + sum_{i<j} G[8blk+j][8blk+i] * delta_i; the terms
                    // with known deltas are summed off the critical path, the newest delta enters last
                    float part = -d8[j];
                    if (j > 1) {
                        const float4 g0 = *reinterpret_cast<const float4*>(gb + 8 * j);
                        part = fmaf(g0.x, dl[0], part);
                        if (j > 2) part = fmaf(g0.y, dl[1], part);
                        if (j > 3) part = fmaf(g0.z, dl[2], part);
                        if (j > 4) part = fmaf(g0.w, dl[3], part);
                    }
                    if (j > 5) {
                        const float4 g1 = *reinterpret_cast<const float4*>(gb + 8 * j + 4);
                        part = fmaf(g1.x, dl[4], part);
                        if (j > 6) part = fmaf(g1.y, dl[5], part);
                    }
                    const float grad = (j > 0) ? fmaf(gb[8 * j + j - 1], dl[j - 1], part) : part;
                    const float inv = inv8[j];
                    const float aq = a8[j];
                    const float pg = (aq == 0.f) ? fminf(0.f, grad) : grad;
                    const float an = fmaxf(fmaf(-grad, inv, aq), 0.f);
                    const bool upd = valid && (inv != 0.f);
                    vsum += valid ? fabsf(pg) : 0.f;
                    dl[j] = upd ? an - aq : 0.f;
                    a8[j] = upd ? an : aq;
                }
                TS_STAMP(1);
                if (valid) {
                    *reinterpret_cast<float4*>(hrow + 8 * blk) = make_float4(a8[0], a8[1], a8[2], a8[3]);
                    *reinterpret_cast<float4*>(hrow + 8 * blk + 4) = make_float4(a8[4], a8[5], a8[6], a8[7]);
                }
                TS_STAMP(2);
                if (blk + 1 < NBLK) {
                    // the next block's accumulator columns: every update up to block blk-1 must have landed ...
                    if (blk >= 1) {
                        TS_TIC();
                        if (lane == 0) mbar_wait(&bars.ddone[wc & 1], (wc >> 1) & 1);
                        ++wc;
                        __syncwarp();
                        tcgen05_fence_after();
                        TS_TOC(c_dd);
                    }
                    TS_STAMP(3);
                    TS_TIC();
                    tmem_ld_32x8(tlane + Cfg::COL_D + buf * KP + (blk + 1) * 8, d8);
                    if (badrow) {
                        d8[0] += vn0.x; d8[1] += vn0.y; d8[2] += vn0.z; d8[3] += vn0.w;
                        d8[4] += vn1.x; d8[5] += vn1.y; d8[6] += vn1.z; d8[7] += vn1.w;
                    }
                    TS_TOC(c_ld);
                    TS_STAMP(4);
                    if (blk + 2 == NBLK) {                            // last read of this accumulator, no update pending
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&bars.dempty[buf]); }
                    } else {                                          // ... and this block's deltas go to the tensor core
                        float hl[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hl[j] = dl[j];
                        cross_pack8(dl, hl + 8, true);
                        TS_TIC();
                        tmem_st_32x16(tlane + Cfg::COL_DELTA + 16 * (ug & 1), hl);
                        tmem_wait_st();
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.dready[ug & 1]);
                        ++ug;
                        TS_TOC(c_st);
                    }
                    TS_STAMP(5);
                    // ... while this thread applies them to the next 8 coordinates itself: D -= delta . G
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 c0 = *reinterpret_cast<const float4*>(gb + 64 + 8 * i);
                        const float4 c1 = *reinterpret_cast<const float4*>(gb + 64 + 8 * i + 4);
                        d8[0] = fmaf(-dl[i], c0.x, d8[0]); d8[1] = fmaf(-dl[i], c0.y, d8[1]);
                        d8[2] = fmaf(-dl[i], c0.z, d8[2]); d8[3] = fmaf(-dl[i], c0.w, d8[3]);
                        d8[4] = fmaf(-dl[i], c1.x, d8[4]); d8[5] = fmaf(-dl[i], c1.y, d8[5]);
                        d8[6] = fmaf(-dl[i], c1.z, d8[6]); d8[7] = fmaf(-dl[i], c1.w, d8[7]);
                    }
                    TS_STAMP(6);
                }
            }
            vsum = warp_sum(vsum);
            if (lane == 0) { s_v[q] = vsum; mbar_arrive(&bars.sempty[buf]); }
            sweep_bar();
            if (row == 0) {
                viol[(long long)b * nH + mt] = ((s_v[0] + s_v[1]) + s_v[2]) + s_v[3];
                if (dbg_on && tl < 64) { dbg[8 * tl + 3] = clock64() - dbg_t0; dbg[8 * tl + 4] = c_ld; dbg[8 * tl + 5] = c_st; dbg[8 * tl + 6] = c_dd; }
            }
            sweep_bar();
            ++tl;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// =====================================================================================================
// W half-step, local part:  [X | Ht]^T Ht  -> X.Ht partials (F x KP) and the Gram Ht^T Ht (KP x KP), split over time.
// Persistent, same scheme as the H step: work items (clip, time split, 128-column tile of the virtual matrix
// [ X (ceil(F/32) slabs of 32 columns) | Ht (KP/32 slabs) ]), round-robin over the CTAs.  Per chunk of 32 frames:
//   TMA: 4 column slabs [32 t][32 cols] of the tile (dense), the Ht chunk as the MN-major tf32 B operand
//        (SWIZZLE_128B_ATOM_32B) and once more as dense slabs for the threads;
//   converters: thread = column of the tile = TMEM lane: 32 frames of its column -> tf32 A operand + bf16 cross operand
//        (tcgen05.st); thread n < KP also turns column n of the dense Ht chunk into row n of the K-major bf16 cross
//        operand of B in shared memory ([hi(8) | lo(8)] per 8 frames);
//   MMA:  per 8 frames one kind::tf32 (A from TMEM, B = raw Ht chunk) and one kind::f16 instruction (cross terms);
//   epilogue warps store finished accumulators (double-buffered) to the partial buffers.
// =====================================================================================================
// role layout of the X.Ht kernel: its converters also build the B cross operand, about twice the H step's work per chunk
// (three groups at K = 64 were measured slower than two: 72 registers per thread, X.Ht 0.55 instead of 0.52 ms)
template <int KP> struct XtShape {
    static constexpr int CONV_GROUPS = (KP == 64) ? 2 : 1;
    static constexpr int SWEEP_WARP0 = 4 + 4 * CONV_GROUPS;      // first epilogue warp
    static constexpr int THREADS = 32 * (SWEEP_WARP0 + 4);
};
template <int KP> struct XtCfg {
    static constexpr int NSS = (KP == 64) ? 5 : 3;               // shared-memory stages
    static constexpr int NAS = (KP == 64) ? 5 : 3;               // TMEM A stages
    static constexpr int SLAB = 32 * TS_BK * 4;                  // 4 KB: 32 frames x 32 columns
    static constexpr int A_BYTES = 4 * SLAB;                     // 16 KB
    static constexpr int B_BYTES = KP * TS_BK * 4;               // 8 / 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + 3 * B_BYTES;    // A slabs | B raw (MN-major) | B dense | B cross (K-major bf16)
    static constexpr int SMEM_BYTES = NSS * STAGE_BYTES + 1024;
    static constexpr int COL_D = 0, COL_A = 2 * KP;
    static_assert(COL_A + NAS * 64 <= 512, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct XtItems {
    int item, step, n, mtiles, S;
    const ClipState* st;
    __device__ __forceinline__ bool next(int& b, int& split, int& mt) {
        while (item < n) {
            mt = item % mtiles;
            const int r = item / mtiles;
            split = r % S;
            b = r / S;
            item += step;
            if (!st[b].done) return true;
        }
        return false;
    }
};

template <int KP>
__global__ void __launch_bounds__(XtShape<KP>::THREADS, 1)
xht_ts_kernel(const __grid_constant__ CUtensorMap mapXs, const __grid_constant__ CUtensorMap mapHs,
              const __grid_constant__ CUtensorMap mapHmn, int F, int T, int B, int S, int mtiles, int frames_per_split,
              float* __restrict__ xht_partial /*[B][S][F][KP]*/, float* __restrict__ gram_partial /*[B][S][KP][KP]*/,
              const ClipState* __restrict__ st, const int* __restrict__ t_good /*good-first frame order or null*/,
              long long* __restrict__ dbg /*AINMF_TC_DEBUG=1: wait / work cycles of CTA 0's roles, or null*/) {
    using Cfg = XtCfg<KP>;
    const bool dbg_on = dbg != nullptr && blockIdx.x == 0;
    constexpr int NSS = Cfg::NSS, NAS = Cfg::NAS, NB = KP / 32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) TsBarriers bars;
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nxs = (F + 31) / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSS; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int a = 0; a < NAS; ++a) { mbar_init(&bars.conv[a], 4); mbar_init(&bars.aempty[a], 1); }
        for (int j = 0; j < 2; ++j) { mbar_init(&bars.dfull[j], 1); mbar_init(&bars.dempty[j], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mapXs); tma_prefetch_desc(&mapHs); tma_prefetch_desc(&mapHmn);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    XtItems items{(int)blockIdx.x, (int)gridDim.x, B * S * mtiles, mtiles, S, st};
    int b, split, mt;
    // chunks of 32 frames of a work item; with the good-first frame order the X tiles stop at the last good frame (the
    // permuted copy holds zeros beyond; the bad frames' share comes back as fill (x) hbad in the W-side kernel), the tile(s)
    // that hold Ht columns (Gram) run over every frame
    auto chunks_of = [&](int bb, int sp, int mm) {
        const int t_begin = sp * frames_per_split;
        int t_end = min(T, t_begin + frames_per_split);
        if (t_good && 4 * mm + 3 < nxs) t_end = min(t_end, t_good[bb]);
        return max(0, (t_end - t_begin + TS_BK - 1) / TS_BK);
    };

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            long long p_w = 0, p_t = 0;
            while (items.next(b, split, mt)) {
                const int nk = chunks_of(b, split, mt), t_begin = split * frames_per_split;
                for (int i = 0; i < nk; ++i, ++it) {
                    const uint32_t s = it % NSS;
                    if (dbg_on) p_t = clock64();
                    mbar_wait(&bars.empty[s], ((it / NSS) & 1) ^ 1);
                    if (dbg_on) p_w += clock64() - p_t;
                    unsigned char* stg = smem + (size_t)s * Cfg::STAGE_BYTES;
                    const int t0 = t_begin + i * TS_BK;           // frames past T load zeros; split boundaries are multiples of 32
                    mbar_arrive_expect_tx(&bars.full[s], Cfg::A_BYTES + 2 * Cfg::B_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int v = 4 * mt + j;
                        if (v < nxs) tma_load_3d(stg + j * Cfg::SLAB, &mapXs, &bars.full[s], 32 * v, t0, b);
                        else tma_load_3d(stg + j * Cfg::SLAB, &mapHs, &bars.full[s], 32 * (v - nxs), t0, b);   // >= KP: zero fill
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        tma_load_3d(stg + Cfg::A_BYTES + j * Cfg::SLAB, &mapHmn, &bars.full[s], 32 * j, t0, b);
                        tma_load_3d(stg + Cfg::A_BYTES + Cfg::B_BYTES + j * Cfg::SLAB, &mapHs, &bars.full[s], 32 * j, t0, b);
                    }
                }
            }
            if (dbg_on) { dbg[0] = p_w; dbg[1] = it; }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(TS_M, KP, 0, 1), idesc16 = make_idesc_bf16(TS_M, KP, 0, 0);
            uint32_t it = 0, tl = 0;
            long long q_d = 0, q_c = 0, q_f = 0, q_i = 0, q_t = 0;
            while (items.next(b, split, mt)) {
                const int nk = chunks_of(b, split, mt);
                const uint32_t buf = tl & 1;
                if (dbg_on) q_t = clock64();
                mbar_wait(&bars.dempty[buf], ((tl >> 1) & 1) ^ 1);
                if (dbg_on) q_d += clock64() - q_t;
                tcgen05_fence_after();
                const uint32_t dcol = tmem + Cfg::COL_D + buf * KP;
                for (int i = 0; i < nk; ++i, ++it) {
                    const uint32_t s = it % NSS, a = it % NAS;
                    if (dbg_on) q_t = clock64();
                    mbar_wait(&bars.conv[a], (it / NAS) & 1);       // the converters waited for full[s] themselves
                    if (dbg_on) { const long long c = clock64(); q_c += c - q_t; q_t = c; }
                    mbar_wait(&bars.full[s], (it / NSS) & 1);
                    if (dbg_on) { const long long c = clock64(); q_f += c - q_t; q_t = c; }
                    tcgen05_fence_after();
                    const uint32_t b_raw = smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES);
                    const uint64_t d_bx = make_smem_desc(b_raw + 2 * Cfg::B_BYTES, 16, 1024);
                    const uint32_t acol = tmem + Cfg::COL_A + a * 64;
#pragma unroll
                    for (int k8 = 0; k8 < TS_BK / 8; ++k8) {
                        const uint64_t d_br = make_smem_desc(b_raw + k8 * 1024, Cfg::SLAB, 512, kLayoutSw128Base32);
                        mma_tf32_ts(dcol, acol + k8 * 8, d_br, idesc, (i > 0 || k8 > 0) ? 1u : 0u);
                        mma_bf16_ts(dcol, acol + 32 + k8 * 8, d_bx + (uint64_t)(k8 * 32 >> 4), idesc16, 1);
                    }
                    mma_commit(&bars.empty[s]);
                    mma_commit(&bars.aempty[a]);
                    if (dbg_on) q_i += clock64() - q_t;
                }
                mma_commit(&bars.dfull[buf]);
                ++tl;
            }
            if (dbg_on) { dbg[2] = q_d; dbg[3] = q_c; dbg[4] = q_f; dbg[5] = q_i; dbg[6] = it; dbg[7] = tl; }
        }
    } else if (warp >= 4 && warp < XtShape<KP>::SWEEP_WARP0) {
        // converter groups of 4 warps take the chunks round-robin, as in the H step
        const int q = warp & 3, col = q * 32 + lane;
        const uint32_t grp = (warp - 4) >> 2;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + Cfg::COL_A;
        uint32_t it = 0;
        long long k_w = 0, k_a = 0, k_b = 0, k_s = 0, k_n = 0, k_t = 0;
        const bool kd = dbg_on && col == 0;
        while (items.next(b, split, mt)) {
            const int nk = chunks_of(b, split, mt);
            for (int i = 0; i < nk; ++i, ++it) {
                if (it % (uint32_t)XtShape<KP>::CONV_GROUPS != grp) continue;
                const uint32_t s = it % NSS, a = it % NAS;
                if (kd) k_t = clock64();
                if (lane == 0) {
                    mbar_wait(&bars.full[s], (it / NSS) & 1);
                    mbar_wait(&bars.aempty[a], ((it / NAS) & 1) ^ 1);
                }
                __syncwarp();
                tcgen05_fence_after();
                if (kd) { const long long c = clock64(); k_w += c - k_t; k_t = c; }
                unsigned char* stg = smem + (size_t)s * Cfg::STAGE_BYTES;
                {   // A: column `col` of the tile over the chunk's 32 frames
                    const float* xs = reinterpret_cast<const float*>(stg + q * Cfg::SLAB) + lane;
                    float hi[32], cx[32];
#pragma unroll
                    for (int t = 0; t < 32; ++t) hi[t] = xs[t * 32];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cross_pack8(hi + 8 * j, cx + 8 * j, true);
                    tmem_st_32x32(tlane + a * 64, hi);
                    tmem_st_32x32(tlane + a * 64 + 32, cx);
                }
                if (kd) { const long long c = clock64(); k_a += c - k_t; k_t = c; }
                if (col < KP) {   // B cross operand: row n = col, 4 groups of 8 frames -> 8 chunks of 16 bytes (128-byte swizzle)
                    const float* hs = reinterpret_cast<const float*>(stg + Cfg::A_BYTES + Cfg::B_BYTES + (col >> 5) * Cfg::SLAB) + (col & 31);
                    unsigned char* br = stg + Cfg::A_BYTES + 2 * Cfg::B_BYTES + col * 128;
                    float v[32], w[8];
#pragma unroll
                    for (int t = 0; t < 32; ++t) v[t] = hs[t * 32];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        cross_pack8(v + 8 * g, w, false);
                        *reinterpret_cast<float4*>(br + (((2 * g) ^ (col & 7)) << 4)) = make_float4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<float4*>(br + (((2 * g + 1) ^ (col & 7)) << 4)) = make_float4(w[4], w[5], w[6], w[7]);
                    }
                    fence_proxy_async_smem();
                }
                if (kd) { const long long c = clock64(); k_b += c - k_t; k_t = c; }
                tmem_wait_st();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.conv[a]);
                if (kd) { k_s += clock64() - k_t; ++k_n; }
            }
        }
        if (kd && grp == 0) { dbg[8] = k_w; dbg[9] = k_a; dbg[10] = k_b; dbg[11] = k_s; dbg[12] = k_n; }
    } else if (warp >= XtShape<KP>::SWEEP_WARP0) {
        // ---------------- epilogue: rows of the accumulator straight to the partial buffers ----------------
        const int q = warp & 3;
        uint32_t tl = 0;
        while (items.next(b, split, mt)) {
            const uint32_t buf = tl & 1;
            const bool empty_item = chunks_of(b, split, mt) == 0;           // nothing was accumulated: the partial is zero
            if (lane == 0) mbar_wait(&bars.dfull[buf], (tl >> 1) & 1);
            __syncwarp();
            tcgen05_fence_after();
            const int vcol = 128 * mt + q * 32 + lane;             // virtual column = output row
            float* dst = nullptr;
            if (vcol < 32 * nxs) {
                if (vcol < F) dst = xht_partial + ((((long long)b * S + split) * F) + vcol) * KP;
            } else if (vcol - 32 * nxs < KP) {
                dst = gram_partial + ((((long long)b * S + split) * KP) + (vcol - 32 * nxs)) * KP;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < KP; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + Cfg::COL_D + buf * KP + c0, v);
                if (empty_item) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                if (dst) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.dempty[buf]);
            ++tl;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}


static inline const CUtensorMap& as_map(const TcMapBlob& b) { return *reinterpret_cast<const CUtensorMap*>(&b); }

template <int KP>
static cudaError_t ts_hstep_impl(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    using Cfg = TsCfg<KP>;
    static int n_sm = 0;
    cudaError_t e;
    if (!n_sm) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if ((e = cudaFuncSetAttribute(h_step_ts_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)) != cudaSuccess) return e;
        attr_set = true;
    }
    static int exp_flags = -1;
    if (exp_flags < 0) { const char* x_ = getenv("AINMF_TS_EXP"); exp_flags = x_ ? atoi(x_) : 0; }
    static long long* dbg = nullptr;
    static int dbg_left = -1;
    if (dbg_left < 0) {
        const char* e_ = getenv("AINMF_TC_DEBUG");
        dbg_left = (e_ && e_[0] == '1') ? 2 : 0;
        if (dbg_left) { cudaMalloc((void**)&dbg, 8 * 84 * sizeof(long long)); }
    }
    if (dbg_left > 0) cudaMemsetAsync(dbg, 0, 8 * 84 * sizeof(long long), s);
    const int nH = ceil_div(p.T, TS_M);
    const long long tiles = (long long)p.B * nH;
    const int grid = (int)(tiles < n_sm ? tiles : n_sm);
    AINMF_LAUNCH(h_step_ts_kernel<KP>, dim3(grid), dim3(TsShape<KP>::THREADS), Cfg::SMEM_BYTES, s, as_map(wk.tc->mapX), as_map(wk.tc->mapWt),
                 as_map(wk.tc->mapWtLo), as_map(wk.tc->mapHk), as_map(wk.tc->mapG), as_map(wk.tc->mapGlo), p.F, p.T, p.B,
                 wk.tc_blobs, wk.tc_scal, p.Ht, p.h_stride, wk.violH, p.state, dbg_left > 0 ? dbg : nullptr, exp_flags, p.t_good, wk.tc_vfill);
    if (dbg_left > 0) {
        --dbg_left;
        long long hbuf[8 * 84];
        cudaStreamSynchronize(s);
        cudaMemcpy(hbuf, dbg, sizeof hbuf, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ts-debug h_step KP=%d grid=%d tiles=%lld] per tile of CTA 0: contraction start/end, sweep start/end (cycles)\n", KP, grid, tiles);
        for (int i = 0; i < 64 && (i == 0 || hbuf[8 * i + 1]); ++i)
            fprintf(stderr, "  tile %2d: mma %8lld .. %8lld   sweep %8lld .. %8lld   (tmem ld %lld, delta st+arrive %lld, update wait %lld)\n", i, hbuf[8 * i], hbuf[8 * i + 1], hbuf[8 * i + 2], hbuf[8 * i + 3], hbuf[8 * i + 4], hbuf[8 * i + 5], hbuf[8 * i + 6]);
        for (int blk = 0; blk < KP / 8; ++blk) {
            const long long* q = hbuf + 8 * 66 + 8 * blk;
            fprintf(stderr, "  tile 2 block %2d: start %lld coords +%lld store +%lld wait +%lld ld +%lld st +%lld lookahead +%lld\n", blk, q[0], q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5]);   // coords | store | delta st issue | update wait | ld issue + look-ahead | st visible + ld arrive
        }
        const long long* k = hbuf + 8 * 64;
        const long long* qq = hbuf + 8 * 65;
        if (qq[4]) fprintf(stderr, "  MMA issuer, cycles per chunk: (unused %lld) wait TMA %lld, wait converters %lld, issue 8 MMAs + 2 commits %lld  (%lld chunks)\n", qq[0] / qq[4], qq[1] / qq[4], qq[2] / qq[4], qq[3] / qq[4], qq[4]);
        if (k[4]) fprintf(stderr, "  converter thread 0, cycles per chunk: wait TMA %lld, wait A-stage free %lld, read+split %lld, tmem st+arrive %lld  (%lld chunks)\n", k[0] / k[4], k[1] / k[4], k[2] / k[4], k[3] / k[4], k[4]);
    }
    return cudaGetLastError();
}

cudaError_t nmf_ts_hstep(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    return p.KP == 64 ? ts_hstep_impl<64>(p, wk, s) : ts_hstep_impl<128>(p, wk, s);
}

template <int KP>
static cudaError_t ts_half1_impl(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    using Cfg = XtCfg<KP>;
    static int n_sm = 0;
    cudaError_t e;
    if (!n_sm) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(xht_ts_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)) != cudaSuccess) return e;
    }
    const long long items = (long long)p.B * wk.tc_splits * wk.tc_mtiles;
    const int grid = (int)(items < n_sm ? items : n_sm);
    static long long* dbg = nullptr;
    static int dbg_left = -1;
    if (dbg_left < 0) {
        const char* e_ = getenv("AINMF_TC_DEBUG");
        dbg_left = (e_ && e_[0] == '1') ? 2 : 0;
        if (dbg_left) cudaMalloc((void**)&dbg, 16 * sizeof(long long));
    }
    if (dbg_left > 0) cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), s);
    AINMF_LAUNCH(xht_ts_kernel<KP>, dim3(grid), dim3(XtShape<KP>::THREADS), Cfg::SMEM_BYTES, s, as_map(wk.tc->mapXs), as_map(wk.tc->mapHs),
                 as_map(wk.tc->mapHmn), p.F, p.T, p.B, wk.tc_splits, wk.tc_mtiles, wk.tc_fps, wk.xht_partial, wk.gram_partial, p.state, p.t_good,
                 dbg_left > 0 ? dbg : nullptr);
    if (dbg_left > 0) {
        --dbg_left;
        long long hb[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, dbg, sizeof hb, cudaMemcpyDeviceToHost);
        const double n = hb[6] > 0 ? (double)hb[6] : 1.0, nc = hb[12] > 0 ? (double)hb[12] : 1.0;
        fprintf(stderr, "[ts-debug xht KP=%d grid=%d items=%lld] CTA 0: %lld chunks in %lld items; cycles per chunk: producer waits for a free stage %.0f | "
                        "issuer: accumulator free %.0f, converters %.0f, TMA %.0f, issue+commit %.0f | converter group 0 (per chunk it takes): waits %.0f, A operand %.0f, "
                        "B cross operand %.0f, st visible + arrive %.0f\n", KP, grid, items, hb[6], hb[7], hb[0] / n, hb[2] / n, hb[3] / n, hb[4] / n, hb[5] / n,
                hb[8] / nc, hb[9] / nc, hb[10] / nc, hb[11] / nc);
    }
    return cudaGetLastError();
}
cudaError_t nmf_ts_half1(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    return p.KP == 64 ? ts_half1_impl<64>(p, wk, s) : ts_half1_impl<128>(p, wk, s);
}

}  // namespace ainmf
#endif  // AINMF_EMU
