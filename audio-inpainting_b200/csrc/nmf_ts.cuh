// nmf_ts.cuh -- pieces shared by the tensor-core kernels (nmf_ts.cu) and the W-side kernel (nmf_cd.cu).
#pragma once
#ifndef AINMF_EMU
#include "kernels.h"
#include "tc.cuh"

namespace ainmf {

constexpr int TS_SC = kSweepScalars;  // floats of sweep scalars per block: G diagonal block 8x8, look-ahead block 8x8, 1/diag

// Operands the H step derives from G = W^T W, for block `blk` of 8 coordinates of one clip (called by `nthreads`
// threads with index `tid`):
//   blob : the rows G[8blk .. 8blk+8)[0..KP) laid out as the K-major (no swizzle) [N = KP][K = 8] operand of the sweep
//          update -- raw values (tf32 main term), then the bf16 cross operand (tc::cross_pack8, B-side order);
//   sc   : what the sweep threads read themselves: the diagonal block G[8blk+j][8blk+i], the look-ahead block
//          G[8blk+i][8(blk+1)+c] and the reciprocals of the block's diagonal (0 where the diagonal is 0);
//   GX   : the bf16 cross operand of G's rows for the contraction's Ht.G chunks (same footprint as G).
__device__ __forceinline__ void g_prep_block(const float* __restrict__ Gb, float* __restrict__ GXb, float* __restrict__ blob,
                                             float* __restrict__ sc, int KP, int blk, int tid, int nthreads) {
    const int nblk = KP / 8;
    for (int idx = tid; idx < 8 * KP; idx += nthreads) {
        const int j = idx / KP, n = idx - j * KP;
        blob[((j >> 2) * KP + n) * 4 + (j & 3)] = Gb[(8 * blk + j) * KP + n];
    }
    for (int n = tid; n < KP; n += nthreads) {
        float v[8], w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = Gb[(8 * blk + j) * KP + n];          // G[8blk+j][n] = G[n][8blk+j]
        tc::cross_pack8(v, w, false);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            blob[8 * KP + ((j >> 2) * KP + n) * 4 + (j & 3)] = w[j];
            GXb[n * KP + 8 * blk + j] = w[j];
        }
    }
    for (int e = tid; e < 64; e += nthreads) {
        const int j = e >> 3, i = e & 7;
        sc[8 * j + i] = Gb[(8 * blk + j) * KP + 8 * blk + i];
        sc[64 + 8 * j + i] = (blk + 1 < nblk) ? Gb[(8 * blk + j) * KP + 8 * (blk + 1) + i] : 0.f;
    }
    for (int e = tid; e < 8; e += nthreads) {
        const float d = Gb[(8 * blk + e) * (KP + 1)];
        sc[128 + e] = (d != 0.f) ? 1.0f / d : 0.f;
    }
}

}  // namespace ainmf
#endif
