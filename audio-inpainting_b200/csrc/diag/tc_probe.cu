// tc_probe.cu -- diagnostic single-CTA GEMM that exercises exactly the TMA / descriptor / tcgen05 / TMEM encodings
// the tensor-core NMF kernels rely on (K-major and MN-major operands, SWIZZLE_128B, kind::tf32, 3-pass split).
// Exported as ainmf_tc_probe for tests/test_gpu_tc.py; not part of the inpainting path.
#include <stdio.h>

#include "../tc.cuh"

#ifndef AINMF_EMU
namespace ainmf {
using namespace tc;

constexpr int PM = 128;       // M
constexpr int PBK = 32;       // contraction elements per stage (one 128-byte swizzle row)

// mode 0: A [128][Kd], B [N][Kd]  (K-major both);  mode 1: At [Kd][128], Bt [Kd][N] (MN-major both)
// mode 2: as mode 0, but the A operand is placed in TMEM by the threads (tcgen05.st: lane = row, one column per
//         contraction element; raw bits as a_hi, a_lo beside it) and the MMA takes A from TMEM, B from shared memory
// split: 0 = single pass on the raw fp32 bits; 1 = hi/lo split + 3 passes
template <int N>
__global__ void __launch_bounds__(128)
tc_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int Kd, int mode,
                int split, float* __restrict__ D, const float* __restrict__ Ag, const float* __restrict__ Bx) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int stages = Kd / PBK;
    const uint32_t a_bytes = PM * PBK * 4, b_bytes = N * PBK * 4;
    float* sA = reinterpret_cast<float*>(smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u));   // 1024-B aligned, [stages][a]
    float* sB = sA + (size_t)stages * PM * PBK;                          // [stages][b]
    float* sAlo = sB + (size_t)stages * N * PBK;
    float* sBlo = sAlo + (size_t)stages * PM * PBK;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
    }
    const uint32_t ncols = (mode >= 2) ? 512 : (N < 32 ? 32 : N);
    if (warp == 1) tmem_alloc(&tmem_slot, ncols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_full, (uint32_t)stages * (a_bytes + b_bytes));
        for (int s = 0; s < stages; ++s) {
            if (mode == 0 || mode == 2 || mode == 3) {
                tma_load_3d(sA + (size_t)s * PM * PBK, &mapA, &bar_full, s * PBK, 0, 0);
                tma_load_3d(sB + (size_t)s * N * PBK, &mapB, &bar_full, s * PBK, 0, 0);
            } else {
                for (int j = 0; j < PM / 32; ++j)
                    tma_load_3d(sA + (size_t)s * PM * PBK + j * (32 * PBK), &mapA, &bar_full, j * 32, s * PBK, 0);
                for (int j = 0; j < N / 32; ++j)
                    tma_load_3d(sB + (size_t)s * N * PBK + j * (32 * PBK), &mapB, &bar_full, j * 32, s * PBK, 0);
            }
        }
    }
    mbar_wait(&bar_full, 0);
    if (split && mode != 3) {       // generic-proxy rewrite of the tiles: hi in place, lo beside; then make it visible to the MMA
        const int na = stages * PM * PBK, nb = stages * N * PBK;
        for (int i = threadIdx.x; i < na; i += blockDim.x) { float h, l; split_tf32(sA[i], h, l); sA[i] = h; sAlo[i] = l; }
        for (int i = threadIdx.x; i < nb; i += blockDim.x) { float h, l; split_tf32(sB[i], h, l); sB[i] = h; sBlo[i] = l; }
        fence_proxy_async_smem();
    }
    if (mode == 3) {
        // A: raw tf32 at [N, N+Kd); packed bf16 cross operand at [N+Kd, N+2Kd): per group of 8 contraction elements,
        // 8 columns = 16 bf16: elements 0-7 = A_lo, 8-15 = A_hi.  B cross tile: copied from Bx into sBlo in the
        // SWIZZLE_128B K-major layout the TMA would produce (row n, 16-byte chunk c at (c ^ (n & 7))).
        const int row = threadIdx.x;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + N;
        for (int c0 = 0; c0 < Kd; c0 += 8) {
            float hi[8], lo[8], px[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float v = Ag[(size_t)row * Kd + c0 + j]; float h; split_tf32(v, h, lo[j]); hi[j] = v; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                px[j] = __uint_as_float(pack_bf16x2(lo[2 * j], lo[2 * j + 1]));
                px[4 + j] = __uint_as_float(pack_bf16x2(hi[2 * j], hi[2 * j + 1]));
            }
            tmem_st_32x8(tbase + c0, hi);
            tmem_st_32x8(tbase + Kd + c0, px);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        for (int s = 0; s < stages; ++s)
            for (int i = threadIdx.x; i < N * PBK; i += blockDim.x) {
                const int n = i / PBK, k = i % PBK;
                const int chunk = k >> 2;
                sBlo[(size_t)s * N * PBK + n * PBK + ((chunk ^ (n & 7)) << 2) + (k & 3)] = Bx[(size_t)n * Kd + s * PBK + k];
            }
        fence_proxy_async_smem();
    }
    if (mode == 2) {   // thread = row of A: raw values -> TMEM columns [N, N+Kd), residuals -> [N+Kd, N+2Kd)
        const int row = threadIdx.x;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + N;
        for (int c0 = 0; c0 < Kd; c0 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float v = Ag[(size_t)row * Kd + c0 + j]; float h; split_tf32(v, h, lo[j]); hi[j] = v; }
            tmem_st_32x8(tbase + c0, hi);
            tmem_st_32x8(tbase + Kd + c0, lo);
        }
        tmem_wait_st();
        tcgen05_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc_tf32(PM, N, mode == 1, mode == 1);
        uint32_t acc = 0;
        for (int s = 0; s < stages; ++s) {
            for (int k8 = 0; k8 < PBK / 8; ++k8) {
                if (mode == 3) {
                    const uint64_t dbh = make_smem_desc(smem_u32(sB + (size_t)s * N * PBK) + k8 * 32, 16, 1024);
                    const uint64_t dbx = make_smem_desc(smem_u32(sBlo + (size_t)s * N * PBK) + k8 * 32, 16, 1024);
                    mma_tf32_ts(tmem, tmem + N + s * PBK + k8 * 8, dbh, idesc, acc);
                    acc = 1;
                    if (split) mma_bf16_ts(tmem, tmem + N + Kd + s * PBK + k8 * 8, dbx, make_idesc_bf16(PM, N, 0, 0), 1);
                    continue;
                }
                const int passes = split ? 3 : 1;
                for (int pss = 0; pss < passes; ++pss) {
                    const float* a = (pss == 2 ? sAlo : sA) + (size_t)s * PM * PBK;
                    const float* b = (pss == 1 ? sBlo : sB) + (size_t)s * N * PBK;
                    uint64_t da, db;
                    if (mode == 2) {
                        db = make_smem_desc(smem_u32(b) + k8 * 32, 16, 1024);
                        mma_tf32_ts(tmem, tmem + N + (pss == 2 ? Kd : 0) + s * PBK + k8 * 8, db, idesc, acc);
                        acc = 1;
                        continue;
                    }
                    if (mode == 0) {
                        da = make_smem_desc(smem_u32(a) + k8 * 32, 16, 1024);
                        db = make_smem_desc(smem_u32(b) + k8 * 32, 16, 1024);
                    } else {
                        da = make_smem_desc(smem_u32(a) + k8 * 1024, 32 * PBK * 4, 512, kLayoutSw128Base32);
                        db = make_smem_desc(smem_u32(b) + k8 * 1024, 32 * PBK * 4, 512, kLayoutSw128Base32);
                    }
                    mma_tf32_ss(tmem, da, db, idesc, acc);
                    acc = 1;
                }
            }
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tcgen05_fence_after();
    // warp w owns TMEM lanes 32w .. 32w+31 = rows of D
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        const int row = warp * 32 + (threadIdx.x & 31);
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, ncols);
}

}  // namespace ainmf

extern "C" int ainmf_tc_probe_x(int mode, int split, int N, int Kd, const float* A, const float* B, const float* Bx, float* D, void* stream);
extern "C" int ainmf_tc_probe(int mode, int split, int N, int Kd, const float* A, const float* B, float* D, void* stream) {
    return ainmf_tc_probe_x(mode, split, N, Kd, A, B, nullptr, D, stream);
}
extern "C" int ainmf_tc_probe_x(int mode, int split, int N, int Kd, const float* A, const float* B, const float* Bx, float* D, void* stream) {
    using namespace ainmf;
    if ((N != 64 && N != 128) || Kd % PBK != 0 || Kd <= 0 || Kd > 128) return -1;
    CUtensorMap ma, mb;
    int rc;
    if (mode == 0 || mode == 2 || mode == 3) {
        rc = make_tensor_map_3d(&ma, A, Kd, PM, 1, Kd, (uint64_t)Kd * PM, PBK, PM);
        if (!rc) rc = make_tensor_map_3d(&mb, B, Kd, N, 1, Kd, (uint64_t)Kd * N, PBK, N);
    } else {
        rc = make_tensor_map_3d(&ma, A, PM, Kd, 1, PM, (uint64_t)Kd * PM, 32, PBK, 1);
        if (!rc) rc = make_tensor_map_3d(&mb, B, N, Kd, 1, N, (uint64_t)Kd * N, 32, PBK, 1);
    }
    if (rc) return 1000 + rc;
    const int stages = Kd / PBK;
    const size_t smem = 2 * (size_t)stages * (PM + N) * PBK * 4 + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (N == 64) {
        e = cudaFuncSetAttribute(tc_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) tc_probe_kernel<64><<<1, 128, smem, s>>>(ma, mb, Kd, mode, split, D, A, Bx);
    } else {
        e = cudaFuncSetAttribute(tc_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) tc_probe_kernel<128><<<1, 128, smem, s>>>(ma, mb, Kd, mode, split, D, A, Bx);
    }
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}
#endif

