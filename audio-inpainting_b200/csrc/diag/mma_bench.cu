// mma_bench.cu -- diagnostic microbenchmark of tcgen05.mma issue rate on sm_100a (libainmf_diag.so; not on the product path).
//
// One thread per CTA issues `iters` x `per` back-to-back tcgen05.mma on garbage operands and waits for the commit of the
// last one; cycles = clock64 from the first issue to the commit's arrival.  What it separates:
//   * n_acc = 1: every instruction accumulates into the SAME TMEM accumulator (a dependent chain);
//     n_acc > 1: instructions rotate over independent accumulators -> the issue-rate floor of the pipe;
//   * N in {8..256}, M in {64, 128}, kind tf32 (K = 8) / bf16 (K = 16), A from shared memory or from TMEM;
//   * cg2 = 1: cta_group::2 over a cluster of two CTAs (M = 128 or 256 per pair).
#include <stdio.h>

#include "../tc.cuh"

#ifdef AINMF_DIAG_LIB
namespace ainmf { unsigned long long g_launch_count = 0; }   // the diagnostics library counts its own launches
#endif

#ifndef AINMF_EMU
namespace ainmf {
using namespace tc;

__device__ __forceinline__ void mma_any(int kind_bf16, int ts, uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d, uint32_t idesc) {
    if (kind_bf16) { if (ts) mma_bf16_ts(d, a_t, b_d, idesc, 1); else mma_bf16_ss(d, a_d, b_d, idesc, 1); }
    else { if (ts) mma_tf32_ts(d, a_t, b_d, idesc, 1); else mma_tf32_ss(d, a_d, b_d, idesc, 1); }
}

// issue = 0: `if (threadIdx.x == 0)` around the whole loop (ptxas cannot prove the branch warp-uniform and wraps every
//            UTCHMMA in an ELECT / BRA.U.ANY loop);  1: the whole warp runs the loop, elect.sync around each instruction
//            (what CUTLASS does);  2: the whole warp enters, one elect.sync around the loop
__global__ void __launch_bounds__(128)
mma_bench_kernel(int M, int N, int kind_bf16, int ts, int n_acc, int iters, int per, int issue, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (issue != 0 && warp == 0) {
        const uint32_t idesc = kind_bf16 ? make_idesc_bf16(M, N, 0, 0) : make_idesc_tf32(M, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(base), 16, 1024);
        const uint64_t db = make_smem_desc(smem_u32(base) + 16384, 16, 1024);
        const long long t0 = clock64();
        long long t_issue = 0;
        if (issue == 1) {
            int acc = 0;
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int j = 0; j < per; ++j) {
                    const uint64_t o = (uint64_t)((j & 3) * 2);
                    if (elect_one()) mma_any(kind_bf16, ts, tmem + acc * N, tmem + 480 + (j & 3) * 8, da + o, db + o, idesc);
                    if (++acc == n_acc) acc = 0;
                }
            }
            t_issue = clock64() - t0;
            if (elect_one()) mma_commit(&bar);
        } else if (elect_one()) {
            int acc = 0;
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int j = 0; j < per; ++j) {
                    const uint64_t o = (uint64_t)((j & 3) * 2);
                    mma_any(kind_bf16, ts, tmem + acc * N, tmem + 480 + (j & 3) * 8, da + o, db + o, idesc);
                    if (++acc == n_acc) acc = 0;
                }
            }
            t_issue = clock64() - t0;
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        if (threadIdx.x == 0) {
            out[2 * blockIdx.x] = clock64() - t0;
            out[2 * blockIdx.x + 1] = t_issue;
        }
    }
    if (issue == 0 && threadIdx.x == 0) {
        const uint32_t idesc = kind_bf16 ? make_idesc_bf16(M, N, 0, 0) : make_idesc_tf32(M, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(base), 16, 1024);               // A: [M][32 B of K] rows of 128 B
        const uint64_t db = make_smem_desc(smem_u32(base) + 16384, 16, 1024);       // B: [N][...]
        const long long t0 = clock64();
        int acc = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int j = 0; j < per; ++j) {
                const uint64_t o = (uint64_t)((j & 3) * 2);
                mma_any(kind_bf16, ts, tmem + acc * N, tmem + 480 + (j & 3) * 8, da + o, db + o, idesc);
                if (++acc == n_acc) acc = 0;
            }
        }
        const long long t_issue = clock64() - t0;
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        out[2 * blockIdx.x] = clock64() - t0;
        out[2 * blockIdx.x + 1] = t_issue;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// Straight-line variant: everything that selects the instruction is a template parameter, the operands of a group of 8
// instructions are precomputed, the group is fully unrolled inside ONE elect.sync region -> the SASS is UTCHMMA after
// UTCHMMA with at most a uniform add between them: the hardware's own issue/execute floor.
template <int N, int BF16, int TS, int NACC>
__global__ void __launch_bounds__(128)
mma_bench_unrolled_kernel(int M, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        // BF16 = 2: kinds alternate tf32, bf16, tf32, ... (what the NMF contractions issue); 3: four tf32 then four bf16
        const uint32_t idesc_b = make_idesc_bf16(M, N, 0, 0), idesc_t = make_idesc_tf32(M, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(base), 16, 1024);
        const uint64_t db = make_smem_desc(smem_u32(base) + 16384, 16, 1024);
        const long long t0 = clock64();
        long long t_issue = 0;
        if (elect_one()) {
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint64_t o = (uint64_t)((j & 3) * 2);
                    const uint32_t d = tmem + (j % NACC) * N, a_t = tmem + 480 + (j & 3) * 8;
                    const bool bf = BF16 == 1 || (BF16 == 2 && (j & 1)) || (BF16 == 3 && j >= 4);
                    if (bf) { if (TS) mma_bf16_ts(d, a_t, db + o, idesc_b, 1); else mma_bf16_ss(d, da + o, db + o, idesc_b, 1); }
                    else { if (TS) mma_tf32_ts(d, a_t, db + o, idesc_t, 1); else mma_tf32_ss(d, da + o, db + o, idesc_t, 1); }
                }
            }
            t_issue = clock64() - t0;
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t_issue = __shfl_sync(0xffffffffu, t_issue, 0) + 0 * t_issue;
        if (threadIdx.x == 0) { out[2 * blockIdx.x] = clock64() - t0; }
        if (t_issue && (threadIdx.x & 31) == 0) out[2 * blockIdx.x + 1] = t_issue;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}
template <int N, int BF16, int TS, int NACC>
static int run_unrolled(int M, int iters, int blocks, long long* out, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(mma_bench_unrolled_kernel<N, BF16, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    if (e != cudaSuccess) return (int)e;
    mma_bench_unrolled_kernel<N, BF16, TS, NACC><<<blocks, 128, 66 * 1024, s>>>(M, iters, out);
    return (int)cudaGetLastError();
}

// cta_group::2: a cluster of two CTAs; CTA 0's thread issues, both CTAs' TMEM receive M/2... (M = 128: 64 rows each;
// M = 256: 128 rows each).  B is split between the two CTAs' shared memory (N/2 rows each), A is each CTA's own rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
mma_bench_cg2_kernel(int M, int N, int kind_bf16, int n_acc, int iters, int per, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (cta_rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = kind_bf16 ? make_idesc_bf16(M, N, 0, 0) : make_idesc_tf32(M, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(base), 16, 1024);
        const uint64_t db = make_smem_desc(smem_u32(base) + 16384, 16, 1024);
        const long long t0 = clock64();
        int acc = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int j = 0; j < per; ++j) {
                const uint64_t o = (uint64_t)((j & 3) * 2);
                const uint32_t d = tmem + acc * N;
                if (kind_bf16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(da + o), "l"(db + o), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(da + o), "l"(db + o), "r"(idesc), "r"(1u) : "memory");
                if (++acc == n_acc) acc = 0;
            }
        }
        const long long t_issue = clock64() - t0;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
        out[blockIdx.x + 1] = t_issue;
    }
    tcgen05_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
}  // namespace ainmf

// iters groups of 8 instructions; (N, bf16, ts, n_acc) must be one of the instantiated combinations
extern "C" int ainmf_diag_mma_bench_unrolled(int M, int N, int bf16, int ts, int n_acc, int iters, int blocks, long long* out, void* stream) {
    using namespace ainmf;
    cudaStream_t s = (cudaStream_t)stream;
#define AINMF_U(n, b, t, a) if (N == n && bf16 == b && ts == t && n_acc == a) return run_unrolled<n, b, t, a>(M, iters, blocks, out, s);
#define AINMF_UN(n) AINMF_U(n, 2, 1, 1) AINMF_U(n, 2, 1, 2) AINMF_U(n, 3, 1, 1) AINMF_U(n, 2, 0, 1) AINMF_U(n, 0, 1, 1) AINMF_U(n, 0, 1, 2) AINMF_U(n, 1, 1, 1) AINMF_U(n, 1, 1, 2) AINMF_U(n, 0, 0, 1) AINMF_U(n, 1, 0, 1)
    AINMF_UN(16) AINMF_UN(32) AINMF_UN(64) AINMF_UN(128) AINMF_UN(256)
    return -1;
}

extern "C" int ainmf_diag_mma_bench(int M, int N, int kind_bf16, int ts, int n_acc, int cg2, int iters, int per, int blocks,
                                    int issue, long long* out, void* stream) {
    using namespace ainmf;
    if (n_acc < 1 || n_acc * N > 480) return -1;
    cudaError_t e;
    if (cg2) {
        e = cudaFuncSetAttribute(mma_bench_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
        if (e != cudaSuccess) return (int)e;
        mma_bench_cg2_kernel<<<blocks & ~1, 128, 66 * 1024, (cudaStream_t)stream>>>(M, N, kind_bf16, n_acc, iters, per, out);
    } else {
        e = cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
        if (e != cudaSuccess) return (int)e;
        mma_bench_kernel<<<blocks, 128, 66 * 1024, (cudaStream_t)stream>>>(M, N, kind_bf16, ts, n_acc, iters, per, issue, out);
    }
    return (int)cudaGetLastError();
}
#endif
