// sweep_test.cu -- diagnostic entry point ainmf_test_sweep_inc: runs cd_sweep_rows_inc<KP,8,4> (the sweep the
// tensor-core h-step uses) on caller-provided A, G, B so that it can be compared with the reference sweep
// (_cdnmf_fast.pyx:8-38) in isolation -- through the CPU emulator harness and on the GPU.  Not on the product path.
#include "../kernels.h"
#include "../nmf_cd.cuh"

namespace ainmf {

template <int KP>
__global__ void __launch_bounds__(kThreads)
sweep_inc_test_kernel(float* __restrict__ A, const float* __restrict__ G, const float* __restrict__ Bm, int rows,
                      float* __restrict__ viol, long long* __restrict__ cyc) {
    constexpr int L = 8, R = 4, S = KP / L, PITCH = KP + 4 * L;
    AINMF_DYN_SMEM(smem_raw);
    float* sG = reinterpret_cast<float*>(smem_raw);      // [KP][PITCH]
    float* sInv = sG + KP * PITCH;                       // [KP]
    __shared__ float s_red[32];
    load_gram_padded<KP, L>(sG, G);
    for (int t = threadIdx.x; t < KP; t += blockDim.x) { const float d = G[t * KP + t]; sInv[t] = (d != 0.f) ? 1.0f / d : 0.f; }
    __syncthreads();
    const int l = threadIdx.x % L, grp = threadIdx.x / L;
    const int row0 = blockIdx.x * 128 + grp * R;
    float a[R][S], g[R][S];
    bool valid[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int row = row0 + i;
        valid[i] = row < rows;
#pragma unroll
        for (int q = 0; q < S; ++q) {
            a[i][q] = valid[i] ? A[(long long)row * KP + l * S + q] : 0.f;
            g[i][q] = 0.f;
        }
        if (valid[i]) {      // g0 = A.G - B, the slow way (the product path gets it from the tensor cores)
#pragma unroll
            for (int q = 0; q < S; ++q) {
                const int t = l * S + q;
                float acc = -Bm[(long long)row * KP + t];
                for (int r = 0; r < KP; ++r) acc = fmaf(A[(long long)row * KP + r], G[r * KP + t], acc);
                g[i][q] = acc;
            }
        }
    }
    __syncthreads();
#ifndef AINMF_EMU
    const long long c0 = clock64();
#endif
    const float v = cd_sweep_rows_inc<KP, L, R>(a, g, sG, sInv, l, valid);
#ifndef AINMF_EMU
    if (cyc && threadIdx.x == 0) cyc[blockIdx.x] = clock64() - c0;
#endif
#pragma unroll
    for (int i = 0; i < R; ++i)
        if (valid[i]) {
#pragma unroll
            for (int q = 0; q < S; ++q) A[(long long)(row0 + i) * KP + l * S + q] = a[i][q];
        }
    const float tot = block_sum(v, s_red);
    if (threadIdx.x == 0) viol[blockIdx.x] = tot;
}

template <int KP>
static cudaError_t run(float* A, const float* G, const float* Bm, int rows, float* viol, long long* cyc, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)KP * (KP + 32) + KP);
    cudaError_t e = cudaFuncSetAttribute(sweep_inc_test_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(sweep_inc_test_kernel<KP>, dim3(ceil_div(rows, 128)), dim3(kThreads), smem, s, A, G, Bm, rows, viol, cyc);
    return cudaGetLastError();
}

}  // namespace ainmf

// A [rows][KP] in/out, G [KP][KP], B [rows][KP], viol [ceil(rows/128)] out (device pointers); KP in {32,64,128}
extern "C" int ainmf_test_sweep_inc_timed(float* A, const float* G, const float* B, int rows, int KP, float* viol,
                                          long long* cycles, void* stream) {
    using namespace ainmf;
    cudaStream_t s = (cudaStream_t)stream;
    switch (KP) {
        case 32: return (int)run<32>(A, G, B, rows, viol, cycles, s);
        case 64: return (int)run<64>(A, G, B, rows, viol, cycles, s);
        case 128: return (int)run<128>(A, G, B, rows, viol, cycles, s);
    }
    return -1;
}

extern "C" int ainmf_test_sweep_inc(float* A, const float* G, const float* B, int rows, int KP, float* viol, void* stream) {
    using namespace ainmf;
    cudaStream_t s = (cudaStream_t)stream;
    switch (KP) {
        case 32: return (int)run<32>(A, G, B, rows, viol, nullptr, s);
        case 64: return (int)run<64>(A, G, B, rows, viol, nullptr, s);
        case 128: return (int)run<128>(A, G, B, rows, viol, nullptr, s);
    }
    return -1;
}
