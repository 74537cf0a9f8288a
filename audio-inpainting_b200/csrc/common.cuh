// common.cuh -- shared declarations for the ainmf kernels (sm_100a).
//
// Conventions used by every kernel in this directory
//   * spectrogram arrays are FRAME-MAJOR: V[b][t][f], leading dimension ldf (multiple of 4 floats,
//     pad bins are kept at zero).  This is the memory order scipy's STFT produces before its
//     moveaxis view ($SP/scipy/signal/_spectral_py.py:2395) and it makes a time-frame shard a
//     contiguous row block.
//   * factors: W[b][f][Kp], Ht[b][t][Kp] row-major; Kp = rank padded to {32,64,128}; pad components
//     are zero, which is inert under the coordinate-descent update (zero Gram diagonal -> skipped).
//   * a kernel never returns from part of a block before its last __syncthreads().
#pragma once

#ifdef AINMF_EMU
#include "emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
namespace ainmf { extern unsigned long long g_launch_count; }
#define AINMF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (++ainmf::g_launch_count, kernel<<<grid, block, smem, stream>>>(__VA_ARGS__))
#define AINMF_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

#include <stdint.h>

namespace ainmf {

constexpr int kThreads = 256;

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole block (kThreads threads); result valid in every thread.  `scratch` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += scratch[i];   // fixed order: deterministic
    return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += scratch[i];
    return r;
}

}  // namespace ainmf
