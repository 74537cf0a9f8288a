// pcm.cu -- device front/back end of the scripts and layout transposes.
//   load_damaged_data: wav int16 -> mono mean -> float32 -> x / max|x|      (main4_NMF_gap.py:21-24)
//   save_result:       clip(-1,1) * 32767 -> int16 (truncation)              (main4_NMF_gap.py:76-77)
#include "kernels.h"

namespace ainmf {

// mono[b][n] = mean over channels (exact in float32 for int16 input with <= 2 channels; float64 mean then cast
// for more); peak_bits[b] = max |mono| as an ordered int (non-negative floats order like their bit patterns).
__global__ void __launch_bounds__(kThreads)
pcm_mono_kernel(const int16_t* __restrict__ pcm, long long N, int channels, float* __restrict__ x,
                int* __restrict__ peak_bits) {
    __shared__ float s_red[32];
    const int b = blockIdx.y;
    float m = 0.f;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
        const int16_t* p = pcm + ((long long)b * N + n) * channels;
        double s = 0.0;
        for (int c = 0; c < channels; ++c) s += (double)p[c];
        const float v = (float)(channels > 1 ? s / (double)channels : s);
        x[(long long)b * N + n] = v;
        m = fmaxf(m, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
        atomicMax(&peak_bits[b], __float_as_int(m));
    }
}

__global__ void __launch_bounds__(kThreads)
pcm_normalise_kernel(float* __restrict__ x, long long N, const int* __restrict__ peak_bits, float* __restrict__ peak) {
    const int b = blockIdx.y;
    const float pk = __int_as_float(peak_bits[b]);
    if (peak && blockIdx.x == 0 && threadIdx.x == 0) peak[b] = pk;
    if (!(pk > 0.f)) return;                                  // `if np.max(np.abs(data)) > 0`
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x)
        x[(long long)b * N + n] = __fdiv_rn(x[(long long)b * N + n], pk);   // IEEE division, not * (1/pk)
}

cudaError_t launch_load_pcm16(const int16_t* pcm, int B, long long N, int channels, float* x, int* peak_bits,
                              float* peak, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(peak_bits, 0, sizeof(int) * (size_t)B, s);
    if (e != cudaSuccess) return e;
    const int gx = (int)((N + kThreads * 8 - 1) / (kThreads * 8));
    dim3 grid(gx < 1 ? 1 : (gx > 4096 ? 4096 : gx), B);
    AINMF_LAUNCH(pcm_mono_kernel, grid, dim3(kThreads), 0, s, pcm, N, channels, x, peak_bits);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    AINMF_LAUNCH(pcm_normalise_kernel, grid, dim3(kThreads), 0, s, x, N, peak_bits, peak);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kThreads)
pcm_store_kernel(const float* __restrict__ y, long long count, int16_t* __restrict__ pcm) {
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += (long long)gridDim.x * blockDim.x) {
        float v = y[n];
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        pcm[n] = (int16_t)(int)(v * 32767.0f);                // float32 product, C cast truncates toward zero
    }
}

cudaError_t launch_store_pcm16(const float* y, long long count, int16_t* pcm, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    long long gx = (count + kThreads * 8 - 1) / (kThreads * 8);
    if (gx > 8192) gx = 8192;
    AINMF_LAUNCH(pcm_store_kernel, dim3((unsigned)gx), dim3(kThreads), 0, s, y, count, pcm);
    return cudaGetLastError();
}

// ---- layout transposes between the reference's (F, T) arrays and the internal [T][ld] -------------------
// dst[c][r] = src[r][c] for r < rows, c < cols; src pitch ld_src, dst pitch ld_dst; pad region of dst rows
// (columns rows..ld_dst) is zeroed when zero_pad is set.  grid = (ceil(cols/32), ceil(rows/32), B)
template <typename T>
__global__ void __launch_bounds__(kThreads)
transpose_kernel(const T* __restrict__ src, long long src_stride, int ld_src, int rows, int cols, T* __restrict__ dst,
                 long long dst_stride, int ld_dst, int zero_pad) {
    __shared__ T tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    T zero;
    memset(&zero, 0, sizeof(T));
    for (int i = ly; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + lx;
        tile[i][lx] = (r < rows && c < cols) ? src[(long long)b * src_stride + (long long)r * ld_src + c] : zero;
    }
    __syncthreads();
    for (int i = ly; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + lx;
        if (c < cols && (r < rows || (zero_pad && r < ld_dst)))
            dst[(long long)b * dst_stride + (long long)c * ld_dst + r] = (r < rows) ? tile[lx][i] : zero;
    }
}

template <typename T>
static cudaError_t launch_transpose_t(const T* src, long long src_stride, int ld_src, int rows, int cols, T* dst,
                                      long long dst_stride, int ld_dst, int zero_pad, int B, cudaStream_t s) {
    if (rows <= 0 || cols <= 0 || B <= 0) return cudaSuccess;
    const int rr = zero_pad ? (ld_dst > rows ? ld_dst : rows) : rows;
    AINMF_LAUNCH(transpose_kernel<T>, dim3(ceil_div(cols, 32), ceil_div(rr, 32), B), dim3(kThreads), 0, s, src,
                 src_stride, ld_src, rows, cols, dst, dst_stride, ld_dst, zero_pad);
    return cudaGetLastError();
}
cudaError_t launch_transpose_f32(const float* src, long long src_stride, int ld_src, int rows, int cols, float* dst,
                                 long long dst_stride, int ld_dst, int zero_pad, int B, cudaStream_t s) {
    return launch_transpose_t<float>(src, src_stride, ld_src, rows, cols, dst, dst_stride, ld_dst, zero_pad, B, s);
}
cudaError_t launch_transpose_c64(const float2* src, long long src_stride, int ld_src, int rows, int cols, float2* dst,
                                 long long dst_stride, int ld_dst, int zero_pad, int B, cudaStream_t s) {
    return launch_transpose_t<float2>(src, src_stride, ld_src, rows, cols, dst, dst_stride, ld_dst, zero_pad, B, s);
}

}  // namespace ainmf
