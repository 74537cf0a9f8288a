// mask.cu -- K2: column (frame) mask from the corrupted waveform, bit-exact with the reference.
//
// get_gap_mask (main4_NMF_gap.py:28-40; threshold 1e-4, fraction 0.9) and get_mask_from_signal
// (main4_NMF_mask.py:28-45; threshold 0.01, fraction 0.8):
//     g[i]   = |x[i]| < float32(threshold)
//     column c is bad  <=>  mean(g[ws:we]) > frac,  ws = max(0, c*hop - hop//2), we = min(N, c*hop + hop//2)
// mean(bool) = cnt/len exactly and cnt/len > num/den <=> den*cnt > num*len (integers; verified
// exhaustively by the CPU test-suite).  An empty window (numpy: mean of empty = nan) is not bad.
#include "kernels.h"

namespace ainmf {

// one warp per column; grid = (ceil(T / warps_per_block), B)
__global__ void __launch_bounds__(kThreads)
gap_mask_kernel(const float* __restrict__ x, long long x_stride, long long x_origin, long long x_avail,
                long long N, int hop, int t_begin, int T, float thr, int num, int den,
                unsigned char* __restrict__ bad, long long bad_stride) {
    const int warps = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int lc = blockIdx.x * warps + (threadIdx.x >> 5);     // local column
    const int b = blockIdx.y;
    if (lc >= T) return;                                         // whole warp; the kernel has no barrier
    const long long centre = (long long)(t_begin + lc) * hop;
    long long ws = centre - hop / 2, we = centre + hop / 2;
    if (ws < 0) ws = 0;
    if (we > N) we = N;
    const float* xb = x + (long long)b * x_stride;
    int cnt = 0;
    for (long long i = ws + lane; i < we; i += 32) {
        const long long li = i - x_origin;
        const float v = (li >= 0 && li < x_avail) ? xb[li] : 1e30f;   // caller provides the halo
        cnt += (fabsf(v) < thr) ? 1 : 0;
    }
    cnt = warp_sum_i(cnt);
    const long long len = we - ws;
    if (lane == 0)
        bad[(long long)b * bad_stride + lc] = (len > 0 && (long long)den * cnt > (long long)num * len) ? 1 : 0;
}

cudaError_t launch_gap_mask(const float* x, long long x_stride, long long x_origin, long long x_avail,
                            int B, long long N, int hop, int t_begin, int T, float thr, int num, int den,
                            unsigned char* bad, long long bad_stride, cudaStream_t s) {
    if (T <= 0 || B <= 0) return cudaSuccess;
    dim3 grid(ceil_div(T, kThreads / 32), B);
    AINMF_LAUNCH(gap_mask_kernel, grid, dim3(kThreads), 0, s, x, x_stride, x_origin, x_avail, N, hop, t_begin,
                 T, thr, num, den, bad, bad_stride);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kThreads)
range_mask_kernel(int T, int col_start, int col_end, unsigned char* __restrict__ bad, long long bad_stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < T) bad[(long long)blockIdx.y * bad_stride + c] = (c >= col_start && c < col_end) ? 1 : 0;
}

cudaError_t launch_range_mask(int B, int T, int col_start, int col_end, unsigned char* bad,
                              long long bad_stride, cudaStream_t s) {
    if (T <= 0 || B <= 0) return cudaSuccess;
    dim3 grid(ceil_div(T, kThreads), B);
    AINMF_LAUNCH(range_mask_kernel, grid, dim3(kThreads), 0, s, T, col_start, col_end, bad, bad_stride);
    return cudaGetLastError();
}

// Ascending indices of the set flags (np.array(bad_cols)) and their count; one block per clip,
// chunked block scan so the order is the natural one.
__global__ void __launch_bounds__(kThreads)
compact_kernel(const unsigned char* __restrict__ bad, long long bad_stride, int T, int* __restrict__ bad_idx,
               long long idx_stride, int* __restrict__ n_bad) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const unsigned char* fb = bad + (long long)b * bad_stride;
    int* ob = bad_idx + (long long)b * idx_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < T; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const int f = (c < T && fb[c]) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        const int before = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (f) ob[off + before] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_bad[b] = s_base;
}

cudaError_t launch_compact(const unsigned char* bad, long long bad_stride, int B, int T, int* bad_idx,
                           long long idx_stride, int* n_bad, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    AINMF_LAUNCH(compact_kernel, dim3(B), dim3(kThreads), 0, s, bad, bad_stride, T, bad_idx, idx_stride, n_bad);
    return cudaGetLastError();
}

}  // namespace ainmf
