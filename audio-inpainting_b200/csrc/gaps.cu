// gaps.cu -- sample-level gap detectors and the linear-interpolation baseline of the sibling scripts, and the Part-0
// post-processing of main4_NMF.py: the callers / baselines either side of the NMF path.
//   find_main_gap   main3_AR_text_gap.py:34-49    first and last sample with |x| < thr
//   find_gaps       main3_AR_text_mask.py:30-52   maximal runs of |x| < thr longer than min_len samples
//   linear interp   linear_interp_part1.py:52-75  valid = |x| > thr; damaged samples <- np.interp over the valid ones
//   _blend_boundaries / SNR   main4_NMF.py:99-126
// One pass over the waveform per kernel (HBM-bound); a clip is cut into chunks of 2048 samples, a chunk summary pass
// and a per-clip carry pass give every chunk the nearest valid sample on either side.
#include "kernels.h"

#ifdef AINMF_EMU      // host build of the test harness: plain IEEE double operations (x86-64 baseline does not fuse them)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
#endif

namespace ainmf {

constexpr int GC = 2048;                 // samples per chunk = 256 threads x 8
constexpr int GPT = GC / kThreads;       // samples per thread

struct GapSummary {                      // per chunk
    long long first_valid, last_valid;   // N / -1 when the chunk has no valid sample
    long long first_gap, last_gap;       // N / -1 when the chunk has no gap sample
    int n_gap;
    int n_runs;                          // qualifying runs that END in this chunk (filled by run_count_kernel)
};

__device__ __forceinline__ bool is_gap_sample(float v, float thr, int inclusive) {
    const float a = fabsf(v);
    return inclusive ? !(a > thr) : (a < thr);       // linear_interp: damaged = not(|x| > thr); detectors: |x| < thr
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}

// grid = (chunks, B)
__global__ void __launch_bounds__(kThreads)
gap_summary_kernel(const float* __restrict__ x, long long x_stride, long long N, float thr, int inclusive,
                   GapSummary* __restrict__ sum /*[B][chunks]*/) {
    __shared__ long long s_fv[8], s_lv[8], s_fg[8], s_lg[8];
    __shared__ int s_ng[8];
    const int b = blockIdx.y, c = blockIdx.x;
    const long long i0 = (long long)c * GC + threadIdx.x * GPT;
    long long fv = N, lv = -1, fg = N, lg = -1;
    int ng = 0;
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
        const long long i = i0 + j;
        if (i < N) {
            if (is_gap_sample(x[(long long)b * x_stride + i], thr, inclusive)) { fg = fg < i ? fg : i; lg = i; ++ng; }
            else { fv = fv < i ? fv : i; lv = i; }
        }
    }
    fv = warp_min_ll(fv); lv = warp_max_ll(lv); fg = warp_min_ll(fg); lg = warp_max_ll(lg); ng = warp_sum_i(ng);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_fv[w] = fv; s_lv[w] = lv; s_fg[w] = fg; s_lg[w] = lg; s_ng[w] = ng; }
    __syncthreads();
    if (threadIdx.x == 0) {
        GapSummary g{N, -1, N, -1, 0, 0};
        for (int k = 0; k < kThreads / 32; ++k) {
            g.first_valid = s_fv[k] < g.first_valid ? s_fv[k] : g.first_valid;
            g.last_valid = s_lv[k] > g.last_valid ? s_lv[k] : g.last_valid;
            g.first_gap = s_fg[k] < g.first_gap ? s_fg[k] : g.first_gap;
            g.last_gap = s_lg[k] > g.last_gap ? s_lg[k] : g.last_gap;
            g.n_gap += s_ng[k];
        }
        sum[(long long)b * gridDim.x + c] = g;
    }
}

// Per clip: left[c] = last valid sample before chunk c (-1: none), right[c] = first valid sample after chunk c (N: none),
// run_off[c] = number of qualifying runs ending before chunk c; totals[b] = {first_gap, last_gap + 1, n_gap, n_runs}.
// One block per clip; each thread owns a contiguous segment of chunks (two-level scan).
__global__ void __launch_bounds__(kThreads)
gap_carry_kernel(const GapSummary* __restrict__ sum, int chunks, long long N, long long* __restrict__ left,
                 long long* __restrict__ right, int* __restrict__ run_off, long long* __restrict__ totals /*[B][4]*/) {
    __shared__ long long s_l[kThreads], s_r[kThreads], s_fg[kThreads], s_lg[kThreads], s_ng[kThreads];
    __shared__ int s_nr[kThreads];
    const int b = blockIdx.x, t = threadIdx.x;
    const GapSummary* sb = sum + (long long)b * chunks;
    const int per = (chunks + kThreads - 1) / kThreads;
    const int c0 = t * per, c1 = min(chunks, c0 + per);
    long long lv = -1, fv = N, fg = N, lg = -1, ng = 0;
    int nr = 0;
    for (int c = c0; c < c1; ++c) {
        lv = sb[c].last_valid > lv ? sb[c].last_valid : lv;
        fv = sb[c].first_valid < fv ? sb[c].first_valid : fv;
        fg = sb[c].first_gap < fg ? sb[c].first_gap : fg;
        lg = sb[c].last_gap > lg ? sb[c].last_gap : lg;
        ng += sb[c].n_gap;
        nr += sb[c].n_runs;
    }
    s_l[t] = lv; s_r[t] = fv; s_fg[t] = fg; s_lg[t] = lg; s_ng[t] = ng; s_nr[t] = nr;
    __syncthreads();
    long long cl = -1, cr = N;              // carries into this thread's segment
    int off = 0;
    for (int k = 0; k < t; ++k) { cl = s_l[k] > cl ? s_l[k] : cl; off += s_nr[k]; }
    for (int k = t + 1; k < kThreads; ++k) cr = s_r[k] < cr ? s_r[k] : cr;
    for (int c = c0; c < c1; ++c) {
        left[(long long)b * chunks + c] = cl;
        run_off[(long long)b * chunks + c] = off;
        cl = sb[c].last_valid > cl ? sb[c].last_valid : cl;
        off += sb[c].n_runs;
    }
    for (int c = c1 - 1; c >= c0; --c) {
        right[(long long)b * chunks + c] = cr;
        cr = sb[c].first_valid < cr ? sb[c].first_valid : cr;
    }
    if (t == 0) {
        long long a = N, z = -1, n = 0;
        int r = 0;
        for (int k = 0; k < kThreads; ++k) { a = s_fg[k] < a ? s_fg[k] : a; z = s_lg[k] > z ? s_lg[k] : z; n += s_ng[k]; r += s_nr[k]; }
        totals[4 * b + 0] = (z < 0) ? -1 : a;
        totals[4 * b + 1] = (z < 0) ? -1 : z + 1;
        totals[4 * b + 2] = n;
        totals[4 * b + 3] = r;
    }
}

// For the 8 consecutive samples of a thread: nearest valid index to the left (exclusive of gaps) and to the right, from a
// block-wide scan of the per-thread aggregates plus the chunk carries.  L[j] / R[j] are meaningful for gap samples.
__device__ __forceinline__ void nearest_valid(const bool (&gap)[GPT], long long i0, long long N, long long carry_l, long long carry_r,
                                              long long (&L)[GPT], long long (&R)[GPT], long long* s_l, long long* s_r) {
    long long lv = -1, fv = N;
#pragma unroll
    for (int j = 0; j < GPT; ++j)
        if (i0 + j < N && !gap[j]) { lv = i0 + j; fv = fv < i0 + j ? fv : i0 + j; }
    s_l[threadIdx.x] = lv;
    s_r[threadIdx.x] = fv;
    __syncthreads();
    long long cl = carry_l, cr = carry_r;
    for (int k = 0; k < (int)threadIdx.x; ++k) cl = s_l[k] > cl ? s_l[k] : cl;          // 256 threads: a short serial scan
    for (int k = kThreads - 1; k > (int)threadIdx.x; --k) cr = s_r[k] < cr ? s_r[k] : cr;
    long long run = cl;
#pragma unroll
    for (int j = 0; j < GPT; ++j) { L[j] = run; if (i0 + j < N && !gap[j]) run = i0 + j; }
    run = cr;
#pragma unroll
    for (int j = GPT - 1; j >= 0; --j) { R[j] = run; if (i0 + j < N && !gap[j]) run = i0 + j; }
    __syncthreads();
}

// linear_interp_part1.py:64-75: y = x on valid samples, np.interp(i, valid positions, valid values) on damaged ones
// (float64 arithmetic: slope = (fr - fl) / (xr - xl); y = slope * (i - xl) + fl; outside the valid range the end value)
__global__ void __launch_bounds__(kThreads)
interp_fill_kernel(const float* __restrict__ x, long long x_stride, long long N, float thr, const long long* __restrict__ left,
                   const long long* __restrict__ right, float* __restrict__ y, long long y_stride) {
    __shared__ long long s_l[kThreads], s_r[kThreads];
    const int b = blockIdx.y, c = blockIdx.x;
    const float* xb = x + (long long)b * x_stride;
    const long long i0 = (long long)c * GC + threadIdx.x * GPT;
    float v[GPT];
    bool gap[GPT];
#pragma unroll
    for (int j = 0; j < GPT; ++j) { v[j] = (i0 + j < N) ? xb[i0 + j] : 0.f; gap[j] = (i0 + j < N) && is_gap_sample(v[j], thr, 1); }
    long long L[GPT], R[GPT];
    nearest_valid(gap, i0, N, left[(long long)b * gridDim.x + c], right[(long long)b * gridDim.x + c], L, R, s_l, s_r);
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
        const long long i = i0 + j;
        if (i >= N) continue;
        float out = v[j];
        if (gap[j]) {
            if (L[j] < 0 && R[j] >= N) out = v[j];                       // no valid sample at all: the caller returns the input
            else if (L[j] < 0) out = xb[R[j]];
            else if (R[j] >= N) out = xb[L[j]];
            else {
                const double fl = (double)xb[L[j]], fr = (double)xb[R[j]];
                const double slope = __ddiv_rn(__dsub_rn(fr, fl), (double)(R[j] - L[j]));
                out = (float)__dadd_rn(__dmul_rn(slope, (double)(i - L[j])), fl);
            }
        }
        y[(long long)b * y_stride + i] = out;
    }
}

// find_gaps: a run [s, e) of gap samples ends in this chunk at e - 1; it qualifies when e - s > min_len.
// write == 0: count into sum[].n_runs;  write == 1: store the pairs at run_off (ascending).
__global__ void __launch_bounds__(kThreads)
gap_runs_kernel(const float* __restrict__ x, long long x_stride, long long N, float thr, int min_len,
                const long long* __restrict__ left, GapSummary* __restrict__ sum, const int* __restrict__ run_off, int write,
                long long* __restrict__ runs /*[B][max_runs][2]*/, int max_runs) {
    __shared__ long long s_l[kThreads], s_r[kThreads];
    __shared__ int s_cnt[kThreads];
    const int b = blockIdx.y, c = blockIdx.x;
    const float* xb = x + (long long)b * x_stride;
    const long long i0 = (long long)c * GC + threadIdx.x * GPT;
    bool gap[GPT];
#pragma unroll
    for (int j = 0; j < GPT; ++j) gap[j] = (i0 + j < N) && is_gap_sample(xb[i0 + j], thr, 0);
    const bool next_gap = (i0 + GPT < N) && is_gap_sample(xb[i0 + GPT], thr, 0);
    long long L[GPT], R[GPT];
    nearest_valid(gap, i0, N, left[(long long)b * gridDim.x + c], N, L, R, s_l, s_r);
    int cnt = 0;
    long long rs[GPT], re[GPT];
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
        const long long i = i0 + j;
        const bool ng = (j + 1 < GPT) ? gap[j + 1] : next_gap;
        if (i < N && gap[j] && !(i + 1 < N && ng)) {                 // last sample of a run
            const long long s = L[j] + 1, e = i + 1;
            if (e - s > min_len) { rs[cnt] = s; re[cnt] = e; ++cnt; }
        }
    }
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    if (!write) {
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < kThreads; ++k) tot += s_cnt[k];
            sum[(long long)b * gridDim.x + c].n_runs = tot;
        }
        return;
    }
    int off = run_off[(long long)b * gridDim.x + c];
    for (int k = 0; k < (int)threadIdx.x; ++k) off += s_cnt[k];
    for (int k = 0; k < cnt; ++k)
        if (off + k < max_runs) {
            runs[((long long)b * max_runs + off + k) * 2] = rs[k];
            runs[((long long)b * max_runs + off + k) * 2 + 1] = re[k];
        }
}

// main4_NMF.py:114-126 (_blend_boundaries): ground truth outside the gap, restored inside, linear cross-fades of
// blend_len samples either side; numpy mixes float32 arrays with a float64 ramp, so the arithmetic is float64.
__global__ void __launch_bounds__(kThreads)
blend_kernel(const float* __restrict__ raw, const float* __restrict__ restored, long long N, long long gs, long long ge,
             int blend_len, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float v = raw[i];
    if (i >= gs && i < ge) v = restored[i];
    else if (i >= gs - blend_len && i < gs) {
        const int k = (int)(i - (gs - blend_len));
        const double m = (k == blend_len - 1) ? 1.0 : __dmul_rn((double)k, __ddiv_rn(1.0, (double)(blend_len - 1)));
        v = (float)__dadd_rn(__dmul_rn((double)raw[i], __dsub_rn(1.0, m)), __dmul_rn((double)restored[i], m));
    } else if (i >= ge && i < ge + blend_len) {
        const int k = (int)(i - ge);
        const double m = (k == blend_len - 1) ? 1.0 : __dmul_rn((double)k, __ddiv_rn(1.0, (double)(blend_len - 1)));
        v = (float)__dadd_rn(__dmul_rn((double)raw[i], m), __dmul_rn((double)restored[i], __dsub_rn(1.0, m)));
    }
    out[i] = v;
}

// sums[0] = sum ref^2, sums[1] = sum (ref - est)^2 over [begin, end) in double; grid-stride, one atomic pair per block
__global__ void __launch_bounds__(kThreads)
snr_sums_kernel(const float* __restrict__ ref, const float* __restrict__ est, long long begin, long long end,
                double* __restrict__ sums) {
    __shared__ double s_a[kThreads / 32], s_b[kThreads / 32];
    double a = 0.0, d = 0.0;
    for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
        const double r = (double)ref[i], e = r - (double)est[i];
        a += r * r;
        d += e * e;
    }
    a = warp_sum_d(a); d = warp_sum_d(d);
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = d; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) { ta += s_a[k]; tb += s_b[k]; }
        sums[2 * blockIdx.x] = ta;
        sums[2 * blockIdx.x + 1] = tb;
    }
}

// Fixture producers: x[b][s : s + l] = 0 for the clip's gap list -- the zeroing loop of
// generate_part1_data.py:44-46 (gaps from create_random_mask, drawn on the host) and generate_part2_data.py:36-43.
// grid = (gaps per clip, B); a gap may extend past N (clipped) and gaps may overlap.
__global__ void __launch_bounds__(kThreads)
apply_gaps_kernel(float* __restrict__ x, long long x_stride, long long N, const long long* __restrict__ starts,
                  const long long* __restrict__ lens, int gaps_per_clip) {
    const int b = blockIdx.y, g = blockIdx.x;
    const long long s = starts[(long long)b * gaps_per_clip + g], l = lens[(long long)b * gaps_per_clip + g];
    if (s < 0 || l <= 0) return;
    const long long e = (s + l < N) ? s + l : N;
    for (long long i = s + threadIdx.x; i < e; i += blockDim.x) x[(long long)b * x_stride + i] = 0.f;
}

// ---- launchers --------------------------------------------------------------------------------------------------
size_t gaps_work_bytes(int B, long long N) {
    const long long chunks = (N + GC - 1) / GC;
    return (size_t)B * chunks * (sizeof(GapSummary) + 2 * sizeof(long long) + sizeof(int)) + (size_t)B * 4 * sizeof(long long) + 1024;
}
struct GapsWork { GapSummary* sum; long long *left, *right, *totals; int* run_off; int chunks; };
static GapsWork gaps_carve(void* base, int B, long long N) {
    GapsWork w;
    w.chunks = (int)((N + GC - 1) / GC);
    char* p = (char*)base;
    w.sum = (GapSummary*)p; p += sizeof(GapSummary) * (size_t)B * w.chunks;
    w.left = (long long*)p; p += sizeof(long long) * (size_t)B * w.chunks;
    w.right = (long long*)p; p += sizeof(long long) * (size_t)B * w.chunks;
    w.totals = (long long*)p; p += sizeof(long long) * (size_t)B * 4;
    w.run_off = (int*)p;
    return w;
}

// span[b] = {first gap sample, last gap sample + 1} or {-1, -1}; n_gap[b] = number of gap samples (either may be null)
cudaError_t launch_gap_span(const float* x, long long x_stride, int B, long long N, float thr, int inclusive, void* work,
                            long long* span, long long* n_gap, cudaStream_t s) {
    GapsWork w = gaps_carve(work, B, N);
    AINMF_LAUNCH(gap_summary_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, inclusive, w.sum);
    AINMF_LAUNCH(gap_carry_kernel, dim3(B), dim3(kThreads), 0, s, w.sum, w.chunks, N, w.left, w.right, w.run_off, w.totals);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (span && (e = cudaMemcpy2DAsync(span, 2 * sizeof(long long), w.totals, 4 * sizeof(long long), 2 * sizeof(long long), B,
                                       cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if (n_gap && (e = cudaMemcpy2DAsync(n_gap, sizeof(long long), w.totals + 2, 4 * sizeof(long long), sizeof(long long), B,
                                        cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_gap_runs(const float* x, long long x_stride, int B, long long N, float thr, int min_len, void* work,
                            long long* runs, int max_runs, int* n_runs, cudaStream_t s) {
    GapsWork w = gaps_carve(work, B, N);
    AINMF_LAUNCH(gap_summary_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, 0, w.sum);
    AINMF_LAUNCH(gap_carry_kernel, dim3(B), dim3(kThreads), 0, s, w.sum, w.chunks, N, w.left, w.right, w.run_off, w.totals);
    AINMF_LAUNCH(gap_runs_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, min_len, w.left, w.sum, w.run_off, 0,
                 runs, max_runs);
    AINMF_LAUNCH(gap_carry_kernel, dim3(B), dim3(kThreads), 0, s, w.sum, w.chunks, N, w.left, w.right, w.run_off, w.totals);
    AINMF_LAUNCH(gap_runs_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, min_len, w.left, w.sum, w.run_off, 1,
                 runs, max_runs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // n_runs[b] = totals[b][3] (int64 -> int32 on the host side of the ABI would need a sync; keep it on the device)
    if (n_runs) {
        static_assert(sizeof(long long) == 8, "layout");
        e = cudaMemcpy2DAsync(n_runs, sizeof(int), w.totals + 3, 4 * sizeof(long long), sizeof(int), B, cudaMemcpyDeviceToDevice, s);
    }
    return e;
}

cudaError_t launch_interp_fill(const float* x, long long x_stride, int B, long long N, float thr, void* work, float* y,
                               long long y_stride, long long* n_damaged, cudaStream_t s) {
    GapsWork w = gaps_carve(work, B, N);
    AINMF_LAUNCH(gap_summary_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, 1, w.sum);
    AINMF_LAUNCH(gap_carry_kernel, dim3(B), dim3(kThreads), 0, s, w.sum, w.chunks, N, w.left, w.right, w.run_off, w.totals);
    AINMF_LAUNCH(interp_fill_kernel, dim3(w.chunks, B), dim3(kThreads), 0, s, x, x_stride, N, thr, w.left, w.right, y, y_stride);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_damaged) e = cudaMemcpy2DAsync(n_damaged, sizeof(long long), w.totals + 2, 4 * sizeof(long long), sizeof(long long), B,
                                         cudaMemcpyDeviceToDevice, s);
    return e;
}

cudaError_t launch_apply_gaps(float* x, long long x_stride, int B, long long N, const long long* starts, const long long* lens,
                              int gaps_per_clip, cudaStream_t s) {
    AINMF_LAUNCH(apply_gaps_kernel, dim3(gaps_per_clip, B), dim3(kThreads), 0, s, x, x_stride, N, starts, lens, gaps_per_clip);
    return cudaGetLastError();
}

cudaError_t launch_blend(const float* raw, const float* restored, long long N, long long gs, long long ge, int blend_len,
                         float* out, cudaStream_t s) {
    AINMF_LAUNCH(blend_kernel, dim3((unsigned)ceil_div64(N, kThreads)), dim3(kThreads), 0, s, raw, restored, N, gs, ge, blend_len, out);
    return cudaGetLastError();
}

// sums: device scratch of 2 * 256 doubles; the caller adds the 256 block pairs (fixed order) on the host
cudaError_t launch_snr_sums(const float* ref, const float* est, long long begin, long long end, double* sums, cudaStream_t s) {
    AINMF_LAUNCH(snr_sums_kernel, dim3(256), dim3(kThreads), 0, s, ref, est, begin, end, sums);
    return cudaGetLastError();
}

}  // namespace ainmf
