// impute.cu -- K3: mean-imputation of the bad frames, initial factors, layout helpers.
//
//   avg_spec = mean(magnitude[:, good_cols], axis=1); current_mag[:, bad_cols] = avg_spec
//                                                        (main4_NMF_gap.py:55-59, main4_NMF.py:79-81)
//   avg = sqrt(X.mean() / n_components); H = |avg * N(0,1)| (K,T) drawn first, then W = |avg * N(0,1)| (F,K)
//                                                        ($SP/sklearn/decomposition/_nmf.py:296-307)
// The normals come from the host (MT19937 + legacy polar method restated in rng.cpp); this file scales them.
#include "kernels.h"

namespace ainmf {

void impute_plan(int T, ImputeWork* wk) {
    int fpc = ceil_div(T, 512);
    if (fpc < 64) fpc = 64;
    wk->frames_per_chunk = fpc;
    wk->n_chunks = ceil_div(T, fpc);
}
size_t impute_work_bytes(int B, int F, const ImputeWork& wk) {
    return sizeof(double) * ((size_t)B * wk.n_chunks * F + (size_t)B * (F + 1)) + 256;
}
void impute_carve(void* base, int B, int F, ImputeWork* wk) {
    wk->colsum = static_cast<double*>(base);
    wk->sums = wk->colsum + ((size_t)B * wk->n_chunks * F + 31) / 32 * 32;
}

// colsum[b][chunk][f] = sum over the frames of the chunk that are NOT excluded of V[t][f]  (double accumulation)
// grid = (n_chunks, ceil(F/256), B); excl may be null (sum everything).
__global__ void __launch_bounds__(kThreads)
colsum_kernel(const float* __restrict__ V, long long v_stride, int ldf, int F, int T,
              const unsigned char* __restrict__ excl, long long excl_stride, int frames_per_chunk,
              double* __restrict__ colsum) {
    const int b = blockIdx.z, chunk = blockIdx.x;
    const int f = blockIdx.y * blockDim.x + threadIdx.x;
    const int t0 = chunk * frames_per_chunk;
    const int t1 = min(T, t0 + frames_per_chunk);
    if (f >= F) return;                                     // no barrier in this kernel
    const float* Vb = V + (long long)b * v_stride;
    const unsigned char* eb = excl ? excl + (long long)b * excl_stride : nullptr;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int t = t0;
    for (; t + 3 < t1; t += 4) {
        const float v0 = Vb[(long long)t * ldf + f], v1 = Vb[(long long)(t + 1) * ldf + f];
        const float v2 = Vb[(long long)(t + 2) * ldf + f], v3 = Vb[(long long)(t + 3) * ldf + f];
        s0 += (eb && eb[t]) ? 0.0 : (double)v0;
        s1 += (eb && eb[t + 1]) ? 0.0 : (double)v1;
        s2 += (eb && eb[t + 2]) ? 0.0 : (double)v2;
        s3 += (eb && eb[t + 3]) ? 0.0 : (double)v3;
    }
    for (; t < t1; ++t) s0 += (eb && eb[t]) ? 0.0 : (double)Vb[(long long)t * ldf + f];
    colsum[((long long)b * gridDim.x + chunk) * F + f] = (s0 + s1) + (s2 + s3);
}

// sums[b][f] = sum_chunks colsum[b][chunk][f]; sums[b][F] = sum_f of those (grand total).  One block per clip.
// (In the time-frame-sharded mode the caller all-reduces `sums` across ranks before the next kernel.)
__global__ void __launch_bounds__(kThreads)
colsum_reduce_kernel(const double* __restrict__ colsum, int n_chunks, int F, double* __restrict__ sums /*[B][F+1]*/) {
    __shared__ double s_red[32];
    const int b = blockIdx.x;
    double local = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < n_chunks; ++c) s += colsum[((long long)b * n_chunks + c) * F + f];
        sums[(long long)b * (F + 1) + f] = s;
        local += s;
    }
    const double tot = block_sum_d(local, s_red);
    if (threadIdx.x == 0) sums[(long long)b * (F + 1) + F] = tot;
}

// fill[f] = sums[f] / n_good, n_good = T_total - n_excl; clip state initialised from the counts.
__global__ void __launch_bounds__(kThreads)
fill_kernel(const double* __restrict__ sums, int F, int ldf, int T_total, const int* __restrict__ n_bad,
            const int* __restrict__ n_excl, float* __restrict__ fill, ClipState* __restrict__ st) {
    const int b = blockIdx.x;
    const int nb = n_bad[b];
    const int n_good = T_total - n_excl[b];
    for (int f = threadIdx.x; f < ldf; f += blockDim.x)
        fill[(long long)b * ldf + f] = (f < F && n_good > 0) ? (float)(sums[(long long)b * (F + 1) + f] / (double)n_good) : 0.f;
    if (threadIdx.x == 0) {
        ClipState s;
        s.n_bad = nb;
        s.status = (nb == 0) ? 1 : (n_good <= 0 ? 2 : 0);
        s.done = (s.status != 0) ? 1 : 0;
        s.n_iter = 0;
        s.viol_init = 0.0;
        s.viol_last = 0.0;
        s.sum_x = 0.0;
        s.mean_x = 0.f;
        s.err = 0.f;
        st[b] = s;
    }
}

// mean_x = sums[F] / (F * T_total); also re-arms the iteration (done = status != 0, n_iter = 0).
__global__ void __launch_bounds__(kThreads)
mean_kernel(const double* __restrict__ sums, int F, int T_total, int B, ClipState* __restrict__ st) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ClipState s = st[b];
    s.sum_x = sums[(long long)b * (F + 1) + F];
    s.mean_x = (float)(s.sum_x / ((double)F * (double)T_total));
    s.done = (s.status != 0) ? 1 : 0;
    s.n_iter = 0;
    s.viol_init = 0.0;
    s.viol_last = 0.0;
    st[b] = s;
}

// bad frames <- fill; one warp per frame; grid = (ceil(T/8), B)
__global__ void __launch_bounds__(kThreads)
fill_rows_kernel(float* __restrict__ V, long long v_stride, int ldf, int T, const unsigned char* __restrict__ bad,
                 long long bad_stride, const float* __restrict__ fill) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T || !bad[(long long)b * bad_stride + t]) return;
    float* row = V + (long long)b * v_stride + (long long)t * ldf;
    const float* fl = fill + (long long)b * ldf;
    for (int f = lane; f < ldf; f += 32) row[f] = fl[f];
}

cudaError_t launch_colsums(const float* V, long long v_stride, int ldf, int F, int T, int B,
                           const unsigned char* excl, long long excl_stride, const ImputeWork& wk, cudaStream_t s) {
    AINMF_LAUNCH(colsum_kernel, dim3(wk.n_chunks, ceil_div(F, kThreads), B), dim3(kThreads), 0, s, V, v_stride, ldf,
                 F, T, excl, excl_stride, wk.frames_per_chunk, wk.colsum);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(colsum_reduce_kernel, dim3(B), dim3(kThreads), 0, s, wk.colsum, wk.n_chunks, F, wk.sums);
    return cudaGetLastError();
}

cudaError_t launch_fill(float* V, long long v_stride, int ldf, int F, int T, int T_total, int B,
                        const unsigned char* bad, long long bad_stride, const int* n_bad, const int* n_excl,
                        float* fill, ClipState* state, const ImputeWork& wk, cudaStream_t s) {
    AINMF_LAUNCH(fill_kernel, dim3(B), dim3(kThreads), 0, s, wk.sums, F, ldf, T_total, n_bad, n_excl, fill, state);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(fill_rows_kernel, dim3(ceil_div(T, kThreads / 32), B), dim3(kThreads), 0, s, V, v_stride, ldf, T,
                 bad, bad_stride, fill);
    return cudaGetLastError();
}

cudaError_t launch_mean(int F, int T_total, int B, ClipState* state, const ImputeWork& wk, cudaStream_t s) {
    AINMF_LAUNCH(mean_kernel, dim3(ceil_div(B, kThreads)), dim3(kThreads), 0, s, wk.sums, F, T_total, B, state);
    return cudaGetLastError();
}

// ---- initial factors ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
init_w_kernel(const float* __restrict__ Wn, int F, int K, int KP, const ClipState* __restrict__ st,
              float* __restrict__ W, long long w_stride) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * KP) return;
    const int f = (int)(i / KP), k = (int)(i % KP);
    const float avg = sqrtf(st[b].mean_x / (float)K);
    W[(long long)b * w_stride + i] = (k < K) ? fabsf(avg * Wn[(long long)f * K + k]) : 0.f;
}
// Ht[t][k] = |avg * Hn[k][t_begin + t]| via a 32x32 shared-memory transpose; grid = (ceil(T/32), ceil(KP/32), B)
__global__ void __launch_bounds__(kThreads)
init_h_kernel(const float* __restrict__ Hn, int T_total, int t_begin, int T, int K, int KP,
              const ClipState* __restrict__ st, float* __restrict__ Ht, long long h_stride) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;        // 32 x 8
    for (int r = ly; r < 32; r += 8) {
        const int k = k0 + r, t = t0 + lx;
        tile[r][lx] = (k < K && t < T) ? Hn[(long long)k * T_total + t_begin + t] : 0.f;
    }
    __syncthreads();
    const float avg = sqrtf(st[b].mean_x / (float)K);
    for (int r = ly; r < 32; r += 8) {
        const int t = t0 + r, k = k0 + lx;
        if (t < T && k < KP) Ht[(long long)b * h_stride + (long long)t * KP + k] = (k < K) ? fabsf(avg * tile[lx][r]) : 0.f;
    }
}

cudaError_t launch_init_factors(const float* Wn, const float* Hn, int T_total, int t_begin, int B, int F, int T,
                                int K, int KP, const ClipState* state, float* W, long long w_stride, float* Ht,
                                long long h_stride, cudaStream_t s) {
    AINMF_LAUNCH(init_w_kernel, dim3((unsigned)ceil_div64((long long)F * KP, kThreads), B), dim3(kThreads), 0, s, Wn,
                 F, K, KP, state, W, w_stride);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(init_h_kernel, dim3(ceil_div(T, 32), ceil_div(KP, 32), B), dim3(kThreads), 0, s, Hn, T_total,
                 t_begin, T, K, KP, state, Ht, h_stride);
    return cudaGetLastError();
}

// ---- user layout <-> internal layout ----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
pack_w_kernel(const float* __restrict__ W0, int F, int K, int KP, float* __restrict__ W, long long w_stride) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * KP) return;
    const int f = (int)(i / KP), k = (int)(i % KP);
    W[(long long)b * w_stride + i] = (k < K) ? W0[((long long)b * F + f) * K + k] : 0.f;
}
__global__ void __launch_bounds__(kThreads)
unpack_w_kernel(const float* __restrict__ W, long long w_stride, int F, int K, int KP, float* __restrict__ Wout) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)F * K) return;
    const int f = (int)(i / K), k = (int)(i % K);
    Wout[(long long)b * F * K + i] = W[(long long)b * w_stride + (long long)f * KP + k];
}
// H0[b][k][t] -> Ht[b][t][k]  (to_internal) or back; 32x32 tiles; grid = (ceil(T/32), ceil(KP/32), B)
__global__ void __launch_bounds__(kThreads)
transpose_h_kernel(float* __restrict__ Hkt /*[B][K][T]*/, float* __restrict__ Ht, long long h_stride, int T, int K,
                   int KP, int to_internal) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    if (to_internal) {
        for (int r = ly; r < 32; r += 8) {
            const int k = k0 + r, t = t0 + lx;
            tile[r][lx] = (k < K && t < T) ? Hkt[((long long)b * K + k) * T + t] : 0.f;
        }
        __syncthreads();
        for (int r = ly; r < 32; r += 8) {
            const int t = t0 + r, k = k0 + lx;
            if (t < T && k < KP) Ht[(long long)b * h_stride + (long long)t * KP + k] = tile[lx][r];
        }
    } else {
        for (int r = ly; r < 32; r += 8) {
            const int t = t0 + r, k = k0 + lx;
            tile[r][lx] = (t < T && k < K) ? Ht[(long long)b * h_stride + (long long)t * KP + k] : 0.f;
        }
        __syncthreads();
        for (int r = ly; r < 32; r += 8) {
            const int k = k0 + r, t = t0 + lx;
            if (k < K && t < T) Hkt[((long long)b * K + k) * T + t] = tile[lx][r];
        }
    }
}

cudaError_t launch_pack_factors(const float* W0, const float* H0, int B, int F, int T, int K, int KP, float* W,
                                long long w_stride, float* Ht, long long h_stride, cudaStream_t s) {
    AINMF_LAUNCH(pack_w_kernel, dim3((unsigned)ceil_div64((long long)F * KP, kThreads), B), dim3(kThreads), 0, s, W0,
                 F, K, KP, W, w_stride);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    AINMF_LAUNCH(transpose_h_kernel, dim3(ceil_div(T, 32), ceil_div(KP, 32), B), dim3(kThreads), 0, s,
                 const_cast<float*>(H0), Ht, h_stride, T, K, KP, 1);
    return cudaGetLastError();
}

cudaError_t launch_unpack_factors(const float* W, long long w_stride, const float* Ht, long long h_stride, int B,
                                  int F, int T, int K, int KP, float* Wout, float* Hout, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    if (Wout) {
        AINMF_LAUNCH(unpack_w_kernel, dim3((unsigned)ceil_div64((long long)F * K, kThreads), B), dim3(kThreads), 0, s,
                     W, w_stride, F, K, KP, Wout);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (Hout) {
        AINMF_LAUNCH(transpose_h_kernel, dim3(ceil_div(T, 32), ceil_div(KP, 32), B), dim3(kThreads), 0, s, Hout,
                     const_cast<float*>(Ht), h_stride, T, K, KP, 0);
        e = cudaGetLastError();
    }
    return e;
}

__global__ void __launch_bounds__(kThreads)
export_state_kernel(const ClipState* __restrict__ st, int B, int* n_bad, int* n_iter, float* err, int* status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const ClipState s = st[b];
    if (n_bad) n_bad[b] = s.n_bad;
    if (n_iter) n_iter[b] = (s.status == 0) ? s.n_iter : 0;
    if (err) err[b] = (s.status == 2) ? __int_as_float(0x7fc00000) : s.err;       // every frame flagged: NaN, as in the reference
    if (status) status[b] = s.status;
}
cudaError_t launch_export_state(const ClipState* st, int B, int* n_bad, int* n_iter, float* err, int* status,
                                cudaStream_t s) {
    AINMF_LAUNCH(export_state_kernel, dim3(ceil_div(B, kThreads)), dim3(kThreads), 0, s, st, B, n_bad, n_iter, err,
                 status);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kThreads)
reset_state_kernel(ClipState* __restrict__ st, int B, int keep_mean) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ClipState s = st[b];
    if (!keep_mean) { s.n_bad = 0; s.status = 0; s.sum_x = 0.0; s.mean_x = 0.f; }
    s.done = (keep_mean && s.status != 0) ? 1 : 0;
    s.n_iter = 0;
    s.viol_init = 0.0;
    s.viol_last = 0.0;
    s.err = 0.f;
    st[b] = s;
}
cudaError_t launch_reset_state(ClipState* st, int B, cudaStream_t s) {
    AINMF_LAUNCH(reset_state_kernel, dim3(ceil_div(B, kThreads)), dim3(kThreads), 0, s, st, B, 0);
    return cudaGetLastError();
}

// ---- frame permutation for the fit (good frames first) ----------------------------------------------------
// Every bad frame of the imputed spectrogram is the same column (the fill spectrum), so the tensor-core kernels can skip
// their share of the two contractions when the frames are ordered good-first.  perm[b][t'] = original frame.
__global__ void __launch_bounds__(kThreads)
invert_flags_kernel(const unsigned char* __restrict__ bad, long long stride, int T, unsigned char* __restrict__ good) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) good[(long long)b * stride + t] = bad[(long long)b * stride + t] ? 0 : 1;
}
__global__ void __launch_bounds__(kThreads)
build_perm_kernel(const int* __restrict__ good_idx, const int* __restrict__ n_good, const int* __restrict__ bad_idx, int T,
                  int* __restrict__ perm) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int ng = n_good[b];
    perm[(long long)b * T + t] = (t < ng) ? good_idx[(long long)b * T + t] : bad_idx[(long long)b * T + (t - ng)];
}
// dst[b][t'][:] = src[b][perm[t']][:] for t' < limit[b] (limit null: all T rows), zeros beyond; one warp per row
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float* __restrict__ src, long long src_stride, float* __restrict__ dst, long long dst_stride, int ld,
                   const int* __restrict__ perm, int T, const int* __restrict__ limit) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= T) return;
    float4* d = reinterpret_cast<float4*>(dst + (long long)b * dst_stride + (long long)t * ld);
    if (limit && t >= limit[b]) {
        for (int i = lane; i < ld / 4; i += 32) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4* sr = reinterpret_cast<const float4*>(src + (long long)b * src_stride + (long long)perm[(long long)b * T + t] * ld);
    for (int i = lane; i < ld / 4; i += 32) d[i] = sr[i];
}
// dst[b][perm[t']][:] = src[b][t'][:]
__global__ void __launch_bounds__(kThreads)
scatter_rows_kernel(const float* __restrict__ src, long long src_stride, float* __restrict__ dst, long long dst_stride, int ld,
                    const int* __restrict__ perm, int T) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= T) return;
    const float4* sr = reinterpret_cast<const float4*>(src + (long long)b * src_stride + (long long)t * ld);
    float4* d = reinterpret_cast<float4*>(dst + (long long)b * dst_stride + (long long)perm[(long long)b * T + t] * ld);
    for (int i = lane; i < ld / 4; i += 32) d[i] = sr[i];
}
cudaError_t launch_invert_flags(const unsigned char* bad, long long stride, int B, int T, unsigned char* good, cudaStream_t s) {
    AINMF_LAUNCH(invert_flags_kernel, dim3(ceil_div(T, kThreads), B), dim3(kThreads), 0, s, bad, stride, T, good);
    return cudaGetLastError();
}
cudaError_t launch_build_perm(const int* good_idx, const int* n_good, const int* bad_idx, int B, int T, int* perm, cudaStream_t s) {
    AINMF_LAUNCH(build_perm_kernel, dim3(ceil_div(T, kThreads), B), dim3(kThreads), 0, s, good_idx, n_good, bad_idx, T, perm);
    return cudaGetLastError();
}
cudaError_t launch_gather_rows(const float* src, long long src_stride, float* dst, long long dst_stride, int ld, const int* perm,
                               int B, int T, const int* limit, cudaStream_t s) {
    AINMF_LAUNCH(gather_rows_kernel, dim3(ceil_div(T, kThreads / 32), B), dim3(kThreads), 0, s, src, src_stride, dst, dst_stride, ld,
                 perm, T, limit);
    return cudaGetLastError();
}
cudaError_t launch_scatter_rows(const float* src, long long src_stride, float* dst, long long dst_stride, int ld, const int* perm,
                                int B, int T, cudaStream_t s) {
    AINMF_LAUNCH(scatter_rows_kernel, dim3(ceil_div(T, kThreads / 32), B), dim3(kThreads), 0, s, src, src_stride, dst, dst_stride, ld,
                 perm, T);
    return cudaGetLastError();
}

}  // namespace ainmf
