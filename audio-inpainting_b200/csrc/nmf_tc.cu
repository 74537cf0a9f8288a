// nmf_tc.cu -- host side and helper kernels of the tensor-core path; the two V-sized contractions of the CD-NMF
// iteration (tcgen05, operands through TMEM) live in nmf_ts.cu.
#include "kernels.h"
#include "nmf_cd.cuh"
#include <stdlib.h>

#include "tc.cuh"

#ifndef AINMF_EMU
namespace ainmf {
using namespace tc;

constexpr int TC_BK = 32;          // contraction elements per stage (one 128-byte row)
constexpr int TC_M = 128;          // MMA M

// out[b][e] = sum_s partial[b][s][e]  (fixed order); grid = (ceil(n4/256), B)
__global__ void __launch_bounds__(kThreads)
reduce_splits_kernel(const float* __restrict__ partial, int S, long long n4, float* __restrict__ out, long long out_stride,
                     const ClipState* __restrict__ st) {
    const int b = blockIdx.y;
    if (st[b].done) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(partial + (((long long)b * S + s) * n4 + i) * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (long long)b * out_stride + i * 4) = acc;
}

// ---- host ----------------------------------------------------------------------------------------------------
static_assert(sizeof(CUtensorMap) == sizeof(TcMapBlob), "CUtensorMap is 128 bytes");
static inline const CUtensorMap& as_map(const TcMapBlob& b) { return *reinterpret_cast<const CUtensorMap*>(&b); }
static inline CUtensorMap* as_map_ptr(TcMapBlob* b) { return reinterpret_cast<CUtensorMap*>(b); }
// W half-step, local part: X.Ht partials and the Gram of Ht from the persistent TMEM-operand kernel, then the ordered
// sum of the Gram partials
cudaError_t nmf_tc_half1(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) {
    cudaError_t e = nmf_ts_half1(p, wk, s);
    if (e != cudaSuccess) return e;
    const long long n4 = (long long)p.KP * p.KP / 4;
    AINMF_LAUNCH(reduce_splits_kernel, dim3((unsigned)ceil_div64(n4, kThreads), p.B), dim3(kThreads), 0, s, wk.gram_partial,
                 wk.tc_splits, n4, wk.HHt, (long long)p.KP * p.KP, p.state);
    return cudaGetLastError();
}

// H half-step: the persistent TMEM-operand kernel (nmf_ts.cu); its W^T operands and the operands derived from W^T W are
// written by the W-side kernel (nmf_cd.cu: w_side_kernel)
cudaError_t nmf_tc_hstep(const NmfProblem& p, const NmfWork& wk, cudaStream_t s) { return nmf_ts_hstep(p, wk, s); }

int nmf_tc_setup(const NmfProblem& p, NmfWork* wk, TcMaps* m) {
    if (!wk->use_tc) return 0;
    const uint64_t B = p.B, T = p.T, F = p.F, ldf = p.ldf, KP = p.KP;
    int rc = make_tensor_map_3d(as_map_ptr(&m->mapX), p.Xt, F, T, B, ldf, (uint64_t)p.x_stride, TC_BK, TC_M, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapWt), wk->tc_Wt, F, KP, B, ldf, KP * ldf, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapWtLo), wk->tc_WtLo, F, KP, B, ldf, KP * ldf, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapXs), p.Xt, F, T, B, ldf, (uint64_t)p.x_stride, 32, TC_BK, 2);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapHs), p.Ht, KP, T, B, KP, (uint64_t)p.h_stride, 32, TC_BK, 2);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapHmn), p.Ht, KP, T, B, KP, (uint64_t)p.h_stride, 32, TC_BK, 1);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapHk), p.Ht, KP, T, B, KP, (uint64_t)p.h_stride, TC_BK, TC_M, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapG), wk->WtW, KP, KP, B, KP, KP * KP, TC_BK, (uint32_t)KP, 0);
    if (!rc) rc = make_tensor_map_3d(as_map_ptr(&m->mapGlo), wk->tc_GLo, KP, KP, B, KP, KP * KP, TC_BK, (uint32_t)KP, 0);
    wk->tc = m;
    return rc;
}

}  // namespace ainmf
#endif  // AINMF_EMU
