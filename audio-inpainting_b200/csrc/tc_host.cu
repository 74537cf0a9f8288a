// tc_host.cu -- host helpers of the tensor-core path: TMA tensor-map construction via the driver entry point.
#include "tc.cuh"

#ifndef AINMF_EMU
namespace ainmf {

int make_tensor_map_3d(CUtensorMap* out, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1,
                       uint64_t pitch2, uint32_t b0, uint32_t b1, int atom32) {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
        fn = (PFN_tensorMapEncodeTiled)p;
    }
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {pitch1 * sizeof(float), pitch2 * sizeof(float)};   // bytes, dims 1 and 2
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace ainmf
#endif
