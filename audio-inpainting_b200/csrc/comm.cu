// comm.cu -- time-frame sharding of one long signal across the GPUs of a box.
#include <string.h>

#include "../../include/ainmf.h"
#include "kernels.h"

extern "C" {

int ainmf_comm_destroy_internal(ainmf_handle) { return AINMF_OK; }

}  // extern "C"

extern "C" {

int ainmf_comm_unique_id(uint8_t*) { return AINMF_ERR_COMM; }
int ainmf_comm_init(ainmf_handle, const uint8_t*, int32_t, int32_t) { return AINMF_ERR_COMM; }
int ainmf_shard_plan(int64_t, int32_t, int32_t, int32_t, int32_t, int32_t*, int32_t*, int64_t*, int64_t*, int64_t*,
                     int64_t*) { return AINMF_ERR_COMM; }
size_t ainmf_sharded_workspace_bytes(ainmf_handle, const ainmf_params*) { return 0; }
int ainmf_inpaint_sharded(ainmf_handle, const ainmf_params*, const float*, float*, int32_t*, float*, float*, float*,
                          int32_t*, void*, size_t, void*) { return AINMF_ERR_COMM; }

}  // extern "C"
