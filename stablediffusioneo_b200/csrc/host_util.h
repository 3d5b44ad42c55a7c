// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdeo {

// Records `msg` as this thread's last error and returns `code` (so callers can `return set_error(...)`).
int set_error(int code, const char* msg);
// cudaGetLastError() after a launch -> 0 or SDEO_ECUDA with the message recorded.
int check_launch(const char* what);

// cuTensorMapEncodeTiled for a bf16 tensor with 128-byte swizzle and zero OOB fill. dims/box are innermost-first;
// strides_bytes has rank-1 entries (dims 1..rank-1). The driver entry point is resolved lazily through
// cudaGetDriverEntryPoint so that the library links without libcuda.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides);

}  // namespace sdeo
