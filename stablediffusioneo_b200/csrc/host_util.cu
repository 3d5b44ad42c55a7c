#include "host_util.h"
#include "../../include/sdeo.h"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace sdeo {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return SDEO_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return SDEO_ECUDA;
}

bool sync_launches() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SDEO_SYNC_LAUNCH");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("SDEO_NO_PDL");
    g_pdl = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else (void)cudaGetLastError();
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(SDEO_ENOSYS, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(SDEO_EINVAL, "tensor map: base not 16-byte aligned");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = elem_strides[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides_bytes[i];
    if (s[i] % 16 != 0) return set_error(SDEO_EINVAL, "tensor map: stride not a multiple of 16 bytes");
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r,
             rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), b[0], rank > 1 ? b[1] : 0);
    return set_error(SDEO_ECUDA, msg);
  }
  return SDEO_OK;
}

int encode_tmap_plain(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(SDEO_ENOSYS, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(SDEO_EINVAL, "tensor map: base not 16-byte aligned");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides_bytes[i];
    if (s[i] % 16 != 0) return set_error(SDEO_EINVAL, "tensor map: stride not a multiple of 16 bytes");
  }
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (plain) failed: CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r,
             rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), b[0], rank > 1 ? b[1] : 0);
    return set_error(SDEO_ECUDA, msg);
  }
  return SDEO_OK;
}

}  // namespace sdeo

extern "C" const char* sdeo_last_error(void) { return sdeo::g_err; }
extern "C" int sdeo_version(void) { return 2; }  // 2: sdeo_conv_args::pad_hi, fp32-mode and sampler-mode entry points
extern "C" int sdeo_trace_set_conv(void*);
extern "C" int sdeo_trace_set_attention(void*);
extern "C" int sdeo_trace_set_norm(void*);
extern "C" int sdeo_trace_set_elementwise(void*);
extern "C" int sdeo_trace_set_precise(void*);
extern "C" int sdeo_trace_set_canny(void*);
extern "C" int sdeo_set_trace(void* buf) {
  int rc = sdeo_trace_set_conv(buf) | sdeo_trace_set_attention(buf) | sdeo_trace_set_norm(buf) | sdeo_trace_set_elementwise(buf) |
           sdeo_trace_set_precise(buf) | sdeo_trace_set_canny(buf);
  return rc ? sdeo::set_error(SDEO_ECUDA, "set_trace: cudaMemcpyToSymbol failed") : SDEO_OK;
}
extern "C" int sdeo_set_pdl(int enable) {
  sdeo::set_pdl(enable);
  return SDEO_OK;
}
extern "C" int sdeo_memset_async(void* p, int value, size_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return sdeo::set_error(SDEO_ECUDA, cudaGetErrorString(e));
  return SDEO_OK;
}
