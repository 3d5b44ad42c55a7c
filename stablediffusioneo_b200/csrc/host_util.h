// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <stdio.h>
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

// The same for an UNSWIZZLED box of a bf16 (elem_bytes 2) or fp32 (elem_bytes 4) tensor: the box lands in shared memory
// dense, innermost dimension first (the conv epilogue's residual tile).
int encode_tmap_plain(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box);

// Programmatic dependent launch switch (default on; SDEO_NO_PDL=1 or sdeo_set_pdl(0) turns it off).
bool pdl_enabled();
bool sync_launches();

// Launches `fn` with an optional thread-block cluster and, when enabled, the programmatic-stream-serialization
// attribute. EVERY kernel launched through here must execute griddep_wait() before its first access to memory that an
// earlier kernel wrote or still reads (common.cuh).
template <typename... KArgs, typename... Args>
int launch_k(const char* what, void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster,
             Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster.x;
    attr[na].val.clusterDim.y = cluster.y;
    attr[na].val.clusterDim.z = cluster.z;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fn, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error(-5 /* SDEO_ECUDA */, cudaGetErrorString(e));
  }
  cudaStreamCaptureStatus cap_ = cudaStreamCaptureStatusNone;
  if (sync_launches() && cudaStreamIsCapturing(st, &cap_) == cudaSuccess && cap_ == cudaStreamCaptureStatusNone) {
    // SDEO_SYNC_LAUNCH=1 (debugging aid): wait for the kernel and report ITS error under its own name
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      char buf[256];
      snprintf(buf, sizeof(buf), "%s: kernel failed: %s (grid %u,%u,%u block %u smem %zu)", what, cudaGetErrorString(e), grid.x, grid.y,
               grid.z, block.x, smem);
      return set_error(-5 /* SDEO_ECUDA */, buf);
    }
  }
  return check_launch(what);
}

}  // namespace sdeo
