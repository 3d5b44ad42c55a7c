// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C driver around the REFERENCE's own GroupNorm kernels (plugin/groupNormPlugin/groupNormKernel.cu, compiled from
// /root/reference where it lies -- see oracle/Makefile): restates the parameter set-up of GroupNormPlugin::enqueue
// (plugin/groupNormPlugin/groupNormPlugin.cpp:179-228, findMaxDivisor :36-57) without TensorRT types, so that tests and
// tools can run the reference kernels on raw device pointers: the numerical pin of the plugin contract (fp16 NHWC in/out,
// fp32 gamma/beta, optional Swish) and the native "kernel to beat" for sdeo_groupnorm_nhwc_f16.
#include "groupNormKernel.h"

#include <cmath>
#include <cstdint>

static int32_t find_max_divisor(int32_t n, int32_t max_allowed) {  // groupNormPlugin.cpp:36-57
  int32_t best = -1;
  for (int32_t i = 1; i <= std::sqrt(n); i++) {
    if (n % i == 0) {
      const int32_t d1 = n / i, d2 = i;
      if (d1 > best && d1 < max_allowed) best = d1;
      if (d2 > best && d2 < max_allowed) best = d2;
    }
  }
  return best;
}

extern "C" size_t ref_groupnorm_workspace_bytes(void) { return (sizeof(float) * 2) * 32 * 32; }  // groupNormPlugin.cpp:173-177

// Shapes the reference kernels accept: c / cPerBlock integral with the plugin's cPerBlock table, hw % hwPerBlock == 0, n <= 32.
extern "C" int ref_groupnorm_enqueue(const void* x, const float* gamma, const float* beta, void* y, int32_t n, int32_t c,
                                     int32_t h, int32_t w, int32_t with_swish, void* workspace, void* stream) {
  int32_t cPerBlock = 320;
  switch (c) {
    case 960:
    case 1920: cPerBlock = 480; break;
    case 512:
    case 256: cPerBlock = 256; break;
    case 128: cPerBlock = 128; break;
    default: cPerBlock = 320;
  }
  if (n <= 0 || n > 32 || c % 32 != 0 || c % cPerBlock != 0) return -1;
  GroupNormNHWCParams p{};
  p.withSwish = with_swish != 0;
  p.dst = static_cast<half*>(y);
  p.src = static_cast<half const*>(x);
  p.gamma = gamma;
  p.beta = beta;
  p.redBuffer = static_cast<float*>(workspace);
  p.n = n;
  p.h = h;
  p.w = w;
  p.c = c;
  p.groups = 32;
  p.hw = h * w;
  const int32_t blocksPerHW = find_max_divisor(p.hw, 1024);
  p.hwPerBlock = divUp(p.hw, blocksPerHW);
  p.cPerBlock = cPerBlock;
  p.cPerGroup = c / p.groups;
  p.hwc = p.hw * c;
  p.invHWC = 1.F / (float)(p.hw * p.cPerGroup);
  p.groupsPerBlock = cPerBlock / p.cPerGroup;
  if (p.hw % p.hwPerBlock != 0 || cPerBlock % p.cPerGroup != 0) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(p.redBuffer, 0, ref_groupnorm_workspace_bytes(), st);
  groupNormNHWCSum(p, st);
  groupNormNHWCScale(p, st);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
