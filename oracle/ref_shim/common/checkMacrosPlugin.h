// ORACLE BUILD SHIM — TEST INFRASTRUCTURE ONLY. Stands in for the reference's plugin/common/checkMacrosPlugin.h (which pulls
// in NvInfer.h; TensorRT is not in this image) when the reference's groupNormKernel.cu is compiled stand-alone from
// /root/reference by oracle/Makefile. Only the three macros the kernel file uses are provided.
#pragma once
#include <cstdio>
#include <cstdlib>
#define PLUGIN_ASSERT(cond)                                                              \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      std::fprintf(stderr, "PLUGIN_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                                      \
    }                                                                                    \
  } while (0)
#define PLUGIN_FAIL(msg)                                                               \
  do {                                                                                 \
    std::fprintf(stderr, "PLUGIN_FAIL: %s (%s:%d)\n", msg, __FILE__, __LINE__);        \
    std::abort();                                                                      \
  } while (0)
#define PLUGIN_CUASSERT(status)                                                        \
  do {                                                                                 \
    if ((status) != 0) {                                                               \
      std::fprintf(stderr, "PLUGIN_CUASSERT: CUDA error %d (%s:%d)\n", (int)(status), __FILE__, __LINE__); \
      std::abort();                                                                    \
    }                                                                                  \
  } while (0)
