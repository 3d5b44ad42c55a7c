// Canny edge detector on the GPU, bit-exact with cv2.Canny(img, low, high) for 8-bit 1- or 3-channel images (aperture 3,
// L1 gradient) -- the hint preprocessing of the reference pipeline (annotator/canny/__init__.py:4-6,
// canny2image_torch.py:30-38; the arithmetic lives in OpenCV's canny.cpp, a third-party dependency of the reference:
// requirements.txt opencv-contrib-python 4.3.0.36): 3x3 Sobel with replicated borders -> per pixel the channel with the largest
// |dx| + |dy| (first channel wins ties) -> non-maximum suppression with the fixed-point tan(22.5 deg) test (TG22 = 13573,
// shift 15) against a zero-bordered magnitude map -> hysteresis (8-connected growth of strong edges through weak ones).
// Runs once per image, outside the denoising loop: simple one-pass kernels, no host synchronisation.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"

namespace sdeo {

__device__ __forceinline__ void canny_prologue() {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
}

// Sobel + channel selection: mag[h*w] int32, dxy[h*w] = (dx & 0xffff) | (dy << 16)
__global__ void canny_gradient_kernel(const uint8_t* __restrict__ img, int h, int w, int c, int* __restrict__ mag,
                                      int* __restrict__ dxy) {
  canny_prologue();
  const int total = h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i % w;
    const int ym = max(y - 1, 0), yp = min(y + 1, h - 1), xm = max(x - 1, 0), xp = min(x + 1, w - 1);
    int best = -1, bdx = 0, bdy = 0;
    for (int k = 0; k < c; ++k) {
      auto px = [&](int yy, int xx) { return (int)img[((size_t)yy * w + xx) * c + k]; };
      const int a00 = px(ym, xm), a01 = px(ym, x), a02 = px(ym, xp);
      const int a10 = px(y, xm), a12 = px(y, xp);
      const int a20 = px(yp, xm), a21 = px(yp, x), a22 = px(yp, xp);
      const int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
      const int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
      const int m = abs(dx) + abs(dy);
      if (m > best) { best = m; bdx = dx; bdy = dy; }
    }
    mag[i] = best;
    dxy[i] = (bdx & 0xffff) | (bdy << 16);
  }
}

// Non-maximum suppression + double threshold: map = 2 strong edge, 0 weak candidate, 1 suppressed
__global__ void canny_nms_kernel(const int* __restrict__ mag, const int* __restrict__ dxy, int h, int w, int low, int high,
                                 uint8_t* __restrict__ map) {
  canny_prologue();
  const int total = h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i % w;
    auto M = [&](int yy, int xx) { return (yy < 0 || yy >= h || xx < 0 || xx >= w) ? 0 : mag[yy * w + xx]; };
    const int m = mag[i];
    uint8_t out = 1;
    if (m > low) {
      const int packed = dxy[i];
      const int xs = (int)(short)(packed & 0xffff), ys = packed >> 16;
      const int ax = abs(xs), ay = abs(ys) << 15;
      const int tg22x = ax * 13573;
      bool keep = false;
      if (ay < tg22x) {
        keep = m > M(y, x - 1) && m >= M(y, x + 1);
      } else {
        const int tg67x = tg22x + (ax << 16);
        if (ay > tg67x) {
          keep = m > M(y - 1, x) && m >= M(y + 1, x);
        } else {
          const int s = (xs ^ ys) < 0 ? -1 : 1;
          keep = m > M(y - 1, x - s) && m > M(y + 1, x + s);
        }
      }
      if (keep) out = m > high ? 2 : 0;
    }
    map[i] = out;
  }
}

// Hysteresis: weak candidates (0) that touch a strong edge (2) become strong, repeated to the fixed point. One block
// sweeps the whole map (the closure is unique, so the sweep order does not matter), then writes 255 / 0.
__global__ void __launch_bounds__(1024) canny_hysteresis_kernel(uint8_t* map, int h, int w, uint8_t* __restrict__ edges) {
  canny_prologue();
  __shared__ int changed;
  const int total = h * w;
  volatile uint8_t* vm = map;
  for (;;) {
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      if (vm[i] != 0) continue;
      const int y = i / w, x = i % w;
      bool hit = false;
      for (int dy = -1; dy <= 1 && !hit; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= w || (dx == 0 && dy == 0)) continue;
          if (vm[yy * w + xx] == 2) { hit = true; break; }
        }
      }
      if (hit) { vm[i] = 2; local = 1; }
    }
    if (local) changed = 1;  // benign race: every writer stores 1
    __syncthreads();
    const int again = changed;
    __syncthreads();
    if (!again) break;
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) edges[i] = vm[i] == 2 ? 255 : 0;
}

// detected_map (uint8 [h*w], 0 / 255) -> control hint fp32 NCHW [n, 3, h, w] = HWC3(map) / 255 stacked n times
// (canny2image_torch.py:34-38)
__global__ void edges_to_hint_kernel(const uint8_t* __restrict__ edges, float* __restrict__ hint, int n, int hw) {
  canny_prologue();
  const long long total = (long long)n * 3 * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    hint[i] = (float)edges[i % hw] / 255.0f;
}

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_canny)

extern "C" size_t sdeo_canny_workspace_bytes(int32_t h, int32_t w) {
  if (h <= 0 || w <= 0) return 0;
  return (size_t)h * w * (2 * sizeof(int) + 1) + 256;
}

extern "C" int sdeo_canny_u8(const uint8_t* img, int32_t h, int32_t w, int32_t c, double low_threshold, double high_threshold,
                             uint8_t* edges, void* workspace, size_t workspace_bytes, void* stream) {
  if (!img || !edges || !workspace || h <= 0 || w <= 0 || (c != 1 && c != 3) || (long long)h * w > (1 << 28))
    return set_error(SDEO_EINVAL, "canny_u8: bad args (8-bit image with 1 or 3 channels)");
  if (workspace_bytes < sdeo_canny_workspace_bytes(h, w)) return set_error(SDEO_EINVAL, "canny_u8: workspace too small");
  if (low_threshold > high_threshold) { const double t = low_threshold; low_threshold = high_threshold; high_threshold = t; }
  const int low = (int)floor(low_threshold), high = (int)floor(high_threshold);
  int* mag = reinterpret_cast<int*>(workspace);
  int* dxy = mag + (size_t)h * w;
  uint8_t* map = reinterpret_cast<uint8_t*>(dxy + (size_t)h * w);
  cudaStream_t st = (cudaStream_t)stream;
  const int total = h * w;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  int rc = launch_k("canny_gradient", canny_gradient_kernel, dim3(blocks), dim3(256), 0, st, dim3(1, 1, 1), img, h, w, c, mag, dxy);
  if (rc) return rc;
  rc = launch_k("canny_nms", canny_nms_kernel, dim3(blocks), dim3(256), 0, st, dim3(1, 1, 1), (const int*)mag, (const int*)dxy, h,
                w, low, high, map);
  if (rc) return rc;
  return launch_k("canny_hysteresis", canny_hysteresis_kernel, dim3(1), dim3(1024), 0, st, dim3(1, 1, 1), map, h, w, edges);
}

extern "C" int sdeo_edges_to_hint(const uint8_t* edges, float* hint, int32_t n, int32_t hw, void* stream) {
  if (!edges || !hint || n <= 0 || hw <= 0) return set_error(SDEO_EINVAL, "edges_to_hint: bad args");
  long long blocks = ((long long)n * 3 * hw + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_k("edges_to_hint", edges_to_hint_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1),
                  edges, hint, n, hw);
}
