// Elementwise / layout passes of the denoising loop: CFG + DDIM update, layout conversion at the NCHW fp32
// boundary, nearest upsample, residual injection, timestep embedding, row softmax, image quantisation.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"

namespace sdeo {

// ---- CFG combine + DDIM update (+ next-step network input) --------------------------------------
__global__ void cfg_ddim_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u, int eps_nhwc,
                                int ld_eps, const float* __restrict__ x, const float* __restrict__ noise,
                                float* __restrict__ x_prev, float* __restrict__ pred_x0,
                                __nv_bfloat16* __restrict__ x_next, int dup, int ldn,
                                const float* __restrict__ coef_table, const int* __restrict__ step_idx, int n, int c,
                                int hw, int noise_table) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const int row = step_idx ? *step_idx : 0;
  const float* cf = coef_table + (size_t)row * 8;
  const float s = cf[0], sqrt_1m_at = cf[1], rsqrt_at = cf[2], sqrt_aprev = cf[3], dir_coef = cf[4], sigma = cf[5];
  const long long total = (long long)n * hw;
  // eta > 0: `noise` is this step's tensor, or (noise_table) a table [steps][n*c*hw] indexed by the device step counter
  // (the captured step graph of the sampler engine); rows with sigma == 0 are never read
  const float* noise_row = (noise && sigma != 0.f) ? noise + (noise_table ? (size_t)row * (size_t)(total * c) : 0) : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int p = (int)(i % hw);
    for (int ch = 0; ch < c; ++ch) {
      const long long nchw = ((long long)b * c + ch) * hw + p;
      const long long ei = eps_nhwc ? (i * ld_eps + ch) : nchw;
      const float ec = eps_c[ei];
      float e = ec;
      if (eps_u) {
        const float eu = eps_u[ei];
        e = eu + s * (ec - eu);
      }
      const float xv = x[nchw];
      const float x0 = (xv - sqrt_1m_at * e) * rsqrt_at;
      float xp = sqrt_aprev * x0 + dir_coef * e;
      if (noise_row) xp += sigma * noise_row[nchw];
      x_prev[nchw] = xp;
      if (pred_x0) pred_x0[nchw] = x0;
      if (x_next) {
        for (int d = 0; d < dup; ++d) x_next[((long long)d * total + i) * ldn + ch] = __float2bfloat16(xp);
      }
    }
    if (x_next) {
      for (int d = 0; d < dup; ++d)
        for (int ch = c; ch < ldn; ++ch) x_next[((long long)d * total + i) * ldn + ch] = __float2bfloat16(0.f);
    }
  }
}

__global__ void counter_add_kernel(int* ctr, int delta) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  *ctr += delta;
}

// ---- layout conversion ---------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 (zero-padded to ldy channels). One thread per (pixel, channel-slot).
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int c,
                                         int hw, int ldy, float scale) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  // tile transpose through shared memory: 32 channels x 32 pixels
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && p < hw) ? x[((long long)b * c + ch) * hw + p] * scale : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, ch = c0 + threadIdx.x;
    if (p < hw && ch < ldy) y[((long long)b * hw + p) * ldy + ch] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}

template <typename TIn>
__global__ void nhwc_to_nchw_f32_kernel(const TIn* __restrict__ x, float* __restrict__ y, int n, int c, int hw,
                                        int ldx) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, ch = c0 + threadIdx.x;
    float v = 0.f;
    if (p < hw && ch < c) v = (float)x[((long long)b * hw + p) * ldx + ch];
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j, p = p0 + threadIdx.x;
    if (ch < c && p < hw) y[((long long)b * c + ch) * hw + p] = tile[threadIdx.x][j];
  }
}

// ---- nearest x2 upsample, NHWC, 16-byte vectors ---------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int cv) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const long long total = (long long)n * (2 * h) * (2 * w) * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int ox = (int)(t % (2 * w));
    t /= (2 * w);
    const int oy = (int)(t % (2 * h));
    const int b = (int)(t / (2 * h));
    y[i] = x[(((long long)b * h + (oy >> 1)) * w + (ox >> 1)) * cv + v];
  }
}

// ---- y = a + alpha * b (bf16) ---------------------------------------------------------------------
__global__ void add_scaled_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, float alpha,
                                  uint4* __restrict__ y, long long nvec, const __nv_bfloat16* a_s,
                                  const __nv_bfloat16* b_s, __nv_bfloat16* y_s, long long count) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 av = a[i], bv = b[i];
    float2 a0 = unpack_bf16x2(av.x), a1 = unpack_bf16x2(av.y), a2 = unpack_bf16x2(av.z), a3 = unpack_bf16x2(av.w);
    float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y), b2 = unpack_bf16x2(bv.z), b3 = unpack_bf16x2(bv.w);
    uint4 o;
    o.x = pack_bf16x2(a0.x + alpha * b0.x, a0.y + alpha * b0.y);
    o.y = pack_bf16x2(a1.x + alpha * b1.x, a1.y + alpha * b1.y);
    o.z = pack_bf16x2(a2.x + alpha * b2.x, a2.y + alpha * b2.y);
    o.w = pack_bf16x2(a3.x + alpha * b3.x, a3.y + alpha * b3.y);
    y[i] = o;
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (long long i = nvec * 8 + threadIdx.x; i < count; i += blockDim.x)
      y_s[i] = __float2bfloat16(__bfloat162float(a_s[i]) + alpha * __bfloat162float(b_s[i]));
  }
}

// ---- timestep embedding: [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / half) ------------
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, const int* __restrict__ step_idx,
                                          __nv_bfloat16* __restrict__ y, int n, int dim, int ldy, float max_period) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ldy) return;
  const int b = i / ldy, col = i % ldy;
  const float tv = (float)(step_idx ? t[*step_idx] : t[b]);
  float v = 0.f;
  if (col < 2 * half) {
    const int k = col < half ? col : col - half;
    const float freq = expf(-logf(max_period) * (float)k / (float)half);
    const float arg = tv * freq;
    v = col < half ? cosf(arg) : sinf(arg);
  }
  y[i] = __float2bfloat16(v);
}

__global__ void silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long count) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(silu_f(__bfloat162float(x[i])));
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long count) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long count) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

// ---- row softmax (VAE AttnBlock): one CTA per row, fp32 math -----------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int cols, int ld, int ldy,
                    float scale_log2) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  __shared__ float red[8];
  __shared__ float bcast;
  const float* xr = x + (size_t)blockIdx.x * ld;
  __nv_bfloat16* yr = y + (size_t)blockIdx.x * ldy;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) m = fmaxf(m, xr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s += exp2f((xr[i] - m) * scale_log2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int i = threadIdx.x; i < cols; i += blockDim.x)
    yr[i] = __float2bfloat16(exp2f((xr[i] - m) * scale_log2) * inv);
}

// Same softmax with the row held in registers: ONE read of the fp32 scores (the three-pass kernel above re-reads the row
// for the sum and for the output and evaluates every exponential twice). 256 threads x kV float4 vectors: cols <= 1024 * kV,
// cols % 4 == 0, 16-byte aligned rows. The VAE mid-block attention has cols = 1536 (256x384 image) ... 9216 (768x768).
template <int kV>
__global__ void __launch_bounds__(256)
softmax_rows_reg_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int cols, int ld, int ldy, float scale_log2) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float red[8];
  __shared__ float red2[8];
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)blockIdx.x * ld);
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)blockIdx.x * ldy);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = cols >> 2;
  float4 v[kV];
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < kV; ++u) {
    const int i = threadIdx.x + u * 256;
    v[u] = i < nv ? __ldg(xr + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
#pragma unroll
  for (int u = 0; u < kV; ++u) m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kV; ++u) {   // (padding lanes hold -inf: exp2 gives 0)
    v[u].x = exp2f((v[u].x - m) * scale_log2); v[u].y = exp2f((v[u].y - m) * scale_log2);
    v[u].z = exp2f((v[u].z - m) * scale_log2); v[u].w = exp2f((v[u].w - m) * scale_log2);
    s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red2[warp] = s;
  __syncthreads();
  s = red2[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) s += red2[i];
  const float inv = 1.0f / s;
#pragma unroll
  for (int u = 0; u < kV; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < nv) yr[i] = make_uint2(pack_bf16x2(v[u].x * inv, v[u].y * inv), pack_bf16x2(v[u].z * inv, v[u].w * inv));
  }
}

__global__ void image_to_u8_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ y, long long npix, int c,
                                   int ldx) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const long long total = npix * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i % c);
    float v = __bfloat162float(x[p * ldx + ch]) * 127.5f + 127.5f;
    v = fminf(fmaxf(v, 0.f), 255.f);
    y[i] = (uint8_t)v;  // truncation, like numpy astype(np.uint8) on a clipped float (canny2image_torch.py:68)
  }
}

// ---- token + position embedding lookup (CLIP text encoder input): y = tok[ids[r]] + pos[r % T], fp32 + bf16 twin ----
__global__ void embedding_add_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                     const float* __restrict__ pos, float* __restrict__ y, __nv_bfloat16* __restrict__ y2,
                                     int rows, int T, int C, int vocab) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const long long total = (long long)rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float v = tok[id * C + c] + pos[(long long)(r % T) * C + c];
    y[i] = v;
    if (y2) y2[i] = __float2bfloat16(v);
  }
}

// ---- fp32 linear combinations of latents (sampler modes beyond plain sampling) ---------------------
// y[i] = a[i / per] * x[i] + b[i / per] * z[i]: per-sample device coefficients. DDIM encode step
// (ddim_hacked.py:262-265), stochastic_encode / q_sample (:290-292), CFG combine outside the fused step kernel.
__global__ void axpby_f32_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ a,
                                 const float* __restrict__ b, float* __restrict__ y, long long count, long long per) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per;
    y[i] = a[s] * x[i] + b[s] * z[i];
  }
}
// Inpainting blend (ddim_hacked.py:154-157): y = mask * (a x0 + b noise) + (1 - mask) * img, the parenthesis being
// q_sample(x0, t). mask has mask_c = 1 or c channels per sample ([n, mask_c, hw]); img / x0 / noise are [n, c, hw].
__global__ void mask_blend_f32_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                      const float* __restrict__ img, const float* __restrict__ mask,
                                      const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                      int n, int c, int mask_c, long long hw) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const long long count = (long long)n * c * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / (c * hw);
    const int ch = (int)((i / hw) % c);
    const long long p = i % hw;
    const float m = mask[(s * mask_c + (mask_c == 1 ? 0 : ch)) * hw + p];
    const float orig = a[s] * x0[i] + b[s] * noise[i];
    y[i] = orig * m + (1.0f - m) * img[i];
  }
}

// The same blend inside the captured step graph: img_orig = q_sample(x0, t_step) of EVERY step is drawn up front into a
// table [S][n, c, hw]; the row is selected on the device by the step counter. y may alias img.
__global__ void mask_blend_table_f32_kernel(const float* __restrict__ orig_table, const float* img,
                                            const float* __restrict__ mask, float* y, const int* __restrict__ step_idx,
                                            int n, int c, int mask_c, long long hw) {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const long long count = (long long)n * c * hw;
  const float* orig = orig_table + (long long)__ldg(step_idx) * count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / (c * hw);
    const int ch = (int)((i / hw) % c);
    const long long p = i % hw;
    const float m = mask[(s * mask_c + (mask_c == 1 ? 0 : ch)) * hw + p];
    y[i] = orig[i] * m + (1.0f - m) * img[i];
  }
}

static int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_elementwise)

extern "C" int sdeo_cfg_ddim_step(const float* eps_c, const float* eps_u, int32_t eps_nhwc, int32_t ld_eps, const float* x,
                                  const float* noise, float* x_prev, float* pred_x0, void* x_next, int32_t dup, int32_t ldn,
                                  const float* coef_table, const int32_t* step_idx, int32_t n, int32_t c, int32_t hw,
                                  void* stream) {
  if (!eps_c || !x || !x_prev || !coef_table || n <= 0 || c <= 0 || hw <= 0) return set_error(SDEO_EINVAL, "cfg_ddim_step: bad args");
  if (x_next && (dup <= 0 || ldn < c)) return set_error(SDEO_EINVAL, "cfg_ddim_step: bad x_next geometry");
  if (eps_nhwc && ld_eps < c) return set_error(SDEO_EINVAL, "cfg_ddim_step: ld_eps < c");
  const long long total = (long long)n * hw;
  return launch_k("cfg_ddim_step", cfg_ddim_kernel, dim3(grid_for(total, 128)), dim3(128), 0, (cudaStream_t)stream, dim3(1, 1, 1), eps_c, eps_u, eps_nhwc, ld_eps, x, noise, x_prev, pred_x0, (__nv_bfloat16*)x_next, dup, ldn, coef_table, step_idx, n, c, hw, 0);
}

extern "C" int sdeo_cfg_ddim_step_noise_table(const float* eps_c, const float* eps_u, int32_t eps_nhwc, int32_t ld_eps,
                                              const float* x, const float* noise_table, float* x_prev, float* pred_x0,
                                              void* x_next, int32_t dup, int32_t ldn, const float* coef_table,
                                              const int32_t* step_idx, int32_t n, int32_t c, int32_t hw, void* stream) {
  if (!eps_c || !x || !x_prev || !coef_table || !noise_table || !step_idx || n <= 0 || c <= 0 || hw <= 0)
    return set_error(SDEO_EINVAL, "cfg_ddim_step_noise_table: bad args");
  if (x_next && (dup <= 0 || ldn < c)) return set_error(SDEO_EINVAL, "cfg_ddim_step_noise_table: bad x_next geometry");
  if (eps_nhwc && ld_eps < c) return set_error(SDEO_EINVAL, "cfg_ddim_step_noise_table: ld_eps < c");
  const long long total = (long long)n * hw;
  return launch_k("cfg_ddim_step", cfg_ddim_kernel, dim3(grid_for(total, 128)), dim3(128), 0, (cudaStream_t)stream, dim3(1, 1, 1), eps_c, eps_u, eps_nhwc, ld_eps, x, noise_table, x_prev, pred_x0, (__nv_bfloat16*)x_next, dup, ldn, coef_table, step_idx, n, c, hw, 1);
}

extern "C" int sdeo_counter_add(int32_t* ctr, int32_t delta, void* stream) {
  if (!ctr) return set_error(SDEO_EINVAL, "counter_add: null");
  return launch_k("counter_add", counter_add_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, dim3(1, 1, 1), ctr, delta);
}

extern "C" int sdeo_nchw_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t hw, int32_t ldy, float scale,
                                      void* stream) {
  if (!x || !y || n <= 0 || c <= 0 || hw <= 0 || ldy < c || n > 65535) return set_error(SDEO_EINVAL, "nchw_to_nhwc: bad args");
  dim3 grid((hw + 31) / 32, (ldy + 31) / 32, n), block(32, 8);
  return launch_k("nchw_to_nhwc", nchw_to_nhwc_bf16_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, (__nv_bfloat16*)y, n, c, hw, ldy, scale);
}
extern "C" int sdeo_nhwc_bf16_to_nchw(const void* x, float* y, int32_t n, int32_t c, int32_t hw, int32_t ldx, void* stream) {
  if (!x || !y || n <= 0 || c <= 0 || hw <= 0 || ldx < c || n > 65535) return set_error(SDEO_EINVAL, "nhwc_to_nchw: bad args");
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  return launch_k("nhwc_to_nchw", nhwc_to_nchw_f32_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const __nv_bfloat16*)x, y, n, c, hw, ldx);
}
extern "C" int sdeo_nhwc_f32_to_nchw(const float* x, float* y, int32_t n, int32_t c, int32_t hw, int32_t ldx, void* stream) {
  if (!x || !y || n <= 0 || c <= 0 || hw <= 0 || ldx < c || n > 65535) return set_error(SDEO_EINVAL, "nhwc_f32_to_nchw: bad args");
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  return launch_k("nhwc_f32_to_nchw", nhwc_to_nchw_f32_kernel<float>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, y, n, c, hw, ldx);
}

extern "C" int sdeo_upsample_nearest2x(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
  if (!x || !y || n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 != 0) return set_error(SDEO_EINVAL, "upsample2x: bad args");
  const long long total = (long long)n * 4 * h * w * (c / 8);
  return launch_k("upsample2x", upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const uint4*)x, (uint4*)y, n, h, w, c / 8);
}

extern "C" int sdeo_add_scaled(const void* a, const void* b, float alpha, void* y, int64_t count, void* stream) {
  if (!a || !b || !y || count <= 0) return set_error(SDEO_EINVAL, "add_scaled: bad args");
  const bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0;
  const long long nvec = aligned ? count / 8 : 0;
  return launch_k("add_scaled", add_scaled_kernel, dim3(grid_for(nvec > 0 ? nvec : 1, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const uint4*)a, (const uint4*)b, alpha, (uint4*)y, nvec, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, count);
}

extern "C" int sdeo_timestep_embedding(const int64_t* t, const int32_t* step_idx, void* y, int32_t n, int32_t dim,
                                       int32_t ldy, float max_period, void* stream) {
  if (!t || !y || n <= 0 || dim <= 0 || ldy < dim) return set_error(SDEO_EINVAL, "timestep_embedding: bad args");
  const int total = n * ldy;
  return launch_k("timestep_embedding", timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const long long*)t, step_idx, (__nv_bfloat16*)y, n, dim, ldy, max_period);
}

extern "C" int sdeo_embedding_add(const int64_t* ids, const float* tok, const float* pos, float* y, void* y2, int32_t rows,
                                  int32_t t, int32_t c, int32_t vocab, void* stream) {
  if (!ids || !tok || !pos || !y || rows <= 0 || t <= 0 || c <= 0 || vocab <= 0) return set_error(SDEO_EINVAL, "embedding_add: bad args");
  return launch_k("embedding_add", embedding_add_kernel, dim3(grid_for((long long)rows * c, 256)), dim3(256), 0,
                  (cudaStream_t)stream, dim3(1, 1, 1), (const long long*)ids, tok, pos, y, (__nv_bfloat16*)y2, rows, t, c, vocab);
}

extern "C" int sdeo_silu(const void* x, void* y, int64_t count, void* stream) {
  if (!x || !y || count <= 0) return set_error(SDEO_EINVAL, "silu: bad args");
  return launch_k("silu", silu_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const __nv_bfloat16*)x, (__nv_bfloat16*)y, count);
}
extern "C" int sdeo_f32_to_bf16(const float* x, void* y, int64_t count, void* stream) {
  if (!x || !y || count <= 0) return set_error(SDEO_EINVAL, "f32_to_bf16: bad args");
  return launch_k("f32_to_bf16", f32_to_bf16_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, (__nv_bfloat16*)y, count);
}
extern "C" int sdeo_bf16_to_f32(const void* x, float* y, int64_t count, void* stream) {
  if (!x || !y || count <= 0) return set_error(SDEO_EINVAL, "bf16_to_f32: bad args");
  return launch_k("bf16_to_f32", bf16_to_f32_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const __nv_bfloat16*)x, y, count);
}

extern "C" int sdeo_softmax_rows(const float* x, void* y, int32_t rows, int32_t cols, int32_t ldx, int32_t ldy, float scale,
                                 void* stream) {
  if (!x || !y || rows <= 0 || cols <= 0 || ldx < cols || ldy < cols) return set_error(SDEO_EINVAL, "softmax_rows: bad args");
  const float sl2 = scale * 1.4426950408889634f;
  // register-resident row (one read) when the row is vector-aligned and fits 256 threads x 12 float4
  if (cols % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0 && cols <= 12288 &&
      !getenv("SDEO_SOFTMAX_THREE_PASS")) {
    auto fn = cols <= 2048 ? softmax_rows_reg_kernel<2> : (cols <= 4096 ? softmax_rows_reg_kernel<4> : (cols <= 8192 ? softmax_rows_reg_kernel<8> : softmax_rows_reg_kernel<12>));
    return launch_k("softmax_rows", fn, dim3(rows), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, (__nv_bfloat16*)y, cols, ldx, ldy, sl2);
  }
  return launch_k("softmax_rows", softmax_rows_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, (__nv_bfloat16*)y, cols, ldx, ldy, sl2);
}

extern "C" int sdeo_image_to_u8(const void* x, uint8_t* y, int32_t npix, int32_t c, int32_t ldx, void* stream) {
  if (!x || !y || npix <= 0 || c <= 0 || ldx < c) return set_error(SDEO_EINVAL, "image_to_u8: bad args");
  return launch_k("image_to_u8", image_to_u8_kernel, dim3(grid_for((long long)npix * c, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), (const __nv_bfloat16*)x, y, npix, c, ldx);
}

extern "C" int sdeo_axpby_f32(const float* x, const float* z, const float* a, const float* b, float* y, int64_t count,
                              int64_t per_sample, void* stream) {
  if (!x || !z || !a || !b || !y || count <= 0 || per_sample <= 0) return set_error(SDEO_EINVAL, "axpby_f32: bad args");
  return launch_k("axpby_f32", axpby_f32_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x, z, a, b, y, (long long)count, (long long)per_sample);
}
extern "C" int sdeo_mask_blend_table_f32(const float* orig_table, const float* img, const float* mask, float* y,
                                         const int32_t* step_idx, int32_t n, int32_t c, int32_t mask_c, int64_t hw, void* stream) {
  if (!orig_table || !img || !mask || !y || !step_idx || n <= 0 || c <= 0 || hw <= 0 || (mask_c != 1 && mask_c != c))
    return set_error(SDEO_EINVAL, "mask_blend_table_f32: bad args");
  return launch_k("mask_blend_table_f32", mask_blend_table_f32_kernel, dim3(grid_for((long long)n * c * hw, 256)), dim3(256), 0,
                  (cudaStream_t)stream, dim3(1, 1, 1), orig_table, img, mask, y, (const int*)step_idx, n, c, mask_c, (long long)hw);
}
extern "C" int sdeo_mask_blend_f32(const float* x0, const float* noise, const float* img, const float* mask, const float* a,
                                   const float* b, float* y, int32_t n, int32_t c, int32_t mask_c, int64_t hw, void* stream) {
  if (!x0 || !noise || !img || !mask || !a || !b || !y || n <= 0 || c <= 0 || hw <= 0 || (mask_c != 1 && mask_c != c))
    return set_error(SDEO_EINVAL, "mask_blend_f32: bad args");
  return launch_k("mask_blend_f32", mask_blend_f32_kernel, dim3(grid_for((long long)n * c * hw, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x0, noise, img, mask, a, b, y, n, c, mask_c, (long long)hw);
}
