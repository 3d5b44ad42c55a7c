// GroupNorm(32)(+SiLU) over NHWC bf16 (optionally over the channel-concat of two tensors) and LayerNorm.
// HBM-bound passes: 16-byte vector loads, fp32 statistics, deterministic two-level reduction (no atomics).
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"
#include <stdlib.h>
#include <cuda_fp16.h>

namespace sdeo {

constexpr int kGNThreads = 256;
constexpr int kGNBigPerSM = 6;  // CTAs per SM the two-pass grid is capped at (gn_geometry)
constexpr int kGNMaxChunks = 384;  // CTAs per sample the two-pass grid is capped at (the apply pass folds that many partials)

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// 8 consecutive channels as fp32 from a bf16 or an fp32 (residual-stream) tensor.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) { unpack8(*reinterpret_cast<const uint4*>(p), f); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// fp16 input (the TensorRT plugin's kHWC8 fp16 contract, groupNormPlugin.cpp:136-160)
__device__ __forceinline__ void load8(const __half* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
// two outputs packed in the kernel's 16-bit output type: bf16, except fp16 in -> fp16 out
template <typename T>
__device__ __forceinline__ uint32_t pack2_out(float a, float b) { return pack_bf16x2(a, b); }
template <>
__device__ __forceinline__ uint32_t pack2_out<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void load8_ro(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Stores 8 values back in the input's own dtype (exact round trip of load8).
__device__ __forceinline__ void store8(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// Loads the 8-channel vector `v` (of the virtual concat [x1 | x2]) at pixel `pix`.
template <typename T>
__device__ __forceinline__ void gn_load8(const T* x1, const T* x2, int c1, int c2, long long pix, int v, float* f) {
  const int c = v * 8;
  if (c < c1) load8(x1 + pix * c1 + c, f);
  else load8(x2 + pix * c2 + (c - c1), f);
}

// ---- software-pipelined row loads (the big-tensor passes are HBM-bound: bytes in flight per SM are what set their speed) ----
// An 8-channel vector as it was loaded, kept packed until it is used (4 registers for the 16-bit types). A thread
// holds two batches of kGNBatch<T> of them: the batch being processed and the next one, already in flight.
template <typename T>
struct Raw8 { uint4 u; };
template <>
struct Raw8<float> { float4 a, b; };
template <typename T>
struct GNBatch { static constexpr int H = 4; };
template <>
struct GNBatch<float> { static constexpr int H = 2; };

__device__ __forceinline__ void raw_load(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) { r.u = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void raw_load(const __half* p, Raw8<__half>& r) { r.u = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void raw_load(const float* p, Raw8<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void raw_zero(Raw8<__nv_bfloat16>& r) { r.u = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void raw_zero(Raw8<__half>& r) { r.u = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void raw_zero(Raw8<float>& r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
__device__ __forceinline__ void raw_unpack(const Raw8<__nv_bfloat16>& r, float* f) { unpack8(r.u, f); }
__device__ __forceinline__ void raw_unpack(const Raw8<__half>& r, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&r.u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float* f) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}

// This thread's column of the virtual concat [x1 | x2]: base pointer of channel vector `v` at row 0 and the row pitch.
template <typename T>
__device__ __forceinline__ const T* gn_column(const T* x1, const T* x2, int c1, int c2, int v, int* ld) {
  const int c = v * 8;
  if (c < c1) { *ld = c1; return x1 + c; }
  *ld = c2;
  return x2 + (c - c1);
}

// Rows pp, pp + R, ... (H of them) of one column; rows at or beyond p_end load nothing and read as zero.
template <typename T>
__device__ __forceinline__ void gn_load_batch(const T* col, int ld, long long row0, int pp, int R, int p_end, Raw8<T>* r) {
#pragma unroll
  for (int u = 0; u < GNBatch<T>::H; ++u) {
    const int p = pp + u * R;
    if (p < p_end) raw_load(col + (row0 + p) * ld, r[u]);
    else raw_zero(r[u]);
  }
}

// Pass 1: per (sample, pixel-chunk) partial sums per group -> ws[n][chunk][group][2]
template <typename T>
__global__ void __launch_bounds__(kGNThreads, 3)
gn_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, float* __restrict__ ws,
                int hw, int c1, int c2, int groups, int chunks, int ppc) {
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  extern __shared__ float sm[];
  const int C = c1 + c2;
  const int cv = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p_begin = chunk * ppc;
  const int p_end = min(hw, p_begin + ppc);
  float* chan_sum = sm;        // [C]
  float* chan_sq = sm + C;     // [C]
  float* part = sm + 2 * C;    // [R][cols][16]

  const int cols = cv < kGNThreads ? cv : kGNThreads;  // channel vectors handled per pass
  const int R = kGNThreads / cols;                     // pixel rows in flight
  const int tr = threadIdx.x / cols, tv = threadIdx.x % cols;
  const bool active = threadIdx.x < R * cols;

  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active && v < cv) {
      constexpr int H = GNBatch<T>::H;
      int ld;
      const T* col = gn_column(x1, x2, c1, c2, v, &ld);
      const long long row0 = (long long)n * hw;
      Raw8<T> cur[H], nxt[H];
      int pp = p_begin + tr;
      gn_load_batch(col, ld, row0, pp, R, p_end, cur);
      while (pp < p_end) {  // the next batch is in flight while this one is summed
        const int pn = pp + H * R;
        gn_load_batch(col, ld, row0, pn, R, p_end, nxt);
#pragma unroll
        for (int u = 0; u < H; ++u) {
          float f[8];
          raw_unpack(cur[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
        }
#pragma unroll
        for (int u = 0; u < H; ++u) cur[u] = nxt[u];
        pp = pn;
      }
    }
    if (active) {
      float* dst = part + ((size_t)tr * cols + tv) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = s[j]; dst[8 + j] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < cols && v < cv) {
      // thread tv (tr == 0) folds the R rows in fixed order
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
      for (int r = 0; r < R; ++r) {
        const float* src = part + ((size_t)r * cols + tv) * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += src[j]; q[j] += src[8 + j]; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { chan_sum[v * 8 + j] = s[j]; chan_sq[v * 8 + j] = q[j]; }
    }
    __syncthreads();
  }
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += chan_sum[c]; q += chan_sq[c]; }
    float* out = ws + (((size_t)n * chunks + chunk) * groups + g) * 2;
    out[0] = s;
    out[1] = q;
  }  trace_mark(trc, 3);
}

// SiLU of two values with ONE reciprocal: x / (1 + e^-x) for both from r = 1 / ((1 + e0)(1 + e1)) -- 1.5 MUFU operations per
// value instead of 2 (ncu: the apply pass of the big-tensor GroupNorm sat at 71 % of the XU pipe while moving 5.96 TB/s).
// The exponent is clamped to 63 so the product of the two denominators stays finite (x < -43.7: the result is ~1e-18 x, i.e.
// zero in any 16-bit output either way).
__device__ __forceinline__ void silu_pair(float& x0, float& x1) {
  constexpr float kNegLog2e = -1.4426950408889634f;
  float e0, e1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x0 * kNegLog2e, 63.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x1 * kNegLog2e, 63.f)));
  const float d0 = 1.0f + e0, d1 = 1.0f + e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  x0 = x0 * (d1 * r);
  x1 = x1 * (d0 * r);
}

// One column (8 channels) of a CTA's pixel rows: y = x * a + b (+SiLU), rows pp0, pp0 + R, ... < p_end; the next batch
// of rows is in flight while this one is normalised and stored. Output: 16-bit, bf16 (fp16 for fp16 inputs).
template <typename T>
__device__ __forceinline__ void gn_apply_column(const T* col, int ld, long long row0, int pp0, int R, int p_end,
                                                const float* a, const float* b, int with_silu, __nv_bfloat16* ycol, int C) {
  constexpr int H = GNBatch<T>::H;
  Raw8<T> cur[H], nxt[H];
  int pp = pp0;
  gn_load_batch(col, ld, row0, pp, R, p_end, cur);
  while (pp < p_end) {
    const int pn = pp + H * R;
    gn_load_batch(col, ld, row0, pn, R, p_end, nxt);
#pragma unroll
    for (int u = 0; u < H; ++u) {
      const int p = pp + u * R;
      if (p < p_end) {
        float f[8];
        raw_unpack(cur[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], a[j], b[j]);
        if (with_silu) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) silu_pair(f[j], f[j + 1]);
        }
        uint4 o;
        o.x = pack2_out<T>(f[0], f[1]); o.y = pack2_out<T>(f[2], f[3]);
        o.z = pack2_out<T>(f[4], f[5]); o.w = pack2_out<T>(f[6], f[7]);
        *reinterpret_cast<uint4*>(ycol + (row0 + p) * C) = o;
      }
    }
#pragma unroll
    for (int u = 0; u < H; ++u) cur[u] = nxt[u];
    pp = pn;
  }
}

// Pass 2: finalize mean / rstd per group from the partials, normalise, affine, optional SiLU, store bf16.
template <typename T>
__global__ void __launch_bounds__(kGNThreads, 3)
gn_apply_kernel(const T* __restrict__ x1, const T* __restrict__ x2,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ ws,
                __nv_bfloat16* __restrict__ y, int hw, int c1, int c2, int groups, int chunks, int ppc, float eps,
                int with_silu) {
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_fold[kGNThreads];
  const int C = c1 + c2;
  const int cv = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y, chunk = blockIdx.x;
  {
    // fold the per-chunk partials: kGNThreads / groups lanes per group take chunks lane, lane + L, ... (independent
    // loads), then one thread per group adds the L lane sums in fixed order (deterministic)
    const int L = kGNThreads / groups;
    const int g = threadIdx.x % groups, lane_k = threadIdx.x / groups;
    float s = 0.f, q = 0.f;
    if (lane_k < L) {
      // eight independent loads per round trip to L2 (a mid-size sample has ~220 chunks: 4 rounds instead of 28), added
      // in chunk order
      const float2* src = reinterpret_cast<const float2*>(ws) + (size_t)n * chunks * groups + g;
      for (int k0 = lane_k; k0 < chunks; k0 += 8 * L) {
        float2 v2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k0 + u * L;
          v2[u] = k < chunks ? __ldcg(src + (size_t)k * groups) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s += v2[u].x; q += v2[u].y; }
      }
    }
    s_fold[threadIdx.x] = make_float2(s, q);
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const int L = kGNThreads / groups;
    float s = 0.f, q = 0.f;
    for (int k = 0; k < L; ++k) {
      s += s_fold[k * groups + g].x;
      q += s_fold[k * groups + g].y;
    }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = s * inv;
    float var = q * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int p_begin = chunk * ppc;
  const int p_end = min(hw, p_begin + ppc);
  const int cols = cv < kGNThreads ? cv : kGNThreads;
  const int R = kGNThreads / cols;
  const int tr = threadIdx.x / cols, tv = threadIdx.x % cols;
  if (threadIdx.x >= R * cols) return;
  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    if (v >= cv) continue;
    float a[8], b[8];
    load8_ro(gamma + v * 8, a);
    load8_ro(beta + v * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (v * 8 + j) / cpg;
      a[j] *= s_rstd[g];
      b[j] -= s_mean[g] * a[j];
    }
    int ld;
    const T* col = gn_column(x1, x2, c1, c2, v, &ld);
    gn_apply_column<T>(col, ld, (long long)n * hw, p_begin + tr, R, p_end, a, b, with_silu, y + v * 8, C);
  }
  trace_mark(trc, 3);
}

// ---------------------------------------------------------------------------------------------------------
// Single-launch GroupNorm for the denoiser's small tensors: ONE THREAD-BLOCK CLUSTER of kGNCluster CTAs per sample.
// Each CTA reduces its share of the pixels to per-group (sum, sumsq) in its shared memory, the cluster exchanges the
// 2*groups partials through distributed shared memory (rank order: deterministic), then every CTA normalises its own
// pixels (second read comes from L2). Replaces the stats + apply pair (2 launches, a global workspace round trip).
// ---------------------------------------------------------------------------------------------------------
constexpr int kGNCluster = 8;       // portable cluster size
constexpr int kGNClusterBig = 16;   // non-portable size, used when a sample has enough rows
constexpr int kGNCThreads = 512;
constexpr int kGNCMaxSmem = 160 * 1024;

template <typename T>
__global__ void __launch_bounds__(kGNCThreads)
gn_cluster_kernel(const T* __restrict__ x1, const T* __restrict__ x2, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int hw, int c1, int c2, int groups,
                  float eps, int with_silu, int cache_rows) {
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  extern __shared__ float sm[];
  __shared__ float2 all[kGNClusterBig * 64];  // [cluster rank][group] (sum, sumsq) partials of the whole cluster
  __shared__ float s_mean[64], s_rstd[64];
  const int C = c1 + c2;
  const int cv = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y, rank = blockIdx.x;  // cluster = the gridDim.x CTAs along x
  const int csize = gridDim.x;
  const int ppc = (hw + csize - 1) / csize;
  const int p_begin = rank * ppc;
  const int p_end = min(hw, p_begin + ppc);
  float* chan_sum = sm;        // [C]
  float* chan_sq = sm + C;     // [C]
  float* part = sm + 2 * C;    // [R][cols][16]
  T* cache = reinterpret_cast<T*>(part + kGNCThreads * 16);  // [cache_rows][C]: this CTA's rows (pass 2 reads them here)
  const int cols = cv < kGNCThreads ? cv : kGNCThreads;
  const int R = kGNCThreads / cols;
  const int tr = threadIdx.x / cols, tv = threadIdx.x % cols;
  const bool active = threadIdx.x < R * cols;

  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active && v < cv) {
      // rows in batches of four independent loads; rows past the end are skipped by predicate (a thread of a small sample has
      // 1-3 rows: one round trip to memory, not one per row)
      int ld;
      const T* col = gn_column(x1, x2, c1, c2, v, &ld);
      const long long row0 = (long long)n * hw;
      for (int pp = p_begin + tr; pp < p_end; pp += 4 * R) {
        Raw8<T> r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (pp + u * R < p_end) raw_load(col + (row0 + pp + u * R) * ld, r[u]);
          else raw_zero(r[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          raw_unpack(r[u], f);
          if (cache_rows && pp + u * R < p_end) store8(cache + (size_t)(pp + u * R - p_begin) * C + v * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
        }
      }
    }
    if (active) {
      float* dst = part + ((size_t)tr * cols + tv) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = s[j]; dst[8 + j] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < cols && v < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
      for (int r = 0; r < R; ++r) {
        const float* src = part + ((size_t)r * cols + tv) * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += src[j]; q[j] += src[8 + j]; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { chan_sum[v * 8 + j] = s[j]; chan_sq[v * 8 + j] = q[j]; }
    }
    __syncthreads();
  }
  // ---- exchange the partials across the cluster: every CTA PUSHES its (sum, sumsq) per group into the all[] array of
  // every peer (remote stores are fire-and-forget; remote loads would cost one round trip each), one cluster barrier
  // publishes them, then each CTA folds its local copy in rank order (deterministic) ----
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += chan_sum[c]; q += chan_sq[c]; }
    const uint32_t slot = smem_u32(all) + (uint32_t)((rank * groups + g) * 8);
    for (int r = 0; r < csize; ++r) {
      uint32_t peer;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(slot), "r"(r));
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(peer), "f"(s), "f"(q) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int r = 0; r < csize; ++r) {
      const float2 v2 = all[r * groups + g];
      s += v2.x;
      q += v2.y;
    }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = s * inv;
    float var = q * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();  // (no peer touches this CTA's shared memory after the cluster barrier above)

  // ---- normalise this CTA's pixels ----
  if (!active) return;
  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    if (v >= cv) continue;
    float a[8], b[8];
    load8_ro(gamma + v * 8, a);
    load8_ro(beta + v * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (v * 8 + j) / cpg;
      a[j] *= s_rstd[g];
      b[j] -= s_mean[g] * a[j];
    }
    for (int pp = p_begin + tr; pp < p_end; pp += R) {
      const long long pix = (long long)n * hw + pp;
      float f[8];
      if (cache_rows) load8(cache + (size_t)(pp - p_begin) * C + v * 8, f);
      else gn_load8(x1, x2, c1, c2, pix, v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = f[j] * a[j] + b[j];
        f[j] = with_silu ? silu_f(t) : t;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(y + pix * C + v * 8) = o;
    }
  }
  trace_mark(trc, 3);
}

template <typename T>
static int launch_gn_cluster(const void* x1, const void* x2, const float* gamma, const float* beta, void* y, int n, int hw,
                             int c1, int c2, int groups, float eps, int with_silu, cudaStream_t st) {
  const int C = c1 + c2;
  static int attr_state = 0;  // 0: not configured, 1: portable sizes only, 2: 16-CTA clusters allowed
  if (!attr_state) {
    cudaError_t e = cudaFuncSetAttribute(gn_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGNCMaxSmem);
    if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(gn_cluster_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    attr_state = (e == cudaSuccess && !getenv("SDEO_GN_CLUSTER8")) ? 2 : 1;
    (void)cudaGetLastError();
  }
  const int csize = (attr_state == 2 && hw >= 4 * kGNClusterBig) ? kGNClusterBig : kGNCluster;
  const int ppc = (hw + csize - 1) / csize;
  const size_t base = (size_t)(2 * C + kGNCThreads * 16) * sizeof(float);
  const size_t cache = (size_t)ppc * C * sizeof(T);
  const int cache_rows = (base + cache <= (size_t)kGNCMaxSmem) ? ppc : 0;  // rows kept in shared memory: single read
  const size_t smem = base + (cache_rows ? cache : 0);
  return launch_k("groupnorm (cluster)", gn_cluster_kernel<T>, dim3((unsigned)csize, (unsigned)n), dim3(kGNCThreads), smem, st,
                  dim3((unsigned)csize, 1, 1), (const T*)x1, (const T*)x2, gamma, beta, (__nv_bfloat16*)y, hw, c1, c2, groups,
                  eps, with_silu, cache_rows);
}

// Grid of the two-pass kernels: `chunks` CTAs per sample of `ppc` pixel rows each. row_bytes = bytes of one pixel row.
static void gn_geometry(int n, int hw, long long row_bytes, int* chunks, int* ppc) {
  // small tensors (the denoiser's): about 1.5 CTAs per SM over the batch, and at least 16 pixels per CTA (per-CTA fixed
  // costs and the apply kernel's per-CTA fold over the chunk partials dominate otherwise)
  static int per_sm_x2 = -1;  // CTAs per SM (x2) over the batch; SDEO_GN_CTAS_X2 overrides (tuning aid)
  if (per_sm_x2 < 0) {
    const char* e = getenv("SDEO_GN_CTAS_X2");
    per_sm_x2 = e ? atoi(e) : 3;
  }
  // larger ones (the VAE decoder: up to 16 x 512 x 512 pixels) are bandwidth-bound: one CTA per 64 KB of the batch, up to
  // kGNBigPerSM per SM = two waves of the three CTAs an SM holds (the pipelined row loads keep 4-8 vectors per thread in
  // flight: 48-96 KB per SM), and never more than kGNMaxChunks per sample: every apply CTA folds its sample's partials
  // (8 lanes x 8 loads per L2 round trip: 384 chunks = 6 round trips). Swept on B200 over 32 / 64 / 128 KB and caps of
  // 128 ... 1024 chunks (tools/sweep_gn_geometry.sh, profiles/r02b_gn_geometry_sweep.txt): single samples of 6-25 MB run
  // 13.1 / 16.9 / 25.5 us here, up to 40 us with 1024 chunks (fold) or 128 chunks (too few bytes in flight)
  static int chunk_kb = -1, max_chunks = -1;  // tuning aids: SDEO_GN_CHUNK_KB, SDEO_GN_MAX_CHUNKS (<= kGNMaxChunks)
  if (chunk_kb < 0) {
    const char* e = getenv("SDEO_GN_CHUNK_KB");
    chunk_kb = e && atoi(e) > 0 ? atoi(e) : 64;
    e = getenv("SDEO_GN_MAX_CHUNKS");
    max_chunks = e && atoi(e) > 0 && atoi(e) <= kGNMaxChunks ? atoi(e) : kGNMaxChunks;
  }
  long long total = ((long long)n * hw * row_bytes) / ((long long)chunk_kb << 10);
  const bool big = total > 148 * kGNBigPerSM;
  if (big) total = 148 * kGNBigPerSM;
  if (total < 148 * per_sm_x2 / 2) total = 148 * per_sm_x2 / 2;
  // big tensors: never more CTAs than the cap (whole waves of the resident set); small ones: round up
  int want = big ? (int)(total / n) : (int)((total + n - 1) / n);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  int p = (hw + want - 1) / want;
  if (p < 16) p = hw < 16 ? hw : 16;
  *ppc = p;
  *chunks = (hw + p - 1) / p;
}
// Upper bound of `chunks` over every row size (the workspace queries do not know the channel count).
static int gn_max_chunks(int n, int hw) {
  (void)n;
  return hw < kGNMaxChunks ? hw : kGNMaxChunks;  // chunks <= want <= kGNMaxChunks, and a chunk holds at least one pixel
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm from precomputed partial statistics: the convolutions that PRODUCED x1 / x2 left per-channel (sum, sum of
// squares) partials of their final outputs (sdeo_conv_args::gn_stats, [sample][part][channel]); every CTA folds the
// partials of its sample in fixed order (deterministic), forms the group statistics (any grouping, including groups
// that straddle the concat seam) and normalises its pixel rows. The tensor is read once; no statistics pass.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGNThreads)
gn_apply_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, const float2* __restrict__ st1, int parts1,
                      const float2* __restrict__ st2, int parts2, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int hw, int c1, int c2, int groups,
                      int gslab, int ppc, float eps, int with_silu) {
  // grid: (pixel-row chunk, channel slab of `gslab` groups, sample). A CTA folds only its slab's partials.
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  extern __shared__ float sm[];
  __shared__ float s_mean[64], s_rstd[64];
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int n = blockIdx.z, chunk = blockIdx.x, slab = blockIdx.y;
  const int g_lo = slab * gslab;
  const int ng = min(gslab, groups - g_lo);
  const int c_lo = g_lo * cpg, cs = ng * cpg;  // this CTA's channels [c_lo, c_lo + cs), cs % 8 == 0
  float* chan_sum = sm;       // [cs]
  float* chan_sq = sm + cs;   // [cs]
  // The fold sits in front of every CTA's first load of the tensor, and each pass of 4 partials is an L2 round trip: the
  // `pgs` thread groups of a channel take partials pg, pg + pgs, ... eight loads at a time (24..72 partials: 1-3 round trips),
  // then one thread per channel adds the group sums in group order (fixed order: deterministic).
  const int pgs = cs <= kGNThreads ? min(8, kGNThreads / cs) : 1;
  float2* part = reinterpret_cast<float2*>(sm + 2 * cs);   // [pgs][cs]
  for (int idx = threadIdx.x; idx < cs * pgs; idx += kGNThreads) {
    const int cl = idx % cs, pg = idx / cs;
    const int c = c_lo + cl;
    const bool first = c < c1;
    const float2* src = first ? st1 + (size_t)n * parts1 * c1 + c : st2 + (size_t)n * parts2 * c2 + (c - c1);
    const int parts = first ? parts1 : parts2, ld = first ? c1 : c2;
    float s = 0.f, q = 0.f;
    for (int k0 = pg; k0 < parts; k0 += 8 * pgs) {  // 8 independent loads per L2 round trip (no serial tail), added in slot order
      float2 v2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * pgs;
        v2[u] = k < parts ? __ldcg(src + (size_t)k * ld) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += v2[u].x; q += v2[u].y; }
    }
    part[(size_t)pg * cs + cl] = make_float2(s, q);
  }
  __syncthreads();
  for (int cl = threadIdx.x; cl < cs; cl += kGNThreads) {
    float s = 0.f, q = 0.f;
    for (int pg = 0; pg < pgs; ++pg) {
      const float2 v2 = part[(size_t)pg * cs + cl];
      s += v2.x; q += v2.y;
    }
    chan_sum[cl] = s;
    chan_sq[cl] = q;
  }
  __syncthreads();
  if (threadIdx.x < ng) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int cl = g * cpg; cl < (g + 1) * cpg; ++cl) { s += chan_sum[cl]; q += chan_sq[cl]; }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = s * inv;
    float var = q * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int p_begin = chunk * ppc;
  const int p_end = min(hw, p_begin + ppc);
  const int cvs = cs / 8, v_lo = c_lo / 8;
  const int cols = cvs < kGNThreads ? cvs : kGNThreads;
  const int R = kGNThreads / cols;
  const int tr = threadIdx.x / cols, tv = threadIdx.x % cols;
  if (threadIdx.x < R * cols) {
    for (int vbase = 0; vbase < cvs; vbase += cols) {
      if (vbase + tv >= cvs) continue;
      const int v = v_lo + vbase + tv;  // 8-channel vector of the virtual concat
      float a[8], b[8];
      load8_ro(gamma + v * 8, a);
      load8_ro(beta + v * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (v * 8 + j) / cpg - g_lo;
        a[j] *= s_rstd[g];
        b[j] -= s_mean[g] * a[j];
      }
      int ld;
      const T* col = gn_column(x1, x2, c1, c2, v, &ld);
      gn_apply_column<T>(col, ld, (long long)n * hw, p_begin + tr, R, p_end, a, b, with_silu, y + v * 8, C);
    }
  }
  trace_mark(trc, 3);
}

// One warp per row, the row lives in registers (C <= 2048): exact two-pass mean/variance.
constexpr int kLNMaxVec = 8;
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int rows, int C, float eps) {
  const int trc = trace_start(4);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int cv = C / 8;
  const T* xr = x + (size_t)warp * C;
  float f[kLNMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLNMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < cv) {
      load8(xr + v * 8, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLNMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; q += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
  __nv_bfloat16* yr = y + (size_t)warp * C;
#pragma unroll
  for (int i = 0; i < kLNMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < cv) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      float r[8];
      r[0] = (f[i][0] - mean) * rstd * g0.x + b0.x; r[1] = (f[i][1] - mean) * rstd * g0.y + b0.y;
      r[2] = (f[i][2] - mean) * rstd * g0.z + b0.z; r[3] = (f[i][3] - mean) * rstd * g0.w + b0.w;
      r[4] = (f[i][4] - mean) * rstd * g1.x + b1.x; r[5] = (f[i][5] - mean) * rstd * g1.y + b1.y;
      r[6] = (f[i][6] - mean) * rstd * g1.z + b1.z; r[7] = (f[i][7] - mean) * rstd * g1.w + b1.w;
      uint4 o;
      o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]);
      o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
      *reinterpret_cast<uint4*>(yr + v * 8) = o;
    }
  }
  trace_mark(trc, 3);
}

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_norm)

extern "C" size_t sdeo_groupnorm_workspace_bytes(int32_t n, int32_t hw, int32_t groups) {
  return (size_t)n * gn_max_chunks(n, hw) * groups * 2 * sizeof(float);
}

extern "C" int sdeo_groupnorm_plan(int32_t n, int32_t hw, int64_t row_bytes, int32_t* plan) {
  if (!plan || n <= 0 || hw <= 0 || row_bytes <= 0) return set_error(SDEO_EINVAL, "groupnorm_plan: bad argument");
  int chunks, ppc;
  gn_geometry(n, hw, (long long)row_bytes, &chunks, &ppc);
  plan[0] = chunks;
  plan[1] = ppc;
  return SDEO_OK;
}

extern "C" int sdeo_groupnorm_nhwc(const void* x1, const void* x2, int32_t x_f32, const float* gamma, const float* beta,
                                   void* y, int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps,
                                   int32_t with_silu, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x1 || !gamma || !beta || !y || !workspace) return set_error(SDEO_EINVAL, "groupnorm: null argument");
  if (!x2) c2 = 0;
  const int C = c1 + c2;
  if (n <= 0 || hw <= 0 || groups <= 0 || groups > 64 || C % groups != 0 || c1 % 8 != 0 || c2 % 8 != 0)
    return set_error(SDEO_EINVAL, "groupnorm: unsupported geometry (need C % groups == 0, channels % 8 == 0, groups <= 64)");
  if (n > 65535) return set_error(SDEO_EINVAL, "groupnorm: batch too large");
  // small tensors: one cluster per sample, single launch; big ones (VAE): two-pass grid
  static long long cluster_max = -1;  // elements per sample up to which the single-launch cluster variant is used
  if (cluster_max < 0) {
    const char* e = getenv("SDEO_GN_CLUSTER_MAX");
    cluster_max = e ? atoll(e) : 300000;  // measured on the step graph: larger tensors are faster as stats + apply over the whole GPU
  }
  if ((long long)hw * C <= cluster_max && hw >= kGNCluster && (size_t)(2 * C + kGNCThreads * 16) * 4 <= 100 * 1024 &&
      !getenv("SDEO_GN_TWO_PASS")) {
    if (x_f32)
      return launch_gn_cluster<float>(x1, x2, gamma, beta, y, n, hw, c1, c2, groups, eps, with_silu, (cudaStream_t)stream);
    return launch_gn_cluster<__nv_bfloat16>(x1, x2, gamma, beta, y, n, hw, c1, c2, groups, eps, with_silu, (cudaStream_t)stream);
  }
  int chunks, ppc;
  gn_geometry(n, hw, (long long)C * (x_f32 ? 4 : 2), &chunks, &ppc);
  if (workspace_bytes < (size_t)n * chunks * groups * 2 * sizeof(float))
    return set_error(SDEO_EINVAL, "groupnorm: workspace too small");
  const size_t smem = (size_t)(2 * C + kGNThreads * 16) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
      attr_set = true;
    }
    if (smem > 100 * 1024) return set_error(SDEO_EINVAL, "groupnorm: too many channels");
  }
  dim3 grid((unsigned)chunks, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 one(1, 1, 1);
  int rc;
  if (x_f32)
    rc = launch_k("groupnorm stats", gn_stats_kernel<float>, grid, dim3(kGNThreads), smem, st, one, (const float*)x1,
                  (const float*)x2, (float*)workspace, hw, c1, c2, groups, chunks, ppc);
  else
    rc = launch_k("groupnorm stats", gn_stats_kernel<__nv_bfloat16>, grid, dim3(kGNThreads), smem, st, one,
                  (const __nv_bfloat16*)x1, (const __nv_bfloat16*)x2, (float*)workspace, hw, c1, c2, groups, chunks, ppc);
  if (rc) return rc;
  if (x_f32)
    return launch_k("groupnorm apply", gn_apply_kernel<float>, grid, dim3(kGNThreads), 0, st, one, (const float*)x1,
                    (const float*)x2, gamma, beta, (const float*)workspace, (__nv_bfloat16*)y, hw, c1, c2, groups, chunks,
                    ppc, eps, with_silu);
  return launch_k("groupnorm apply", gn_apply_kernel<__nv_bfloat16>, grid, dim3(kGNThreads), 0, st, one,
                  (const __nv_bfloat16*)x1, (const __nv_bfloat16*)x2, gamma, beta, (const float*)workspace,
                  (__nv_bfloat16*)y, hw, c1, c2, groups, chunks, ppc, eps, with_silu);
}

// The plugin contract (fp16 NHWC in / out) as the two-launch grid (statistics + apply): the A/B partner and fallback of
// the streamed kernel in groupnorm_stream.cu, which owns the C entry point sdeo_groupnorm_nhwc_f16.
namespace sdeo {
size_t groupnorm_two_pass_workspace_bytes(int32_t n, int32_t hw, int32_t groups) {
  return (size_t)n * gn_max_chunks(n, hw) * groups * 2 * sizeof(float);
}
int groupnorm_f16_two_pass(const void* x, const float* gamma, const float* beta, void* y, int32_t n, int32_t hw,
                                       int32_t c, int32_t groups, float eps, int32_t with_silu, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!x || !gamma || !beta || !y || !workspace) return set_error(SDEO_EINVAL, "groupnorm_f16: null argument");
  if (n <= 0 || n > 65535 || hw <= 0 || groups <= 0 || groups > 64 || c % groups != 0 || c % 8 != 0)
    return set_error(SDEO_EINVAL, "groupnorm_f16: unsupported geometry (need C % groups == 0, C % 8 == 0, groups <= 64)");
  int chunks, ppc;
  gn_geometry(n, hw, (long long)c * 2, &chunks, &ppc);
  if (workspace_bytes < (size_t)n * chunks * groups * 2 * sizeof(float))
    return set_error(SDEO_EINVAL, "groupnorm_f16: workspace too small");
  const size_t smem = (size_t)(2 * c + kGNThreads * 16) * sizeof(float);
  if (smem > 100 * 1024) return set_error(SDEO_EINVAL, "groupnorm_f16: too many channels");
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      if (cudaFuncSetAttribute(gn_stats_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
        return set_error(SDEO_ECUDA, "groupnorm_f16: cannot raise the shared-memory limit");
      attr_set = true;
    }
  }
  dim3 grid((unsigned)chunks, (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 one(1, 1, 1);
  int rc = launch_k("groupnorm_f16 stats", gn_stats_kernel<__half>, grid, dim3(kGNThreads), smem, st, one, (const __half*)x,
                    (const __half*)nullptr, (float*)workspace, hw, c, 0, groups, chunks, ppc);
  if (rc) return rc;
  return launch_k("groupnorm_f16 apply", gn_apply_kernel<__half>, grid, dim3(kGNThreads), 0, st, one, (const __half*)x,
                  (const __half*)nullptr, gamma, beta, (const float*)workspace, (__nv_bfloat16*)y, hw, c, 0, groups, chunks, ppc,
                  eps, with_silu);
}
}  // namespace sdeo

extern "C" int sdeo_groupnorm_apply_stats(const void* x1, const void* x2, int32_t x_f32, const float* stats1, int32_t parts1,
                                          const float* stats2, int32_t parts2, const float* gamma, const float* beta, void* y,
                                          int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps,
                                          int32_t with_silu, void* stream) {
  if (!x1 || !stats1 || !gamma || !beta || !y || parts1 <= 0) return set_error(SDEO_EINVAL, "groupnorm_apply_stats: null argument");
  if (!x2) c2 = 0;
  if (x2 && (!stats2 || parts2 <= 0)) return set_error(SDEO_EINVAL, "groupnorm_apply_stats: x2 given without its statistics");
  const int C = c1 + c2;
  if (n <= 0 || n > 65535 || hw <= 0 || groups <= 0 || groups > 64 || C % groups != 0 || c1 % 8 != 0 || c2 % 8 != 0)
    return set_error(SDEO_EINVAL, "groupnorm_apply_stats: unsupported geometry");
  // grid: (row chunks, channel slabs, samples). A slab = a quarter of the groups when that keeps 8-channel vectors whole:
  // every CTA then folds a quarter of the partial statistics. About two CTAs per SM, at least 4 pixel rows each.
  const int cpg = C / groups;
  int gslab = groups;
  if (groups % 4 == 0 && ((groups / 4) * cpg) % 8 == 0) gslab = groups / 4;
  const int slabs = (groups + gslab - 1) / gslab;
  if ((gslab * cpg) % 8 != 0 && slabs > 1) return set_error(SDEO_EINVAL, "groupnorm_apply_stats: slab not vector aligned");
  int want = (296 + n * slabs - 1) / (n * slabs);
  if (want < 1) want = 1;
  int ppc = (hw + want - 1) / want;
  if (ppc < 4) ppc = hw < 4 ? hw : 4;
  const int chunks = (hw + ppc - 1) / ppc;
  // channel sums [2][cs] + the fold's group partials [pgs][cs] float2 (pgs as in the kernel)
  const int cs_max = gslab * cpg;
  const int pgs = cs_max <= kGNThreads ? (kGNThreads / cs_max < 8 ? kGNThreads / cs_max : 8) : 1;
  const size_t smem = (size_t)2 * cs_max * sizeof(float) * (size_t)(1 + pgs);
  const dim3 grid((unsigned)chunks, (unsigned)slabs, (unsigned)n), one(1, 1, 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_f32)
    return launch_k("groupnorm apply (stats)", gn_apply_stats_kernel<float>, grid, dim3(kGNThreads), smem, st, one,
                    (const float*)x1, (const float*)x2, (const float2*)stats1, parts1, (const float2*)stats2, parts2, gamma, beta,
                    (__nv_bfloat16*)y, hw, c1, c2, groups, gslab, ppc, eps, with_silu);
  return launch_k("groupnorm apply (stats)", gn_apply_stats_kernel<__nv_bfloat16>, grid, dim3(kGNThreads), smem, st, one,
                  (const __nv_bfloat16*)x1, (const __nv_bfloat16*)x2, (const float2*)stats1, parts1, (const float2*)stats2,
                  parts2, gamma, beta, (__nv_bfloat16*)y, hw, c1, c2, groups, gslab, ppc, eps, with_silu);
}

// ---------------------------------------------------------------------------------------------------------
// Folds GroupNorm partial statistics [n][parts][c] (sdeo_conv_args::gn_stats) down to [n][out_parts][c], out_parts =
// ceil(parts / 256): large feature maps leave thousands of partial slots per sample (one per 128-pixel M tile), too many
// for every consumer CTA to fold on its own (folded GroupNorm, sdeo_conv_args::gnf_*). Fixed summation order.
// grid: (32-channel slabs, out_parts, n); 256 threads = 8 slot lanes x 32 channels.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_fold_kernel(const float2* __restrict__ st, float2* __restrict__ out, int parts, int out_parts, int c) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float2 sm[8][32];
  const int cl = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  const int n = blockIdx.z, op = blockIdx.y;
  const int k_lo = op * 256, k_hi = min(parts, k_lo + 256);
  float s = 0.f, q = 0.f;
  if (ch < c) {
    const float2* src = st + ((size_t)n * parts) * c + ch;
    for (int k0 = k_lo + pg; k0 < k_hi; k0 += 64) {  // 8 independent loads per round trip, added in slot order
      float2 v2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 8;
        v2[u] = k < k_hi ? __ldcg(src + (size_t)k * c) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += v2[u].x; q += v2[u].y; }
    }
  }
  sm[pg][cl] = make_float2(s, q);
  __syncthreads();
  if (pg == 0 && ch < c) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ts += sm[j][cl].x; tq += sm[j][cl].y; }
    out[((size_t)n * out_parts + op) * c + ch] = make_float2(ts, tq);
  }
}

extern "C" int sdeo_gn_stats_fold(const float* stats, float* out, int32_t n, int32_t parts, int32_t c, int32_t* out_parts,
                                  void* stream) {
  if (n <= 0 || parts <= 0 || c <= 0 || n > 65535) return set_error(SDEO_EINVAL, "gn_stats_fold: bad geometry");
  const int op = (parts + 255) / 256;
  if (out_parts) *out_parts = op;
  if (!stats || !out) return SDEO_OK;  // geometry query
  const dim3 grid((unsigned)((c + 31) / 32), (unsigned)op, (unsigned)n), one(1, 1, 1);
  return launch_k("gn_stats_fold", gn_stats_fold_kernel, grid, dim3(256), 0, (cudaStream_t)stream, one, (const float2*)stats,
                  (float2*)out, parts, op, c);
}

extern "C" int sdeo_layernorm(const void* x, int32_t x_f32, const float* gamma, const float* beta, void* y, int32_t rows,
                              int32_t c, float eps, void* stream) {
  if (!x || !gamma || !beta || !y) return set_error(SDEO_EINVAL, "layernorm: null argument");
  if (rows <= 0 || c % 8 != 0 || c > kLNMaxVec * 32 * 8) return set_error(SDEO_EINVAL, "layernorm: need C % 8 == 0 and C <= 2048");
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  const dim3 one(1, 1, 1);
  if (x_f32)
    return launch_k("layernorm", layernorm_kernel<float>, dim3(blocks), dim3(warps_per_block * 32), 0, (cudaStream_t)stream,
                    one, (const float*)x, gamma, beta, (__nv_bfloat16*)y, rows, c, eps);
  return launch_k("layernorm", layernorm_kernel<__nv_bfloat16>, dim3(blocks), dim3(warps_per_block * 32), 0,
                  (cudaStream_t)stream, one, (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, rows, c, eps);
}
