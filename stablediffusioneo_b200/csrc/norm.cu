// GroupNorm(32)(+SiLU) over NHWC bf16 (optionally over the channel-concat of two tensors) and LayerNorm.
// HBM-bound passes: 16-byte vector loads, fp32 statistics, deterministic two-level reduction (no atomics).
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"
#include <stdlib.h>

namespace sdeo {

constexpr int kGNThreads = 256;

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

// Loads the 8-channel vector `v` (of the virtual concat [x1 | x2]) at pixel `pix`.
template <typename T>
__device__ __forceinline__ void gn_load8(const T* x1, const T* x2, int c1, int c2, long long pix, int v, float* f) {
  const int c = v * 8;
  if (c < c1) load8(x1 + pix * c1 + c, f);
  else load8(x2 + pix * c2 + (c - c1), f);
}

// Pass 1: per (sample, pixel-chunk) partial sums per group -> ws[n][chunk][group][2]
template <typename T>
__global__ void __launch_bounds__(kGNThreads)
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
      int pp = p_begin + tr;
      for (; pp + 3 * R < p_end; pp += 4 * R) {  // 4 independent loads in flight per thread
        float f0[8], f1[8], f2[8], f3[8];
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp, v, f0);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + R, v, f1);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + 2 * R, v, f2);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + 3 * R, v, f3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
          q[j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
        }
      }
      for (; pp < p_end; pp += R) {
        float f[8];
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp, v, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
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
  }
}

// Pass 2: finalize mean / rstd per group from the partials, normalise, affine, optional SiLU, store bf16.
template <typename T>
__global__ void __launch_bounds__(kGNThreads)
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
      for (int k = lane_k; k < chunks; k += L) {
        const float2 v2 = __ldcg(reinterpret_cast<const float2*>(ws + (((size_t)n * chunks + k) * groups + g) * 2));
        s += v2.x;
        q += v2.y;
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
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      const float ga = gamma[c] * s_rstd[g];
      a[j] = ga;
      b[j] = beta[c] - s_mean[g] * ga;
    }
    int pp = p_begin + tr;
    for (; pp + 3 * R < p_end; pp += 4 * R) {  // 4 independent loads in flight per thread
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + u * R, v, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = f[u][j] * a[j] + b[j];
          f[u][j] = with_silu ? silu_f(t) : t;
        }
        uint4 o;
        o.x = pack_bf16x2(f[u][0], f[u][1]); o.y = pack_bf16x2(f[u][2], f[u][3]);
        o.z = pack_bf16x2(f[u][4], f[u][5]); o.w = pack_bf16x2(f[u][6], f[u][7]);
        *reinterpret_cast<uint4*>(y + ((long long)n * hw + pp + u * R) * C + v * 8) = o;
      }
    }
    for (; pp < p_end; pp += R) {
      const long long pix = (long long)n * hw + pp;
      float f[8];
      gn_load8(x1, x2, c1, c2, pix, v, f);
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

// ---------------------------------------------------------------------------------------------------------
// Single-launch GroupNorm for the denoiser's small tensors: ONE THREAD-BLOCK CLUSTER of kGNCluster CTAs per sample.
// Each CTA reduces its share of the pixels to per-group (sum, sumsq) in its shared memory, the cluster exchanges the
// 2*groups partials through distributed shared memory (rank order: deterministic), then every CTA normalises its own
// pixels (second read comes from L2). Replaces the stats + apply pair (2 launches, a global workspace round trip).
// ---------------------------------------------------------------------------------------------------------
constexpr int kGNCluster = 8;
constexpr int kGNCThreads = 512;

template <typename T>
__global__ void __launch_bounds__(kGNCThreads)
gn_cluster_kernel(const T* __restrict__ x1, const T* __restrict__ x2, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int hw, int c1, int c2, int groups,
                  float eps, int with_silu) {
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  extern __shared__ float sm[];
  __shared__ float grp[128];            // [2][groups] this CTA's partial sums, read by the peers
  __shared__ float s_mean[64], s_rstd[64];
  const int C = c1 + c2;
  const int cv = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y, rank = blockIdx.x;  // cluster = the kGNCluster CTAs along x
  const int ppc = (hw + kGNCluster - 1) / kGNCluster;
  const int p_begin = rank * ppc;
  const int p_end = min(hw, p_begin + ppc);
  float* chan_sum = sm;        // [C]
  float* chan_sq = sm + C;     // [C]
  float* part = sm + 2 * C;    // [R][cols][16]
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
      int pp = p_begin + tr;
      for (; pp + 3 * R < p_end; pp += 4 * R) {  // 4 independent loads in flight per thread
        float f0[8], f1[8], f2[8], f3[8];
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp, v, f0);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + R, v, f1);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + 2 * R, v, f2);
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp + 3 * R, v, f3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
          q[j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
        }
      }
      for (; pp < p_end; pp += R) {
        float f[8];
        gn_load8(x1, x2, c1, c2, (long long)n * hw + pp, v, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
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
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += chan_sum[c]; q += chan_sq[c]; }
    grp[g] = s;
    grp[groups + g] = q;
  }
  // ---- exchange the partials across the cluster ----
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    const uint32_t local = smem_u32(grp);
    for (int r = 0; r < kGNCluster; ++r) {
      uint32_t peer;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(local), "r"(r));
      float a, b;
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(a) : "r"(peer + (uint32_t)g * 4u));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(b) : "r"(peer + (uint32_t)(groups + g) * 4u));
      s += a;
      q += b;
    }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = s * inv;
    float var = q * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  // nobody may exit (or overwrite grp) while a peer still reads it; also publishes s_mean / s_rstd CTA-wide
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

  // ---- normalise this CTA's pixels ----
  if (!active) return;
  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    if (v >= cv) continue;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      const float ga = gamma[c] * s_rstd[g];
      a[j] = ga;
      b[j] = beta[c] - s_mean[g] * ga;
    }
    for (int pp = p_begin + tr; pp < p_end; pp += R) {
      const long long pix = (long long)n * hw + pp;
      float f[8];
      gn_load8(x1, x2, c1, c2, pix, v, f);
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
  const size_t smem = (size_t)(2 * C + kGNCThreads * 16) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  return launch_k("groupnorm (cluster)", gn_cluster_kernel<T>, dim3(kGNCluster, (unsigned)n), dim3(kGNCThreads), smem, st,
                  dim3(kGNCluster, 1, 1), (const T*)x1, (const T*)x2, gamma, beta, (__nv_bfloat16*)y, hw, c1, c2, groups, eps,
                  with_silu);
}

// ---------------------------------------------------------------------------------------------------------
// Single-launch, single-read GroupNorm for the denoiser's tensors: P CTAs per sample, every CTA keeps its pixel rows
// in shared memory. Pass 1 reduces them to per-group (sum, sumsq) partials in global memory; a software barrier among
// the P CTAs of the sample (arrival counter + generation word in the workspace; all CTAs of the grid are co-resident:
// grid <= 148, <= 96 KB of shared memory) publishes them; every CTA then folds the P partials in rank order
// (deterministic) and normalises its rows straight from shared memory. HBM/L2 traffic: 1 read + 1 write of the tensor.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGNGThreads = 256;
constexpr int kGNGMaxSmem = 96 * 1024;
constexpr int kGNGSyncBytes = 256;  // head of the workspace: {count, generation} per sample, zero-initialised once

__device__ __forceinline__ void copy8_raw(__nv_bfloat16* dst, const __nv_bfloat16* src) {
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
}
__device__ __forceinline__ void copy8_raw(float* dst, const float* src) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  *reinterpret_cast<float4*>(dst) = a;
  *reinterpret_cast<float4*>(dst + 4) = b;
}

template <typename T>
__global__ void __launch_bounds__(kGNGThreads)
gn_grid_kernel(const T* __restrict__ x1, const T* __restrict__ x2, const float* __restrict__ gamma,
               const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ part,
               unsigned int* __restrict__ sync, int hw, int c1, int c2, int groups, float eps, int with_silu, int ppc) {
  extern __shared__ uint8_t gsm[];
  __shared__ float s_mean[64], s_rstd[64];
  const int trc = trace_start(3);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
  const int C = c1 + c2;
  const int cv = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y, rank = blockIdx.x, P = gridDim.x;
  const int p_begin = rank * ppc;
  const int p_end = min(hw, p_begin + ppc);
  T* data = reinterpret_cast<T*>(gsm);                                   // [ppc][C] raw rows
  float* chan_sum = reinterpret_cast<float*>(gsm + (size_t)ppc * C * sizeof(T));  // [C]
  float* chan_sq = chan_sum + C;                                         // [C]
  float* red = chan_sq + C;                                              // [R][cols][16]
  const int cols = cv < kGNGThreads ? cv : kGNGThreads;
  const int R = kGNGThreads / cols;
  const int tr = threadIdx.x / cols, tv = threadIdx.x % cols;
  const bool active = threadIdx.x < R * cols;

  // ---- pass 1: global -> shared memory, per-channel sums ----
  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active && v < cv) {
      const int c = v * 8;
      const T* src = c < c1 ? x1 + c : x2 + (c - c1);
      const int ld = c < c1 ? c1 : c2;
      for (int pp = p_begin + tr; pp < p_end; pp += R) {
        T* d = data + (size_t)(pp - p_begin) * C + c;
        copy8_raw(d, src + ((long long)n * hw + pp) * ld);
        float f[8];
        load8(d, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
      }
    }
    if (active) {
      float* dst = red + ((size_t)tr * cols + tv) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = s[j]; dst[8 + j] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < cols && v < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
      for (int r = 0; r < R; ++r) {
        const float* src = red + ((size_t)r * cols + tv) * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += src[j]; q[j] += src[8 + j]; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { chan_sum[v * 8 + j] = s[j]; chan_sq[v * 8 + j] = q[j]; }
    }
    __syncthreads();
  }
  float gs = 0.f, gq = 0.f;
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { gs += chan_sum[c]; gq += chan_sq[c]; }
    if (P > 1) {
      float* out = part + (((size_t)n * P + rank) * groups + g) * 2;
      __stcg(out, gs);
      __stcg(out + 1, gq);
    }
  }
  // ---- barrier among the P CTAs of this sample ----
  if (P > 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int* count = sync + 2 * n;
      volatile unsigned int* gen = sync + 2 * n + 1;
      const unsigned int g0 = *gen;  // read before arriving: the generation cannot advance until this CTA has arrived
      __threadfence();
      if (atomicAdd(count, 1u) == (unsigned int)(P - 1)) {
        *count = 0u;  // ready for the next launch (stream-ordered after this one)
        __threadfence();
        atomicAdd((unsigned int*)gen, 1u);
      } else {
        const long long t0 = clock64();
        while (*gen == g0) {
          if (clock64() - t0 > 4000000000LL) {
            printf("sdeo: groupnorm grid barrier timed out (block %d,%d)\n", (int)blockIdx.x, (int)blockIdx.y);
            __trap();
          }
        }
      }
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    if (P > 1) {
      gs = 0.f; gq = 0.f;
      for (int r = 0; r < P; ++r) {
        const float* in = part + (((size_t)n * P + r) * groups + g) * 2;
        gs += __ldcg(in);
        gq += __ldcg(in + 1);
      }
    }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = gs * inv;
    float var = gq * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  // ---- pass 2: normalise from shared memory ----
  if (!active) return;
  for (int vbase = 0; vbase < cv; vbase += cols) {
    const int v = vbase + tv;
    if (v >= cv) continue;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      const float ga = gamma[c] * s_rstd[g];
      a[j] = ga;
      b[j] = beta[c] - s_mean[g] * ga;
    }
    for (int pp = p_begin + tr; pp < p_end; pp += R) {
      float f[8];
      load8(data + (size_t)(pp - p_begin) * C + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = f[j] * a[j] + b[j];
        f[j] = with_silu ? silu_f(t) : t;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(y + ((long long)n * hw + pp) * C + v * 8) = o;
    }
  }
  trace_mark(trc, 3);
}

// Geometry of the grid variant: P CTAs per sample x ppc pixel rows each; false if the tensor does not fit.
static bool gn_grid_geometry(int n, int hw, int C, int elem, int* P, int* ppc, size_t* smem) {
  if (n > 16) return false;
  int per = 148 / n;
  if (per < 1) per = 1;
  int rows = (hw + per - 1) / per;
  if (rows < 4) rows = hw < 4 ? hw : 4;
  *ppc = rows;
  *P = (hw + rows - 1) / rows;
  *smem = (size_t)rows * C * elem + (size_t)(2 * C + kGNGThreads * 16) * sizeof(float);
  return *smem <= (size_t)kGNGMaxSmem;
}

template <typename T>
static int launch_gn_grid(const void* x1, const void* x2, const float* gamma, const float* beta, void* y, void* workspace,
                          int n, int hw, int c1, int c2, int groups, float eps, int with_silu, int P, int ppc, size_t smem,
                          cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_grid_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGNGMaxSmem);
    if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  unsigned int* sync = reinterpret_cast<unsigned int*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kGNGSyncBytes);
  return launch_k("groupnorm (grid)", gn_grid_kernel<T>, dim3((unsigned)P, (unsigned)n), dim3(kGNGThreads), smem, st,
                  dim3(1, 1, 1), (const T*)x1, (const T*)x2, gamma, beta, (__nv_bfloat16*)y, part, sync, hw, c1, c2, groups,
                  eps, with_silu, ppc);
}

static void gn_geometry(int n, int hw, int* chunks, int* ppc) {
  // about 1.5 CTAs per SM over the batch, and at least 16 pixels per CTA (these tensors are small: per-CTA fixed
  // costs and the apply kernel's per-CTA fold over the chunk partials dominate otherwise)
  int want = (148 * 3 / 2 + n - 1) / n;
  if (want < 1) want = 1;
  int p = (hw + want - 1) / want;
  if (p < 16) p = hw < 16 ? hw : 16;
  *ppc = p;
  *chunks = (hw + p - 1) / p;
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
  int chunks, ppc;
  gn_geometry(n, hw, &chunks, &ppc);
  size_t two_pass = (size_t)n * chunks * groups * 2 * sizeof(float);
  size_t grid = (size_t)n * 148 * groups * 2 * sizeof(float);
  return kGNGSyncBytes + (two_pass > grid ? two_pass : grid);
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
  // SDEO_GN_GRID=1: single-read variant (P CTAs per sample + software barrier). Measured on the two-stream step graph
  // it is SLOWER than the cluster variant below (4.90 vs 4.56 ms/step): its ~128 CTAs must all become resident before
  // any can pass the barrier, which serialises it against the other branch's 200 KB-smem GEMM CTAs. Off by default.
  {
    int P, ppc;
    size_t smem;
    const int elem = x_f32 ? 4 : 2;
    if (getenv("SDEO_GN_GRID") && workspace_bytes >= kGNGSyncBytes + (size_t)n * 148 * groups * 2 * sizeof(float) &&
        gn_grid_geometry(n, hw, C, elem, &P, &ppc, &smem)) {
      if (x_f32)
        return launch_gn_grid<float>(x1, x2, gamma, beta, y, workspace, n, hw, c1, c2, groups, eps, with_silu, P, ppc, smem,
                                     (cudaStream_t)stream);
      return launch_gn_grid<__nv_bfloat16>(x1, x2, gamma, beta, y, workspace, n, hw, c1, c2, groups, eps, with_silu, P, ppc,
                                           smem, (cudaStream_t)stream);
    }
  }
  workspace = reinterpret_cast<uint8_t*>(workspace) + kGNGSyncBytes;  // the head holds the grid variant's barrier words
  workspace_bytes = workspace_bytes > (size_t)kGNGSyncBytes ? workspace_bytes - kGNGSyncBytes : 0;
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
  gn_geometry(n, hw, &chunks, &ppc);
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
