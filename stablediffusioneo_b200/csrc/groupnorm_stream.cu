// GroupNorm(+Swish) on the TensorRT plugin's contract (fp16 NHWC "kHWC8" in / out, fp32 gamma / beta; GroupNormPlugin::enqueue,
// plugin/groupNormPlugin/groupNormPlugin.cpp:179-228) as ONE persistent, shared-memory-staged kernel.
//
// Layout: an NHWC sample is one contiguous byte range, so a TILE (ppc pixels x C channels) is a contiguous range too and
// moves with 1-D bulk TMA copies (cp.async.bulk), no tensor map. Each tile is visited twice:
//   visit 0 (statistics): HBM -> shared memory, per-group (sum, sum of squares) of the tile -> workspace; the CTA that
//           delivers the LAST tile of a sample folds the sample's partials in fixed order (deterministic) into (mean, rstd)
//           and raises the sample's ready flag;
//   visit 1 (apply): tile -> shared memory again (an L2 hit: the visit order below keeps the re-read distance at about one
//           sample + two tiles per SM), normalise + affine (+ Swish) in place, shared memory -> HBM with a bulk store.
// Visits are numbered in one global sequence ("tickets"): statistics of tile s, then apply of tile s - lag; CTA b takes
// tickets b, b + G, b + 2G, ... in order. An apply visit only ever waits for statistics visits with LOWER ticket numbers
// and a statistics visit never waits, so with the G <= #SM CTAs co-resident the lowest unfinished ticket can always run:
// no deadlock for any tile count. Loads run kGSBufs - 1 tiles ahead of the arithmetic (mbarrier per buffer), stores drain
// behind it (bulk groups): HBM traffic is 1 read + 1 write of the tensor (4 bytes per element) while the sample fits L2.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace sdeo {

constexpr int kGSThreads = 512;
constexpr int kGSBufs = 4;
constexpr int kGSSmemTotal = 227 * 1024 - 6 * 1024;  // dynamic shared memory budget (static arrays take the rest)
constexpr unsigned long long kL2EvictFirst = 0x12F0000000000000ull;  // createpolicy.fractional.L2::evict_first, fraction 1.0

__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, bool evict_first) {
  if (evict_first)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                     "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(kL2EvictFirst) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes, bool evict_first) {
  if (evict_first)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::
                     "l"(dst), "r"(smem_u32(src_smem)), "r"(bytes), "l"(kL2EvictFirst) : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                     "l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void h8_to_f(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Swish on a pair of normalised values. kMode 1: fp32, ex2.approx + rcp.approx per value (2 MUFU / value).
// kMode 2: the exponentials of the pair in ONE packed ex2.approx.f16x2 (1.5 MUFU / value): e carries fp16 rounding
// (2^-11 relative) into 1 / (1 + e), i.e. at most ~half an fp16 ulp on the result, which is stored as fp16 anyway.
template <int kMode>
__device__ __forceinline__ uint32_t swish_pack(float t0, float t1) {
  if (kMode == 0) return f2_to_h2(t0, t1);
  constexpr float kNegLog2e = -1.4426950408889634f;
  float e0, e1;
  if (kMode == 2) {
    // exponent clamped to 15.9: 2^15.9 < fp16 max (the packed ex2 would return +inf above 16, and 1/(1+inf) = 0 is the
    // right limit anyway; the clamp keeps the value finite for t < -11 where swish(t) ~ -1.8e-4 .. 0)
    const uint32_t m = f2_to_h2(fminf(t0 * kNegLog2e, 15.9f), fminf(t1 * kNegLog2e, 15.9f));
    uint32_t e;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(m));
    const float2 ef = __half22float2(*reinterpret_cast<const __half2*>(&e));
    e0 = ef.x;
    e1 = ef.y;
  } else {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0 * kNegLog2e));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1 * kNegLog2e));
  }
  return f2_to_h2(__fdividef(t0, 1.0f + e0), __fdividef(t1, 1.0f + e1));
}

struct GSGeom {
  int n, hw, C, groups, chunks, ppc, lag, tile_stride;  // tile_stride: bytes between tile buffers (multiple of 128)
};

// ticket j -> (visit, tile): tickets 0 .. lag-1 are statistics of tiles 0 .. lag-1; then pairs (statistics s, apply s - lag);
// then the last `lag` applies.
__host__ __device__ __forceinline__ void gs_decode(int j, int tiles, int lag, int* visit, int* tile) {
  if (j < lag) {
    *visit = 0;
    *tile = j;
    return;
  }
  const int pairs = 2 * (tiles - lag);
  const int r = j - lag;
  if (r < pairs) {
    const int s = lag + (r >> 1);
    *visit = r & 1;
    *tile = (r & 1) ? s - lag : s;
    return;
  }
  *visit = 1;
  *tile = (r - pairs) + tiles - lag;
}

template <int kMode>
__global__ void __launch_bounds__(kGSThreads, 1)
gn_stream_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __half* __restrict__ y, float* __restrict__ part_ws, float2* __restrict__ final_ws, int* __restrict__ flags,
                 GSGeom gm, float eps, int hints) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(128) unsigned char gs_smem[];
  __shared__ __align__(8) uint64_t bars[kGSBufs];
  __shared__ float2 s_mr[64];
  __shared__ float2 s_fold[kGSThreads];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int C = gm.C, hw = gm.hw, groups = gm.groups, chunks = gm.chunks, ppc = gm.ppc;
  const int cv = C / 8, cpg = C / groups;
  const int tiles = gm.n * chunks, total = 2 * tiles, G = gridDim.x;
  float* chan = reinterpret_cast<float*>(gs_smem + (size_t)kGSBufs * gm.tile_stride);  // [cv][16]: 8 sums, 8 sums of squares
  float* part = chan + 2 * C;                                                             // [R][cols][16]
  float* part2 = part + kGSThreads * 16;                                                  // [nparts][S], S * nparts <= 512
  int* arrived = flags;
  int* ready = flags + gm.n;

  const int cols = cv < kGSThreads ? cv : kGSThreads;
  const int R = kGSThreads / cols;
  const int tr = tid / cols, tv = tid % cols;
  const bool active = tid < R * cols;
  const int S = cols * 16;

  if (tid == 0) {
    for (int b = 0; b < kGSBufs; ++b) mbar_init(&bars[b], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int j, int buf) {  // thread 0: start the load of ticket j's tile into buffer buf
    int visit, tile;
    gs_decode(j, tiles, gm.lag, &visit, &tile);
    const int img = tile / chunks, ch = tile - img * chunks;
    const int p0 = ch * ppc;
    const int rows = min(ppc, hw - p0);
    const uint32_t bytes = (uint32_t)rows * (uint32_t)C * 2u;
    mbar_expect_tx(&bars[buf], bytes);
    bulk_load(gs_smem + (size_t)buf * gm.tile_stride, x + ((long long)img * hw + p0) * C, bytes, &bars[buf],
              hints && visit == 1);
  };

  if (tid == 0) {
    for (int i = 0; i < kGSBufs - 1; ++i) {
      const int j = blockIdx.x + i * G;
      if (j < total) issue(j, i);
    }
  }
  uint32_t phases = 0;
  int k = 0;
  for (int j = blockIdx.x; j < total; j += G, ++k) {
    const int buf = k % kGSBufs;
    int visit, tile;
    gs_decode(j, tiles, gm.lag, &visit, &tile);
    const int img = tile / chunks, ch = tile - img * chunks;
    const int p0 = ch * ppc;
    const int rows = min(ppc, hw - p0);
    __half* tp = reinterpret_cast<__half*>(gs_smem + (size_t)buf * gm.tile_stride);

    if (visit == 1) {
      // the sample's statistics: published by the CTA that delivered the sample's last statistics tile
      if (tid == 0) {
        if (ld_acquire(ready + img) == 0) {
          const long long t0 = clock64();
          uint32_t spins = 0;
          while (ld_acquire(ready + img) == 0) {
            __nanosleep(64);
            if ((++spins & 0xFFFu) == 0 && clock64() - t0 > 4000000000LL) {
              printf("sdeo: groupnorm_f16 waited too long for sample %d (block %d)\n", img, (int)blockIdx.x);
              __trap();
            }
          }
        }
      }
      __syncthreads();
      if (tid < groups) s_mr[tid] = __ldcg(final_ws + (size_t)img * groups + tid);
    }
    mbar_wait(&bars[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    if (visit == 1) __syncthreads();  // s_mr

    if (visit == 0) {
      // ---- per-channel (sum, sum of squares) of this tile ----
      for (int vbase = 0; vbase < cv; vbase += cols) {
        const int v = vbase + tv;
        float s[8], q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { s[u] = 0.f; q[u] = 0.f; }
        if (active && v < cv) {
          int pp = tr;
          for (; pp + R < rows; pp += 2 * R) {
            float f0[8], f1[8];
            h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C + v * 8), f0);
            h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)(pp + R) * C + v * 8), f1);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              s[u] += f0[u] + f1[u];
              q[u] += f0[u] * f0[u] + f1[u] * f1[u];
            }
          }
          if (pp < rows) {
            float f0[8];
            h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C + v * 8), f0);
#pragma unroll
            for (int u = 0; u < 8; ++u) { s[u] += f0[u]; q[u] += f0[u] * f0[u]; }
          }
        }
        if (active) {
          float4* dst = reinterpret_cast<float4*>(part + (size_t)tr * S + tv * 16);
          dst[0] = make_float4(s[0], s[1], s[2], s[3]);
          dst[1] = make_float4(s[4], s[5], s[6], s[7]);
          dst[2] = make_float4(q[0], q[1], q[2], q[3]);
          dst[3] = make_float4(q[4], q[5], q[6], q[7]);
        }
        __syncthreads();
        // fold the R pixel rows with every thread: scalar i = column * 16 + slot, rows split in nparts interleaved parts
        // (fixed order: deterministic)
        const int nparts = S >= kGSThreads ? 1 : min(R, kGSThreads / S);
        for (int idx = tid; idx < S * nparts; idx += kGSThreads) {
          const int i = idx % S, rp = idx / S;
          float acc = 0.f;
          for (int r = rp; r < R; r += nparts) acc += part[(size_t)r * S + i];
          if (nparts == 1) chan[vbase * 16 + i] = acc;
          else part2[rp * S + i] = acc;
        }
        if (nparts > 1) {
          __syncthreads();
          for (int i = tid; i < S; i += kGSThreads) {
            float acc = 0.f;
            for (int rp = 0; rp < nparts; ++rp) acc += part2[rp * S + i];
            chan[vbase * 16 + i] = acc;
          }
        }
        __syncthreads();
      }
      if (tid < groups) {
        float s = 0.f, q = 0.f;
        for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) {
          s += chan[(c >> 3) * 16 + (c & 7)];
          q += chan[(c >> 3) * 16 + 8 + (c & 7)];
        }
        *reinterpret_cast<float2*>(part_ws + ((size_t)tile * groups + tid) * 2) = make_float2(s, q);
        __threadfence();
      }
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(arrived + img, 1) == chunks - 1);
      __syncthreads();
      if (s_last) {
        // this CTA delivered the sample's last tile: fold the sample's partials in fixed order, publish (mean, rstd)
        __threadfence();
        const int L = kGSThreads / groups;
        const int g = tid % groups, lane = tid / groups;
        float s = 0.f, q = 0.f;
        if (lane < L) {
          for (int c2 = lane; c2 < chunks; c2 += L) {
            const float2 p2 = __ldcg(reinterpret_cast<const float2*>(part_ws + (((size_t)img * chunks + c2) * groups + g) * 2));
            s += p2.x;
            q += p2.y;
          }
        }
        s_fold[tid] = make_float2(s, q);
        __syncthreads();
        if (tid < groups) {
          s = 0.f;
          q = 0.f;
          for (int l = 0; l < L; ++l) {
            s += s_fold[l * groups + tid].x;
            q += s_fold[l * groups + tid].y;
          }
          const float inv = 1.0f / ((float)hw * (float)cpg);
          const float mean = s * inv;
          float var = q * inv - mean * mean;
          var = var < 0.f ? 0.f : var;
          final_ws[(size_t)img * groups + tid] = make_float2(mean, rsqrtf(var + eps));
          __threadfence();
        }
        __syncthreads();
        if (tid == 0) st_release(ready + img, 1);
      }
    } else {
      // ---- normalise + affine (+ Swish) in place ----
      for (int vbase = 0; vbase < cv; vbase += cols) {
        const int v = vbase + tv;
        if (!active || v >= cv) continue;
        float a[8], b[8];
        {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
          a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w; a[4] = g1.x; a[5] = g1.y; a[6] = g1.z; a[7] = g1.w;
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
          int g = (v * 8) / cpg, r = v * 8 - g * cpg;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float2 mr = s_mr[g];
            a[u] *= mr.y;
            b[u] -= mr.x * a[u];
            if (++r == cpg) { r = 0; ++g; }
          }
        }
        int pp = tr;
        for (; pp + R < rows; pp += 2 * R) {
          uint4* p0v = reinterpret_cast<uint4*>(tp + (size_t)pp * C + v * 8);
          uint4* p1v = reinterpret_cast<uint4*>(tp + (size_t)(pp + R) * C + v * 8);
          float f0[8], f1[8];
          h8_to_f(*p0v, f0);
          h8_to_f(*p1v, f1);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            f0[u] = f0[u] * a[u] + b[u];
            f1[u] = f1[u] * a[u] + b[u];
          }
          uint4 o0, o1;
          o0.x = swish_pack<kMode>(f0[0], f0[1]); o0.y = swish_pack<kMode>(f0[2], f0[3]);
          o0.z = swish_pack<kMode>(f0[4], f0[5]); o0.w = swish_pack<kMode>(f0[6], f0[7]);
          o1.x = swish_pack<kMode>(f1[0], f1[1]); o1.y = swish_pack<kMode>(f1[2], f1[3]);
          o1.z = swish_pack<kMode>(f1[4], f1[5]); o1.w = swish_pack<kMode>(f1[6], f1[7]);
          *p0v = o0;
          *p1v = o1;
        }
        if (pp < rows) {
          uint4* p0v = reinterpret_cast<uint4*>(tp + (size_t)pp * C + v * 8);
          float f0[8];
          h8_to_f(*p0v, f0);
#pragma unroll
          for (int u = 0; u < 8; ++u) f0[u] = f0[u] * a[u] + b[u];
          uint4 o0;
          o0.x = swish_pack<kMode>(f0[0], f0[1]); o0.y = swish_pack<kMode>(f0[2], f0[3]);
          o0.z = swish_pack<kMode>(f0[4], f0[5]); o0.w = swish_pack<kMode>(f0[6], f0[7]);
          *p0v = o0;
        }
      }
      fence_proxy_async_smem();  // the in-place results are read by the bulk store (async proxy)
    }
    __syncthreads();  // every thread is done with buffer `buf` (and with the one the next load lands in)
    if (tid == 0) {
      if (visit == 1)
        bulk_store(y + ((long long)img * hw + p0) * C, tp, (uint32_t)rows * (uint32_t)C * 2u, hints != 0);
      bulk_commit();        // one group per ticket (empty for statistics visits)
      bulk_wait_read<1>();  // the store of the PREVIOUS ticket has left its buffer: that buffer takes the next load
      const int jn = j + (kGSBufs - 1) * G;
      if (jn < total) issue(jn, (k + kGSBufs - 1) % kGSBufs);
    }
  }
  if (tid == 0) bulk_wait_all();
}

// tile geometry: `chunks` tiles of `ppc` pixels per sample
static int gs_geometry(int n, int hw, int c, int sms, GSGeom* g, size_t* smem) {
  const size_t aux = ((size_t)2 * c + kGSThreads * 16 + 512) * sizeof(float);
  if (aux + (size_t)kGSBufs * c * 2 > (size_t)kGSSmemTotal) return -1;
  const size_t tile_cap = ((kGSSmemTotal - aux) / kGSBufs) & ~(size_t)127;
  const int ppc_cap = (int)(tile_cap / ((size_t)c * 2));
  int ppc_min = (int)((4096 + (size_t)c * 2 - 1) / ((size_t)c * 2));  // >= 4 KB per tile
  if (ppc_min > ppc_cap) ppc_min = ppc_cap;
  // small tensors: about one tile per SM over the batch; large ones: tiles as big as the buffers allow, evenly sized
  int want = (sms + n - 1) / n;
  int ppc = (hw + want - 1) / want;
  if (ppc < ppc_min) ppc = ppc_min;
  if (ppc > ppc_cap) {
    const int ch = (hw + ppc_cap - 1) / ppc_cap;
    ppc = (hw + ch - 1) / ch;
  }
  if (ppc > hw) ppc = hw;
  g->n = n; g->hw = hw; g->C = c;
  g->ppc = ppc;
  g->chunks = (hw + ppc - 1) / ppc;
  g->tile_stride = (int)((((size_t)ppc * c * 2) + 127) & ~(size_t)127);
  *smem = (size_t)kGSBufs * g->tile_stride + aux;
  return 0;
}

// geometry + grid + apply lag of one call; non-zero when the streamed kernel cannot take the shape
static int gs_plan(int n, int hw, int c, int groups, int sms, int lag_env, GSGeom* g, size_t* smem, int* grid) {
  if (gs_geometry(n, hw, c, sms, g, smem)) return -1;
  const long long tiles = (long long)n * g->chunks;
  if (tiles > (1 << 29)) return -1;
  g->groups = groups;
  const int G = (int)(tiles < sms ? tiles : sms);
  // apply lag: a sample's statistics tiles plus two rounds of the grid, so that the statistics an apply visit needs are
  // (almost always) complete when its turn comes
  long long lag = lag_env >= 0 ? (long long)lag_env : (long long)g->chunks + 2LL * G;
  if (lag < g->chunks) lag = g->chunks;  // ordering requirement: apply(t) after statistics of every tile of t's sample
  if (lag > tiles) lag = tiles;
  g->lag = (int)lag;
  *grid = G;
  return 0;
}

static int gs_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

int groupnorm_f16_two_pass(const void* x, const float* gamma, const float* beta, void* y, int32_t n, int32_t hw, int32_t c,
                           int32_t groups, float eps, int32_t with_silu, void* workspace, size_t workspace_bytes, void* stream);
size_t groupnorm_two_pass_workspace_bytes(int32_t n, int32_t hw, int32_t groups);

}  // namespace sdeo

using namespace sdeo;

extern "C" size_t sdeo_groupnorm_f16_workspace_bytes(int32_t n, int32_t hw, int32_t c, int32_t groups) {
  if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0) return 0;
  size_t two_pass = groupnorm_two_pass_workspace_bytes(n, hw, groups);
  GSGeom g;
  size_t smem;
  // the tile count depends on the SM count only for small tensors, where fewer SMs mean fewer tiles: 148 is an upper bound
  // for every part with <= 148 SMs; parts with more SMs get the exact figure
  int sms = 148;
  int dev_sms = 0, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
      dev_sms > sms)
    sms = dev_sms;
  (void)cudaGetLastError();
  if (gs_geometry(n, hw, c, sms, &g, &smem)) return two_pass;
  const size_t stream_bytes = (size_t)n * g.chunks * groups * 2 * sizeof(float) + (size_t)n * groups * sizeof(float2) +
                              ((size_t)2 * n + 4) * sizeof(int);
  return stream_bytes > two_pass ? stream_bytes : two_pass;
}

// Host-side view of the schedule (tests): plan[0..5] = tiles per sample, pixels per tile, apply lag (tiles),
// grid size, dynamic shared memory bytes, tile buffer stride; returns non-zero when the shape falls back to two launches.
extern "C" int sdeo_groupnorm_f16_plan(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t sms, int32_t* plan) {
  if (!plan || n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || c % 8 != 0) return set_error(SDEO_EINVAL, "groupnorm_f16_plan: bad argument");
  GSGeom g;
  size_t smem = 0;
  int G = 0;
  if (gs_plan(n, hw, c, groups, sms > 0 ? sms : 148, -1, &g, &smem, &G)) return 1;
  plan[0] = g.chunks; plan[1] = g.ppc; plan[2] = g.lag; plan[3] = G; plan[4] = (int32_t)smem; plan[5] = g.tile_stride;
  return 0;
}
// ticket j of the visit sequence -> out[0] = visit (0 statistics, 1 apply), out[1] = tile
extern "C" void sdeo_groupnorm_f16_ticket(int32_t j, int32_t tiles, int32_t lag, int32_t* out) {
  int visit, tile;
  gs_decode(j, tiles, lag, &visit, &tile);
  out[0] = visit;
  out[1] = tile;
}

// x / y fp16 NHWC, gamma / beta fp32, one tensor, optional Swish (GroupNormPlugin::enqueue, groupNormPlugin.cpp:179-228;
// unlike groupNormKernel.cu:190-194, epsilon IS applied). SDEO_GN_F16_TWO_PASS=1 selects the two-launch grid variant,
// SDEO_GN_F16_SWISH=1|2 the Swish arithmetic (see swish_pack; default 2), SDEO_GN_F16_LAG the apply lag in tiles.
extern "C" int sdeo_groupnorm_nhwc_f16(const void* x, const float* gamma, const float* beta, void* y, int32_t n, int32_t hw,
                                       int32_t c, int32_t groups, float eps, int32_t with_silu, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!x || !gamma || !beta || !y || !workspace) return set_error(SDEO_EINVAL, "groupnorm_f16: null argument");
  if (n <= 0 || n > 65535 || hw <= 0 || groups <= 0 || groups > 64 || c % groups != 0 || c % 8 != 0)
    return set_error(SDEO_EINVAL, "groupnorm_f16: unsupported geometry (need C % groups == 0, C % 8 == 0, groups <= 64)");
  // tuning / A-B switches, read per call (a getenv is noise next to a launch)
  int two_pass = getenv("SDEO_GN_F16_TWO_PASS") ? 1 : 0, swish_mode = 2, lag_env = -1, hints_env = -1;
  {
    const char* e = getenv("SDEO_GN_F16_SWISH");
    if (e && atoi(e) == 1) swish_mode = 1;
    e = getenv("SDEO_GN_F16_LAG");
    if (e) lag_env = atoi(e);
    e = getenv("SDEO_GN_F16_HINTS");
    if (e) hints_env = atoi(e);
  }
  GSGeom g;
  size_t smem = 0;
  int G = 0;
  if (two_pass || gs_plan(n, hw, c, groups, gs_sm_count(), lag_env, &g, &smem, &G))
    return groupnorm_f16_two_pass(x, gamma, beta, y, n, hw, c, groups, eps, with_silu, workspace, workspace_bytes, stream);
  const long long tiles = (long long)n * g.chunks;
  const size_t part_bytes = (size_t)tiles * groups * 2 * sizeof(float);
  const size_t final_bytes = (size_t)n * groups * sizeof(float2);
  const size_t flag_bytes = ((size_t)2 * n + 4) * sizeof(int);
  if (workspace_bytes < part_bytes + final_bytes + flag_bytes)
    return set_error(SDEO_EINVAL, "groupnorm_f16: workspace too small (sdeo_groupnorm_f16_workspace_bytes)");
  float* part_ws = (float*)workspace;
  float2* final_ws = (float2*)((char*)workspace + part_bytes);
  int* flags = (int*)((char*)workspace + part_bytes + final_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(flags, 0, flag_bytes, st) != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error(SDEO_ECUDA, "groupnorm_f16: cudaMemsetAsync failed");
  }
  // output / re-read eviction hints only when the tensor cannot stay in L2 for its consumer anyway
  const int hints = hints_env >= 0 ? hints_env : ((long long)n * hw * c * 4 > (96LL << 20) ? 1 : 0);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
    if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  const dim3 one(1, 1, 1);
  const int mode = with_silu ? swish_mode : 0;
  auto fn = mode == 0 ? gn_stream_kernel<0> : (mode == 1 ? gn_stream_kernel<1> : gn_stream_kernel<2>);
  return launch_k("groupnorm_f16 (streamed)", fn, dim3((unsigned)G), dim3(kGSThreads), smem, st, one, (const __half*)x, gamma,
                  beta, (__half*)y, part_ws, final_ws, flags, g, eps, hints);
}
