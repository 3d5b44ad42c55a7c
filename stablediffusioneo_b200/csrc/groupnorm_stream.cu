// GroupNorm(+Swish) on the TensorRT plugin's contract (fp16 NHWC "kHWC8" in / out, fp32 gamma / beta; GroupNormPlugin::enqueue,
// plugin/groupNormPlugin/groupNormPlugin.cpp:179-228) as ONE persistent, shared-memory-staged kernel.
//
// Layout: an NHWC sample is one contiguous byte range, so a TILE (ppc pixels x C channels) is a contiguous range too and
// moves with 1-D bulk TMA copies (cp.async.bulk), no tensor map. Each tile is visited twice:
//   visit 0 (statistics): HBM -> shared memory, per-group (sum, sum of squares) of the tile, accumulated in registers over
//           the tiles this CTA takes of one sample and PUBLISHED once per (CTA, sample) as self-validating 64-bit slots
//           (see slot_publish): no fence, no atomic, no flag on the arithmetic threads' path;
//   visit 1 (apply): tile -> shared memory again (an L2 hit while `lag` tiles of input + output fit L2), normalise +
//           affine (+ Swish) in place, shared memory -> HBM with a bulk store. The control warp folds the sample's slots
//           (<= one per CTA) in slot order into (mean, rstd) the first time the CTA meets the sample: deterministic.
// Visits are grouped in UNITS: unit u = statistics of tile u (u < tiles) then apply of tile u - lag (u >= lag); CTA b takes
// units b, b + G, b + 2G, ... in order, so every CTA carries the same mix of (light) statistics and (heavy) apply visits and
// the re-read distance stays at `lag` tiles. An apply visit only waits for statistics visits in LOWER units (lag >= tiles per
// sample) and a statistics visit never waits, so with the G <= #SM CTAs co-resident the lowest unfinished visit can always
// run: no deadlock for any tile count. One control warp per CTA recycles the kGSBufs tile buffers (empty / full mbarriers),
// issues the loads bufs - 1 visits ahead and supplies the sample's (mean, rstd) row; the 512 arithmetic threads never
// wait on global memory.
// Stores drain behind the arithmetic (bulk groups): HBM traffic is 1 read + 1 write of the tensor (4 bytes per element)
// while `lag` tiles of input and output fit L2.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace sdeo {

constexpr int kGSThreads = 512;             // arithmetic threads
constexpr int kGSAll = kGSThreads + 32;     // + one control warp (loads, buffer recycling, ready flags)
constexpr int kGSTeam = 16;                 // partial slots per team in the tree fold (one batch of loads)
constexpr int kGSMaxBufs = 8;               // tile buffers per CTA: GSGeom::bufs of them are used
constexpr int kGSSmemTotal = 227 * 1024 - 7 * 1024;  // dynamic shared memory budget (static arrays take the rest)
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
  int n, hw, C, groups, chunks, ppc, lag, bufs, two_level, tree, ngroups, tile_stride;  // tile_stride: bytes between tile buffers (multiple of 128)
};

// A CTA's position in the visit sequence: unit u, visit v (0 statistics of tile u, 1 apply of tile u - lag).
struct GSIter {
  int u, v;
};
// moves `it` to the first existing visit at or after (u, v) among the units u, u + G, ...; false when the CTA is done
__host__ __device__ __forceinline__ bool gs_settle(GSIter& it, int tiles, int lag, int G) {
  while (it.u < tiles + lag) {
    if (it.v == 0) {
      if (it.u < tiles) return true;
      it.v = 1;
    }
    if (it.u >= lag) return true;
    it.u += G;
    it.v = 0;
  }
  return false;
}
__host__ __device__ __forceinline__ bool gs_next(GSIter& it, int tiles, int lag, int G) {
  if (it.v == 0) {
    it.v = 1;
  } else {
    it.u += G;
    it.v = 0;
  }
  return gs_settle(it, tiles, lag, G);
}

// One published partial: (sum, sum of squares) of one group over the tiles ONE CTA took of one sample, as one 64-bit word
// (a naturally aligned 64-bit scalar store is single-copy atomic). The call presets every slot to all-ones; a slot whose
// high word is still 0xFFFFFFFF has not been written. Self-validating slots need no fence, no counter and no flag: the
// writer's other stores never have to be ordered against this one.
__device__ __forceinline__ void slot_publish(unsigned long long* p, float s, float q) {
  uint32_t hi = __float_as_uint(q);
  if (hi == 0xFFFFFFFFu) hi = 0x7FFFFFFFu;  // a NaN sum keeps meaning "NaN", never "empty"
  const unsigned long long v = ((unsigned long long)hi << 32) | (unsigned long long)__float_as_uint(s);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long slot_peek(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int kMode>
__global__ void __launch_bounds__(kGSAll, 1)
gn_stream_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __half* __restrict__ y, unsigned long long* __restrict__ slots, GSGeom gm, float eps, int hints) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(128) unsigned char gs_smem[];
  __shared__ __align__(8) uint64_t full[kGSMaxBufs];   // tile (and, for an apply visit, its (mean, rstd) row) has landed
  __shared__ __align__(8) uint64_t empty[kGSMaxBufs];  // the arithmetic threads and the bulk store are done with the buffer
  __shared__ float2 s_mr[kGSMaxBufs][32];
  const int tid = threadIdx.x;
  const int C = gm.C, hw = gm.hw, groups = gm.groups, chunks = gm.chunks, ppc = gm.ppc, lag = gm.lag;
  const int cv = C / 8, cpg = C / groups;
  const int tiles = gm.n * chunks, G = gridDim.x, kGSBufs = gm.bufs;
  const int P = chunks < gm.ngroups * G ? chunks : gm.ngroups * G;  // partials per sample: one per (CTA, thread group) with tiles of it

  if (tid == 0) {
    for (int b = 0; b < kGSBufs; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&empty[b], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= kGSThreads) {
    // ---- control warp: buffer recycling, tile loads, the sample's (mean, rstd) row for apply visits ----
    const int lane = tid - kGSThreads;
    GSIter it{(int)blockIdx.x, 0};
    int k = 0, cur_img = -1, kg0 = 0, kg1 = 0;
    const int bpg = kGSBufs / gm.ngroups;  // buffers per thread group
    float2 my_mr = make_float2(0.f, 1.f);  // lane = group: (mean, rstd) of sample cur_img
    // lane = group: sums `count` slots (stride `groups` words) in slot order - the same order wherever it runs: deterministic;
    // spins on slots that are not written yet (their writers never wait on anything this warp holds back)
    auto fold_slots = [&](const unsigned long long* first, int count, int what) -> float2 {
      float fs = 0.f, fq = 0.f;
      if (lane < groups) {
        const unsigned long long* base = first + lane;
        const long long t0 = clock64();
        for (int j0 = 0; j0 < count; j0 += 16) {
          unsigned long long v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = j0 + i < count ? slot_peek(base + (size_t)(j0 + i) * groups) : 0ull;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (j0 + i >= count) break;
            uint32_t spins = 0;
            while ((uint32_t)(v[i] >> 32) == 0xFFFFFFFFu) {
              __nanosleep(32);
              if ((++spins & 0xFFFu) == 0 && clock64() - t0 > 4000000000LL) {
                printf("sdeo: groupnorm_f16 waited too long for slot %d of %d (kind %d, block %d)\n", j0 + i, count, what, (int)blockIdx.x);
                __trap();
              }
              v[i] = slot_peek(base + (size_t)(j0 + i) * groups);
            }
            fs += __uint_as_float((uint32_t)v[i]);
            fq += __uint_as_float((uint32_t)(v[i] >> 32));
          }
        }
      }
      return make_float2(fs, fq);
    };
    auto finish = [&](float2 sums) -> float2 {  // (sum, sum of squares) of a whole sample -> (mean, rstd)
      const float inv = 1.0f / ((float)hw * (float)cpg);
      const float mean = sums.x * inv;
      float var = sums.y * inv - mean * mean;
      var = var < 0.f ? 0.f : var;
      return make_float2(mean, rsqrtf(var + eps));
    };
    // Slot areas: [n][P] partials, [n] (mean, rstd) rows ("finals"), [n][M] team sums ("mids", tree fold only).
    unsigned long long* finals = slots + (size_t)gm.n * P * groups;
    const int M = (P + kGSTeam - 1) / kGSTeam;  // teams of kGSTeam partial slots
    unsigned long long* mids = finals + (size_t)gm.n * groups;
    // sample img's sums: straight from the P partials, or (tree) from the M team sums other CTAs publish
    auto sample_sums = [&](int img) -> float2 {
      return gm.tree ? fold_slots(mids + (size_t)img * M * groups, M, 1) : fold_slots(slots + (size_t)img * P * groups, P, 0);
    };
    int next_fold = (int)blockIdx.x;  // folder duty: samples b, b + G, ... of CTA b
    int next_team = (int)blockIdx.x;  // team duty (tree fold): (sample, team) pairs b, b + G, ... in sample-major order
    for (bool ok = gs_settle(it, tiles, lag, G); ok; ok = gs_next(it, tiles, lag, G), ++k) {
      // Each thread group owns bufs / ng buffers and walks them with its own visit counter: a barrier is then waited on by ONE
      // group, phase after phase (a group that skipped the other's phases could not use parity waits).
      const int vg = ((it.u - (int)blockIdx.x) / G) % gm.ngroups;
      const int kgv = vg ? kg1++ : kg0++;
      const int buf = vg * bpg + kgv % bpg, use = kgv / bpg;
      // Folder duty (two-level fold): sample f is folded by CTA f mod G when that CTA first reaches a unit >= (f + 1) * chunks,
      // i.e. after every statistics tile of f in unit order and (lag >= chunks + G) before every apply visit of f.
      // Tree fold (single-round regime, where the fold's latency is exposed): team duties first - they wait on partials only -
      // then the folder sums M team slots instead of P partials: 2 dependent batches of loads instead of P / 16.
      while (gm.tree && next_team < gm.n * M && it.u >= (next_team / M + 1) * chunks && !(hints & 0x100)) {
        const int f = next_team / M, m = next_team - f * M;
        const int cnt = min(kGSTeam, P - m * kGSTeam);
        const float2 sums = fold_slots(slots + ((size_t)f * P + (size_t)m * kGSTeam) * groups, cnt, 2);
        if (lane < groups) slot_publish(mids + ((size_t)f * M + m) * groups + lane, sums.x, sums.y);
        next_team += G;
      }
      while (gm.two_level && next_fold < gm.n && it.u >= (next_fold + 1) * chunks && !(hints & 0x100)) {
        const float2 mr = finish(sample_sums(next_fold));
        if (lane < groups) slot_publish(finals + (size_t)next_fold * groups + lane, mr.x, mr.y);
        next_fold += G;
      }
      if (use > 0) mbar_wait(&empty[buf], (uint32_t)(use - 1) & 1u);
      const int tile = it.v ? it.u - lag : it.u;
      const int img = tile / chunks, ch = tile - img * chunks;
      const int p0 = ch * ppc;
      const int rows = min(ppc, hw - p0);
      const uint32_t bytes = (uint32_t)rows * (uint32_t)C * 2u;
      // the copy may complete before the expect_tx below: the phase cannot close while this warp's arrival is pending
      if (lane == 0)
        bulk_load(gs_smem + (size_t)buf * gm.tile_stride, x + ((long long)img * hw + p0) * C, bytes, &full[buf],
                  (hints & 1) && it.v == 1);
      if (it.v == 1) {
        if (img != cur_img && !(hints & 0x100)) {
          if (!gm.two_level) {
            my_mr = finish(fold_slots(slots + (size_t)img * P * groups, P, 0));  // every CTA folds for itself
          } else if (lane < groups) {
            // the sample's folder CTA published (mean, rstd) in a lower unit than this one (lag >= chunks + G)
            const unsigned long long* fp = finals + (size_t)img * groups + lane;
            unsigned long long v = slot_peek(fp);
            const long long t0 = clock64();
            uint32_t spins = 0;
            while ((uint32_t)(v >> 32) == 0xFFFFFFFFu) {
              __nanosleep(32);
              if ((++spins & 0xFFFu) == 0 && clock64() - t0 > 4000000000LL) {
                printf("sdeo: groupnorm_f16 waited too long for the statistics of sample %d (block %d)\n", img, (int)blockIdx.x);
                __trap();
              }
              v = slot_peek(fp);
            }
            my_mr = make_float2(__uint_as_float((uint32_t)v), __uint_as_float((uint32_t)(v >> 32)));
          }
          cur_img = img;
        }
        if (lane < groups) s_mr[buf][lane] = my_mr;
        __syncwarp();
      }
      if (lane == 0) mbar_expect_tx(&full[buf], bytes);  // arrive (releases the s_mr row) + expect the tile's bytes
    }
    return;
  }

  // ---- arithmetic threads: ng groups of nthr threads, each with its own named barrier and reduction scratch; group g takes
  //      every ng-th unit of the CTA, so one group's barriers and hand-overs hide behind the other's arithmetic ----
  const int ng = gm.ngroups, nthr = kGSThreads / ng;
  const int grp = tid / nthr, t = tid - grp * nthr;
  const uint32_t bar_id = 1u + (uint32_t)grp;
  float* aux = reinterpret_cast<float*>(gs_smem + (size_t)kGSBufs * gm.tile_stride);
  float* chan = aux + (size_t)grp * 2 * C;                         // [cv][16]: 8 sums, 8 sums of squares
  float* part = aux + (size_t)ng * 2 * C + (size_t)grp * nthr * 16;  // [R][cols][16]
  float* part2 = aux + (size_t)ng * 2 * C + kGSThreads * 16 + (size_t)grp * nthr;  // [nparts][S], S * nparts <= nthr
  const int cols = cv;  // <= nthr (host check)
  const int R = nthr / cols;
  const int tr = t / cols, tv = t % cols;
  const bool active = t < R * cols;
  const int S = cols * 16;
  const int nparts = S >= nthr ? 1 : min(R, nthr / S);
  const int gfirst = (tv * 8) / cpg, rfirst = tv * 8 - gfirst * cpg;
  const int W = ng * G;  // distance between two statistics tiles of one group

  GSIter it{(int)blockIdx.x, 0};
  int k = 0;
  bool prev_apply = false;
  int prev_buf = 0, kg = 0;
  const int bpg = kGSBufs / ng;
  float acc_s = 0.f, acc_q = 0.f;  // threads 0 .. groups-1: this CTA's running (sum, sum of squares) of the current sample
  for (bool ok = gs_settle(it, tiles, lag, G); ok; ok = gs_next(it, tiles, lag, G), ++k) {
    if (((it.u - (int)blockIdx.x) / G) % ng != grp) continue;  // the other group's unit
    const int buf = grp * bpg + kg % bpg;  // this group's own buffers, walked with its own visit counter (as the control warp does)
    const uint32_t parity = (uint32_t)(kg / bpg) & 1u;
    ++kg;
    const int tile = it.v ? it.u - lag : it.u;
    const int img = tile / chunks, ch = tile - img * chunks;
    const int p0 = ch * ppc;
    const int rows = min(ppc, hw - p0);
    __half* tp = reinterpret_cast<__half*>(gs_smem + (size_t)buf * gm.tile_stride);

    if (it.v == 0) {
      mbar_wait(&full[buf], parity);
      // ---- per-channel (sum, sum of squares) of this tile ----
      float s[8], q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { s[u] = 0.f; q[u] = 0.f; }
      if (active && !(hints & 0x1000)) {
        int pp = tr;
        for (; pp + R < rows; pp += 2 * R) {
          float f0[8], f1[8];
          h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C + tv * 8), f0);
          h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)(pp + R) * C + tv * 8), f1);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            s[u] += f0[u] + f1[u];
            q[u] += f0[u] * f0[u] + f1[u] * f1[u];
          }
        }
        if (pp < rows) {
          float f0[8];
          h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C + tv * 8), f0);
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
      bar_sync(bar_id, nthr);
      if (t == 0) mbar_arrive(&empty[buf]);  // the tile itself is no longer needed
      // fold the R pixel rows with every thread: scalar i = column * 16 + slot, rows split in nparts interleaved parts
      // (fixed order: deterministic)
      for (int idx = t; idx < S * nparts; idx += nthr) {
        const int i = idx % S, rp = idx / S;
        float acc = 0.f;
        for (int r = rp; r < R; r += nparts) acc += part[(size_t)r * S + i];
        if (nparts == 1) chan[i] = acc;
        else part2[rp * S + i] = acc;
      }
      if (nparts > 1) {
        bar_sync(bar_id, nthr);
        for (int i = t; i < S; i += nthr) {
          float acc = 0.f;
          for (int rp = 0; rp < nparts; ++rp) acc += part2[rp * S + i];
          chan[i] = acc;
        }
      }
      bar_sync(bar_id, nthr);
      if (t < groups) {
        float gs = 0.f, gq = 0.f;
        for (int c = t * cpg; c < (t + 1) * cpg; ++c) {
          gs += chan[(c >> 3) * 16 + (c & 7)];
          gq += chan[(c >> 3) * 16 + 8 + (c & 7)];
        }
        acc_s += gs;
        acc_q += gq;
        // this group's next statistics tile is it.u + W: publish when that one belongs to another sample (or does not exist).
        // Slot (it.u - first tile of the sample) mod W is the same for every tile this group takes of the sample, and the
        // slots 0 .. min(chunks, W) - 1 of a sample are each written exactly once.
        const int nxt = it.u + W;
        if (nxt >= tiles || nxt / chunks != img) {
          const int j = (it.u - img * chunks) % W;
          slot_publish(slots + ((size_t)img * P + j) * groups + t, acc_s, acc_q);
          acc_s = 0.f;
          acc_q = 0.f;
        }
      }
    } else {
      // ---- normalise + affine (+ Swish) in place ----
      float a[8], b[8];
      if (active) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + tv * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + tv * 8) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + tv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + tv * 8) + 1);
        a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w; a[4] = g1.x; a[5] = g1.y; a[6] = g1.z; a[7] = g1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      }
      mbar_wait(&full[buf], parity);
      if (active && !(hints & 0x800)) {
        int g = gfirst, r = rfirst;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float2 mr = s_mr[buf][g];
          a[u] *= mr.y;
          b[u] -= mr.x * a[u];
          if (++r == cpg) { r = 0; ++g; }
        }
        int pp = tr;
        for (; pp + R < rows; pp += 2 * R) {
          uint4* p0v = reinterpret_cast<uint4*>(tp + (size_t)pp * C + tv * 8);
          uint4* p1v = reinterpret_cast<uint4*>(tp + (size_t)(pp + R) * C + tv * 8);
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
          uint4* p0v = reinterpret_cast<uint4*>(tp + (size_t)pp * C + tv * 8);
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
      bar_sync(bar_id, nthr);
      if (t == 0 && !(hints & 0x400)) bulk_store(y + ((long long)img * hw + p0) * C, tp, (uint32_t)rows * (uint32_t)C * 2u, (hints & 1) != 0);
    }
    if (t == 0) {
      bulk_commit();  // one bulk group per visit (empty for statistics visits)
      if (ng == 1) {
        bulk_wait_read<1>();  // the store of the PREVIOUS visit has left its buffer: hand that one back
        if (prev_apply) mbar_arrive(&empty[prev_buf]);
      } else if (it.v == 1) {
        // Two thread groups: this group's next visit is a whole unit of the other group away, and the control warp hands
        // buffers out in visit order, so the hand-back cannot wait for it (it would deadlock with few buffers). Wait for
        // the store to have read the buffer now; the other group's arithmetic goes on meanwhile.
        bulk_wait_read<0>();
        mbar_arrive(&empty[buf]);
      }
    }
    prev_apply = it.v == 1;
    prev_buf = buf;
  }
  if (t == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------------
// Resident variant, for samples that fit the shared memory of one thread-block cluster: cluster = sample, CTA rank r keeps
// pixel rows [r * rpc, (r + 1) * rpc) in shared memory for the whole call. HBM -> shared memory once (kRChunks bulk copies,
// statistics start when the first lands), per-group sums exchanged through distributed shared memory (every CTA folds the
// ranks in rank order: deterministic), normalise in place, shared memory -> HBM chunk by chunk. One read + one write of the
// tensor from anywhere, no workspace, no second launch.
// ---------------------------------------------------------------------------------------------------------------------
static int gs_env_int_fwd(const char* name, int dflt, int lo, int hi);
constexpr int kRChunks = 4;
struct RGeom {
  int n, hw, C, groups, cs, rpc, tile_bytes;
};
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float2 ld_cluster_f2(const void* local_smem, uint32_t rank) {
  uint32_t ra;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra) : "memory");
  return v;
}

template <int kMode>
__global__ void __launch_bounds__(kGSThreads, 1)
gn_resident_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   __half* __restrict__ y, RGeom gm, float eps) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(128) unsigned char gs_smem[];
  __shared__ __align__(8) uint64_t bars[kRChunks];
  __shared__ float2 s_group[32];
  __shared__ float2 s_mr[32];
  const int tid = threadIdx.x;
  const int C = gm.C, hw = gm.hw, groups = gm.groups;
  const int cv = C / 8, cpg = C / groups;
  const uint32_t rank = cluster_rank();
  const int img = blockIdx.x / gm.cs;
  const int r0 = (int)rank * gm.rpc;
  const int rows = max(0, min(gm.rpc, hw - r0));
  const int rch = (rows + kRChunks - 1) / kRChunks;  // rows per chunk
  __half* tile = reinterpret_cast<__half*>(gs_smem);
  float* chan = reinterpret_cast<float*>(gs_smem + gm.tile_bytes);  // [cv][16]
  float* part = chan + 2 * C;                                        // [R][cols][16]
  float* part2 = part + kGSThreads * 16;
  const int cols = cv, R = kGSThreads / cols;
  const int tr = tid / cols, tv = tid % cols;
  const bool active = tid < R * cols;
  const int S = cols * 16;
  const int nparts = S >= kGSThreads ? 1 : min(R, kGSThreads / S);

  if (tid == 0) {
    for (int c = 0; c < kRChunks; ++c) mbar_init(&bars[c], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long base = ((long long)img * hw + r0) * C;
  if (tid == 0) {
    for (int c = 0; c < kRChunks; ++c) {
      const int cr = min(rch, rows - c * rch);
      if (cr <= 0) break;
      const uint32_t bytes = (uint32_t)cr * (uint32_t)C * 2u;
      mbar_expect_tx(&bars[c], bytes);
      bulk_load(tile + (size_t)c * rch * C, x + base + (long long)c * rch * C, bytes, &bars[c], false);
    }
  }
  float a[8], b[8];
  if (active) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + tv * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + tv * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + tv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + tv * 8) + 1);
    a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w; a[4] = g1.x; a[5] = g1.y; a[6] = g1.z; a[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  // ---- statistics of this CTA's rows ----
  float s[8], q[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { s[u] = 0.f; q[u] = 0.f; }
  for (int c = 0; c < kRChunks; ++c) {
    const int cr = min(rch, rows - c * rch);
    if (cr <= 0) break;
    mbar_wait(&bars[c], 0);
    if (active) {
      const __half* tp = tile + (size_t)c * rch * C + tv * 8;
      int pp = tr;
      for (; pp + R < cr; pp += 2 * R) {
        float f0[8], f1[8];
        h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C), f0);
        h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)(pp + R) * C), f1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          s[u] += f0[u] + f1[u];
          q[u] += f0[u] * f0[u] + f1[u] * f1[u];
        }
      }
      if (pp < cr) {
        float f0[8];
        h8_to_f(*reinterpret_cast<const uint4*>(tp + (size_t)pp * C), f0);
#pragma unroll
        for (int u = 0; u < 8; ++u) { s[u] += f0[u]; q[u] += f0[u] * f0[u]; }
      }
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
  for (int idx = tid; idx < S * nparts; idx += kGSThreads) {
    const int i = idx % S, rp = idx / S;
    float acc = 0.f;
    for (int r = rp; r < R; r += nparts) acc += part[(size_t)r * S + i];
    if (nparts == 1) chan[i] = acc;
    else part2[rp * S + i] = acc;
  }
  if (nparts > 1) {
    __syncthreads();
    for (int i = tid; i < S; i += kGSThreads) {
      float acc = 0.f;
      for (int rp = 0; rp < nparts; ++rp) acc += part2[rp * S + i];
      chan[i] = acc;
    }
  }
  __syncthreads();
  {
    // per-group sums: warp w takes groups w, w + 16; lanes stride the group's channels, then a fixed shuffle tree
    const int warp = tid >> 5, lane = tid & 31;
    for (int g = warp; g < groups; g += kGSThreads / 32) {
      float gs = 0.f, gq = 0.f;
      for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
        gs += chan[(c >> 3) * 16 + (c & 7)];
        gq += chan[(c >> 3) * 16 + 8 + (c & 7)];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(0xFFFFFFFFu, gs, o);
        gq += __shfl_xor_sync(0xFFFFFFFFu, gq, o);
      }
      if (lane == 0) s_group[g] = make_float2(gs, gq);
    }
  }
  cluster_sync_all();  // every rank's s_group is written
  if (tid < groups) {
    float2 p2[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)  // all remote loads in flight at once (cluster size <= 16)
      p2[r] = r < gm.cs ? ld_cluster_f2(&s_group[tid], (uint32_t)r) : make_float2(0.f, 0.f);
    float fs = 0.f, fq = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {  // rank order, the same in every CTA
      fs += p2[r].x;
      fq += p2[r].y;
    }
    const float inv = 1.0f / ((float)hw * (float)cpg);
    const float mean = fs * inv;
    float var = fq * inv - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mr[tid] = make_float2(mean, rsqrtf(var + eps));
  }
  // this CTA has read its peers' sums: arrive now, wait at the very end (a peer's shared memory must outlive the reads)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  __syncthreads();
  // ---- normalise + affine (+ Swish) in place, chunk by chunk; each chunk leaves as soon as it is done ----
  if (active) {
    int g = (tv * 8) / cpg, r = tv * 8 - g * cpg;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float2 mr = s_mr[g];
      a[u] *= mr.y;
      b[u] -= mr.x * a[u];
      if (++r == cpg) { r = 0; ++g; }
    }
  }
  for (int c = 0; c < kRChunks; ++c) {
    const int cr = min(rch, rows - c * rch);
    if (cr <= 0) break;
    if (active) {
      __half* tp = tile + (size_t)c * rch * C + tv * 8;
      for (int pp = tr; pp < cr; pp += R) {
        uint4* pv = reinterpret_cast<uint4*>(tp + (size_t)pp * C);
        float f0[8];
        h8_to_f(*pv, f0);
#pragma unroll
        for (int u = 0; u < 8; ++u) f0[u] = f0[u] * a[u] + b[u];
        uint4 o0;
        o0.x = swish_pack<kMode>(f0[0], f0[1]); o0.y = swish_pack<kMode>(f0[2], f0[3]);
        o0.z = swish_pack<kMode>(f0[4], f0[5]); o0.w = swish_pack<kMode>(f0[6], f0[7]);
        *pv = o0;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_store(y + base + (long long)c * rch * C, tile + (size_t)c * rch * C, (uint32_t)cr * (uint32_t)C * 2u, false);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // peers are done reading this CTA's s_group
}

// ---------------------------------------------------------------------------------------------------------
// Slab kernel (small samples: every UNet GroupNorm of the 256x384 workload): ONE CTA per (sample, slab of `sg` whole
// groups), the slab -- hw rows of sg * cpg channels -- parked in the CTA's shared memory by per-thread cp.async. A CTA
// owns complete groups, so nothing is exchanged between CTAs: no cluster, no barrier across CTAs, no workspace; the
// tensor is read once and written once. A thread keeps one 8-channel column and rows tr, tr + R, ... through all three
// phases (load, statistics, normalise).
// sg = the smallest group count whose channels fill whole 16-byte vectors (cpg 10 -> 4 groups = 40 channels, ...).
// Few, large slabs (2 x 320 @ 32x48: 16 slabs of 123 KB) are SPLIT: `split` CTAs each load the whole slab and compute the
// same statistics redundantly (cheap: adds and FMAs out of shared memory, the extra loads are L2 hits), but each
// normalises and stores only its own rows_per rows (the expensive part: Swish, stores) -- still nothing exchanged.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSlabThreads = 512;
constexpr int kSlabDefaultKB = 128;  // largest slab the default dispatch gives to the slab kernel (SDEO_GN_F16_SLAB_KB)
struct SlabGeom {
  int hw, C, cpg, sg, sv, slabs, tile_bytes;  // sv = 16-byte vectors per slab row, slabs per sample
  int split, rows_per;                        // CTAs per slab (each normalises rows_per rows), see slab_plan
  int threads;                                // CTA size: 256 when a slab holds few vectors, else kSlabThreads
};
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

template <int kMode>
__global__ void __launch_bounds__(kSlabThreads, 1)
gn_slab_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
               __half* __restrict__ y, SlabGeom gm, float eps) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(128) unsigned char gs_smem[];
  __shared__ float s_mean[8], s_rstd[8];
  const int tid = threadIdx.x;
  const int sv = gm.sv, hw = gm.hw, C = gm.C, cpg = gm.cpg;
  const int piece = blockIdx.x % gm.split, is = blockIdx.x / gm.split;
  const int img = is / gm.slabs, slab = is - img * gm.slabs;
  const int c0 = slab * gm.sg * cpg;
  const int T = (int)blockDim.x;  // 256 for small slabs, 512 otherwise (slab_plan)
  const int R = T / sv;
  const int tr = tid / sv, tj = tid - tr * sv;
  const bool active = tr < R;
  uint4* tile = reinterpret_cast<uint4*>(gs_smem);                       // [hw][sv]
  float* part = reinterpret_cast<float*>(gs_smem + gm.tile_bytes);       // [T][16]
  float* part2 = part + T * 16;                                          // [P][W], P * W <= T
  const long long col0 = ((long long)img * hw) * C + c0 + tj * 8;
  if (active)
    for (int r = tr; r < hw; r += R) cp_async16(&tile[r * sv + tj], x + col0 + (long long)r * C);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // the affine parameters of this thread's column travel while the tile lands
  float a[8], b[8];
  if (active) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + tj * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + tj * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0 + tj * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + tj * 8) + 1);
    a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w; a[4] = g1.x; a[5] = g1.y; a[6] = g1.z; a[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- statistics: per-thread column sums, then a fixed-order fold (deterministic) ----
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active)
      for (int r = tr; r < hw; r += R) {
        float f[8];
        h8_to_f(tile[r * sv + tj], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
      }
    float4* dst = reinterpret_cast<float4*>(part + (size_t)tid * 16);
    dst[0] = make_float4(s[0], s[1], s[2], s[3]); dst[1] = make_float4(s[4], s[5], s[6], s[7]);
    dst[2] = make_float4(q[0], q[1], q[2], q[3]); dst[3] = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  // stage A: part is [R][W] (W = sv * 16 values per row of threads); P thread groups each sum rows p, p + P, ...
  const int W = sv * 16, P = T / W;
  {
    const int p = tid / W, col = tid - p * W;
    if (p < P) {
      float acc = 0.f;
      for (int r = p; r < R; r += P) acc += part[(size_t)r * W + col];
      part2[p * W + col] = acc;
    }
  }
  __syncthreads();
  // stage B: warp g folds group g: cpg channels x P parts, lanes strided, then a shuffle tree (fixed pattern)
  {
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < gm.sg) {
      float ss = 0.f, qq = 0.f;
      for (int i = lane; i < cpg * P; i += 32) {
        const int p = i / cpg, ch = warp * cpg + (i - p * cpg);  // slab-relative channel
        const int col = (ch >> 3) * 16 + (ch & 7);
        ss += part2[p * W + col];
        qq += part2[p * W + col + 8];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        qq += __shfl_xor_sync(0xffffffffu, qq, o);
      }
      if (lane == 0) {
        const float inv = 1.0f / ((float)hw * (float)cpg);
        const float mean = ss * inv;
        float var = qq * inv - mean * mean;
        var = var < 0.f ? 0.f : var;
        s_mean[warp] = mean;
        s_rstd[warp] = rsqrtf(var + eps);
      }
    }
  }
  __syncthreads();
  // ---- normalise this thread's own vectors out of shared memory, straight to global ----
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (tj * 8 + j) / cpg;
      a[j] *= s_rstd[g];
      b[j] -= s_mean[g] * a[j];
    }
    // (rows of this CTA's piece; with split > 1 they were loaded by other threads: the barriers above order that)
    const int r_hi = min(hw, (piece + 1) * gm.rows_per);
    for (int r = piece * gm.rows_per + tr; r < r_hi; r += R) {
      float f[8];
      h8_to_f(tile[r * sv + tj], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], a[j], b[j]);
      uint4 o;
      o.x = swish_pack<kMode>(f[0], f[1]); o.y = swish_pack<kMode>(f[2], f[3]);
      o.z = swish_pack<kMode>(f[4], f[5]); o.w = swish_pack<kMode>(f[6], f[7]);
      *reinterpret_cast<uint4*>(y + col0 + (long long)r * C) = o;
    }
  }
}

// Slab geometry; non-zero when the shape does not suit the slab kernel (no slab of <= 8 whole groups fills 16-byte
// vectors, or a slab does not fit `max_kb` of shared memory).
static int slab_plan(int n, int hw, int c, int groups, int max_kb, SlabGeom* g, size_t* smem) {
  if (c % 8 != 0 || groups <= 0 || c % groups != 0 || max_kb <= 0) return -1;
  const int cpg = c / groups;
  int sg = 1;
  while (sg <= 8 && ((sg * cpg) % 8 != 0 || groups % sg != 0)) sg *= 2;
  if (sg > 8) return -1;
  const int sv = sg * cpg / 8;
  if (sv * 16 > kSlabThreads) return -1;  // stage A needs one thread per (column, value)
  const size_t tile = ((size_t)hw * sv * 16 + 127) & ~(size_t)127;
  if (tile > ((size_t)max_kb << 10)) return -1;
  const size_t total = tile + (size_t)kSlabThreads * 16 * sizeof(float) + (size_t)kSlabThreads * sizeof(float);
  if (total > (size_t)kGSSmemTotal) return -1;
  // big slabs are split into pieces of <= 16 KB (measured on 2 x 320 @ 32x48, 123 KB slabs: 12.4 / 10.1 / 9.1 / 8.6 us at
  // 1 / 2 / 4 / 8 CTAs per slab); slabs under 48 KB are not (2 x 640 @ 16x24, 30 KB: 6.0 us alone, 7.8 split in two).
  // SDEO_GN_F16_SLAB_SPLIT forces 1 / 2 / 4 / 8.
  int split = 1;
  if (tile >= (size_t)(48 << 10))
    while (split < 8 && tile / (size_t)split > (size_t)(16 << 10)) split *= 2;
  const int force = gs_env_int_fwd("SDEO_GN_F16_SLAB_SPLIT", 0, 0, 8);
  if (force == 1 || force == 2 || force == 4 || force == 8) split = force;
  if ((long long)n * (groups / sg) * split > 0x7fffffffLL) return -1;
  g->hw = hw; g->C = c; g->cpg = cpg; g->sg = sg; g->sv = sv; g->slabs = groups / sg; g->tile_bytes = (int)tile;
  g->split = split;
  g->rows_per = (hw + split - 1) / split;
  // CTA size: 256 threads whenever the fold fits (one thread per (column, value): sv <= 16) -- measured 8.1 / 5.5 / 4.6 / 4.8 us
  // on the four UNet shapes against 8.6 / 6.1 / 5.9 / 6.9 with 512
  int threads = sv * 16 <= 256 ? 256 : kSlabThreads;
  const int tf = gs_env_int_fwd("SDEO_GN_F16_SLAB_THREADS", 0, 0, kSlabThreads);
  if ((tf == 256 || tf == 512) && sv * 16 <= tf) threads = tf;
  g->threads = threads;
  *smem = total;
  return 0;
}

// cluster size + rows per CTA of the resident variant; non-zero when the sample does not fit `max_cs` CTAs
static int gr_plan(int n, int hw, int c, int groups, int sms, int max_cs, RGeom* g, size_t* smem) {
  if (c % 8 != 0 || c / 8 > kGSThreads || groups > 32 || c % groups != 0) return -1;
  const size_t aux = ((size_t)2 * c + kGSThreads * 16 + 512) * sizeof(float);
  if (aux + (size_t)c * 2 > (size_t)kGSSmemTotal) return -1;
  const size_t cap = (kGSSmemTotal - aux) & ~(size_t)127;
  const size_t row = (size_t)c * 2;
  int cs = 1;
  while (cs <= max_cs && (size_t)((hw + cs - 1) / cs) * row > cap) cs *= 2;
  if (cs > max_cs) return -1;
  // more CTAs per sample while SMs are idle and a CTA still has a worthwhile piece
  while (cs * 2 <= max_cs && (long long)n * cs * 2 <= sms && (size_t)((hw + cs - 1) / cs) * row > (size_t)(24 << 10)) cs *= 2;
  const int force = gs_env_int_fwd("SDEO_GN_F16_CLUSTER", 0, 0, 16);
  if (force && (force & (force - 1)) == 0 && force <= max_cs && (size_t)((hw + force - 1) / force) * row <= cap) cs = force;
  g->n = n; g->hw = hw; g->C = c; g->groups = groups; g->cs = cs;
  g->rpc = (hw + cs - 1) / cs;
  g->tile_bytes = (int)(((size_t)g->rpc * row + 127) & ~(size_t)127);
  *smem = (size_t)g->tile_bytes + aux;
  return 0;
}

// SDEO_GN_F16_HINTS: bit 0 = L2 evict_first hints; bits 0x100 (no fold) / 0x400 (no store) / 0x800 (no apply arithmetic) /
// 0x1000 (no statistics arithmetic) are TIMING-ONLY ablations that produce wrong results (tools/sweep_groupnorm.py).
// tuning switches (read per call): SDEO_GN_F16_BUFS = tile buffers per CTA (2..8, default 4), SDEO_GN_F16_TILE_KB = upper
// bound of a tile in KB (default: what the buffers allow)
static int gs_env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}
static int gs_env_int_fwd(const char* name, int dflt, int lo, int hi) { return gs_env_int(name, dflt, lo, hi); }

static int gs_geometry(int n, int hw, int c, int sms, GSGeom* g, size_t* smem) {
  // large tensors: fewer, larger tiles (the per-visit costs weigh more than the extra look-ahead); small ones: 4 buffers
  int kGSBufs = gs_env_int("SDEO_GN_F16_BUFS", (long long)n * hw * c * 2 > (64LL << 20) ? 2 : 4, 2, kGSMaxBufs);
  const size_t tile_kb = (size_t)gs_env_int("SDEO_GN_F16_TILE_KB", 1024, 1, 1024);
  if (c % 8 != 0 || c / 8 > kGSThreads) return -1;  // one 8-channel vector column per arithmetic thread
  // SDEO_GN_F16_GROUPS=2: two thread groups of 256 on alternate units (needs C <= 2048: a group covers a pixel row). Measured
  // on B200 and NOT the default: 16x256@256x256 591 us with two groups vs 542 us with one (4 buffers), 439 us with one group and
  // 2 large buffers - the per-visit cost is arithmetic latency inside a group, not idle time at its barriers.
  const int ngroups = gs_env_int("SDEO_GN_F16_GROUPS", 1, 1, c / 8 <= kGSThreads / 2 ? 2 : 1);
  if (ngroups == 2) kGSBufs = kGSBufs < 4 ? 4 : (kGSBufs & ~1);  // each group needs a tile in hand and one in flight
  const size_t aux = ((size_t)ngroups * 2 * c + kGSThreads * 16 + 512) * sizeof(float);
  if (aux + (size_t)kGSBufs * c * 2 > (size_t)kGSSmemTotal) return -1;
  size_t tile_cap = ((kGSSmemTotal - aux) / kGSBufs) & ~(size_t)127;
  if (tile_cap > tile_kb * 1024 && tile_kb * 1024 >= (size_t)c * 2) tile_cap = tile_kb * 1024;
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
  g->n = n; g->hw = hw; g->C = c; g->bufs = kGSBufs; g->ngroups = ngroups;
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
  if (tiles > (1 << 29) || groups > 32) return -1;  // the (mean, rstd) row is fetched by one warp
  g->groups = groups;
  const int G = (int)(tiles < sms ? tiles : sms);
  // apply lag: a sample's statistics tiles plus two rounds of the grid, so that the statistics an apply visit needs are
  // (almost always) complete when its turn comes
  long long lag = lag_env >= 0 ? (long long)lag_env : (long long)g->chunks + 2LL * G;
  if (lag < g->chunks) lag = g->chunks;  // ordering requirement: apply(t) after statistics of every tile of t's sample
  if (lag > tiles) lag = tiles;
  g->lag = (int)lag;
  // One folder CTA per sample is safe when no apply visit of the folder can precede its fold point: the fold point
  // (< (f + 1) * chunks + G) lies below the sample's first apply unit, or (lag == tiles) every CTA makes all its statistics
  // visits before its first apply visit and the remaining folds fire at the top of that one.
  g->two_level = (lag >= (long long)g->chunks + G || lag == tiles) && !getenv("SDEO_GN_F16_ONE_LEVEL") ? 1 : 0;
  // tree fold: only where every duty fires before the CTA's first apply visit (lag == tiles) - team duties then wait on
  // statistics visits only - and where the folder would otherwise walk more than two batches of slots
  {
    const long long P = g->chunks < (long long)g->ngroups * G ? g->chunks : (long long)g->ngroups * G;
    g->tree = g->two_level && lag == tiles && P > 2 * kGSTeam && !getenv("SDEO_GN_F16_NO_TREE") ? 1 : 0;
  }
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

// Largest usable cluster: 16 (non-portable) when the device can host at least one such cluster of full-size CTAs, else 8.
// Also raises the kernels' dynamic shared memory limit (first call).
static int gr_max_cluster() {
  static int max_cs = 0;
  if (max_cs) return max_cs;
  int best = 8;
  bool ok = true;
  void (*fns[3])(const __half*, const float*, const float*, __half*, RGeom, float) = {gn_resident_kernel<0>, gn_resident_kernel<1>,
                                                                                     gn_resident_kernel<2>};
  for (int i = 0; i < 3; ++i)
    ok = ok && cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal) == cudaSuccess;
  if (ok && !getenv("SDEO_GN_F16_NO_CLUSTER16")) {
    bool np = true;
    for (int i = 0; i < 3; ++i)
      np = np && cudaFuncSetAttribute(fns[i], cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (np) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16, 1, 1);
      cfg.blockDim = dim3(kGSThreads, 1, 1);
      cfg.dynamicSmemBytes = kGSSmemTotal;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, fns[2], &cfg) == cudaSuccess && nclusters >= 1) best = 16;
    }
  }
  (void)cudaGetLastError();
  max_cs = ok ? best : 1;
  return max_cs;
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
  const size_t parts = (size_t)(g.chunks < 2 * sms ? g.chunks : 2 * sms);  // partial slots per sample (<= thread groups x grid size)
  const size_t stream_bytes = (size_t)n * (parts + 1 + (parts + kGSTeam - 1) / kGSTeam) * groups * sizeof(unsigned long long);  // partial + (mean, rstd) + team slots
  return stream_bytes > two_pass ? stream_bytes : two_pass;
}

// Host-side view of the schedule (tests): plan[0..6] = tiles per sample, pixels per tile, apply lag (tiles),
// grid size, dynamic shared memory bytes, tile buffer stride, tile buffers; returns non-zero when the shape falls back to two launches.
extern "C" int sdeo_groupnorm_f16_plan(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t sms, int32_t* plan) {
  if (!plan || n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || c % 8 != 0) return set_error(SDEO_EINVAL, "groupnorm_f16_plan: bad argument");
  GSGeom g;
  size_t smem = 0;
  int G = 0;
  if (gs_plan(n, hw, c, groups, sms > 0 ? sms : 148, -1, &g, &smem, &G)) return 1;
  plan[0] = g.chunks; plan[1] = g.ppc; plan[2] = g.lag; plan[3] = G; plan[4] = (int32_t)smem; plan[5] = g.tile_stride; plan[6] = g.bufs;
  return 0;
}
// visits of CTA `cta` of `grid`, in order: out[2 i] = visit kind (0 statistics, 1 apply), out[2 i + 1] = tile; returns the count
// (at most `cap` pairs are written). The kernel walks the same iterator.
extern "C" int32_t sdeo_groupnorm_f16_visits(int32_t cta, int32_t grid, int32_t tiles, int32_t lag, int32_t* out, int32_t cap) {
  if (cta < 0 || grid <= 0 || cta >= grid || tiles <= 0 || lag <= 0 || lag > tiles) return 0;
  GSIter it{cta, 0};
  int32_t cnt = 0;
  for (bool ok = gs_settle(it, tiles, lag, grid); ok; ok = gs_next(it, tiles, lag, grid), ++cnt) {
    if (out && cnt < cap) {
      out[2 * cnt] = it.v;
      out[2 * cnt + 1] = it.v ? it.u - lag : it.u;
    }
  }
  return cnt;
}

// which kernel a call with this geometry runs on a device with `sms` SMs and clusters of up to `max_cluster` CTAs (<= 0: 148 / 8):
// 3 slab (info[0] = groups per slab, info[1] = 16-byte vectors per slab row, info[2] = shared memory bytes), 2 resident (info[0] =
// cluster size, info[1] = pixel rows per CTA, info[2] = shared memory bytes), 1 two launches (the default beyond a cluster),
// 0 streamed (only under SDEO_GN_F16_VARIANT=stream)
extern "C" int sdeo_groupnorm_f16_variant(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t sms, int32_t max_cluster,
                                          int32_t* info) {
  if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || c % 8 != 0 || c % groups != 0) return set_error(SDEO_EINVAL, "groupnorm_f16_variant: bad argument");
  const char* variant = getenv("SDEO_GN_F16_VARIANT");
  if (!variant || variant[0] == 's') {
    if (!variant || variant[1] == 'l') {  // unset or "slab"
      SlabGeom sg_;
      size_t ssmem = 0;
      if (slab_plan(n, hw, c, groups, gs_env_int("SDEO_GN_F16_SLAB_KB", kSlabDefaultKB, 0, 200), &sg_, &ssmem) == 0) {
        if (info) { info[0] = sg_.sg; info[1] = sg_.sv; info[2] = (int32_t)ssmem; }
        return 3;
      }
    }
  }
  RGeom rg;
  size_t rsmem = 0;
  if (gr_plan(n, hw, c, groups, sms > 0 ? sms : 148, max_cluster > 0 ? max_cluster : 8, &rg, &rsmem) == 0) {
    if (info) { info[0] = rg.cs; info[1] = rg.rpc; info[2] = (int32_t)rsmem; }
    return 2;
  }
  GSGeom g;
  size_t smem = 0;
  int G = 0;
  // samples that do not fit a cluster take the two-launch grid (round 2: its pipelined row loads run both passes at ~5.9 TB/s,
  // ahead of the streamed kernel on every shape measured); SDEO_GN_F16_VARIANT=stream opts into the streamed kernel
  if (!(variant && variant[0] == 's' && variant[1] == 't')) return 1;
  return gs_plan(n, hw, c, groups, sms > 0 ? sms : 148, -1, &g, &smem, &G) ? 1 : 0;
}

// Host-side view of the slab kernel's geometry (tests): plan[0..6] = groups per slab, 16-byte vectors per slab row, slabs per
// sample, CTAs per slab (split), pixel rows each of them normalises, threads per CTA, dynamic shared memory bytes; non-zero
// when the shape does not suit the slab kernel within `max_kb` (<= 0: the default limit) of shared memory per slab.
extern "C" int sdeo_groupnorm_f16_slab_plan(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t max_kb, int32_t* plan) {
  if (!plan || n <= 0 || hw <= 0 || c <= 0 || groups <= 0) return set_error(SDEO_EINVAL, "groupnorm_f16_slab_plan: bad argument");
  SlabGeom g;
  size_t smem = 0;
  if (slab_plan(n, hw, c, groups, max_kb > 0 ? max_kb : kSlabDefaultKB, &g, &smem)) return 1;
  plan[0] = g.sg; plan[1] = g.sv; plan[2] = g.slabs; plan[3] = g.split; plan[4] = g.rows_per; plan[5] = g.threads;
  plan[6] = (int32_t)smem;
  return 0;
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
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 one(1, 1, 1);
  const int mode = with_silu ? swish_mode : 0;
  // "slab" | "resident" | "stream" | unset: slab when a slab of whole groups fits kSlabDefaultKB of shared memory, else
  // resident when the sample fits a cluster, else two launches
  const char* variant = getenv("SDEO_GN_F16_VARIANT");
  const bool want_stream = variant && variant[0] == 's' && variant[1] == 't';
  const bool want_slab = variant && variant[0] == 's' && variant[1] == 'l';
  if (!two_pass && (!variant || want_slab)) {
    SlabGeom sg_;
    size_t ssmem = 0;
    if (slab_plan(n, hw, c, groups, gs_env_int("SDEO_GN_F16_SLAB_KB", kSlabDefaultKB, 0, 200), &sg_, &ssmem) == 0) {
      static bool slab_attr = false;
      if (!slab_attr) {
        cudaError_t e = cudaFuncSetAttribute(gn_slab_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_slab_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_slab_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSSmemTotal);
        if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
        slab_attr = true;
      }
      auto sfn = mode == 0 ? gn_slab_kernel<0> : (mode == 1 ? gn_slab_kernel<1> : gn_slab_kernel<2>);
      return launch_k("groupnorm_f16 (slab)", sfn, dim3((unsigned)(n * sg_.slabs * sg_.split)), dim3((unsigned)sg_.threads), ssmem, st, one,
                      (const __half*)x, gamma, beta, (__half*)y, sg_, eps);
    }
    if (want_slab) return set_error(SDEO_EINVAL, "groupnorm_f16: the shape does not suit the slab kernel (slab variant forced)");
  }
  if (!two_pass && !want_stream && !want_slab) {
    RGeom rg;
    size_t rsmem = 0;
    if (gr_plan(n, hw, c, groups, gs_sm_count(), gr_max_cluster(), &rg, &rsmem) == 0) {
      auto rfn = mode == 0 ? gn_resident_kernel<0> : (mode == 1 ? gn_resident_kernel<1> : gn_resident_kernel<2>);
      return launch_k("groupnorm_f16 (resident)", rfn, dim3((unsigned)(n * rg.cs)), dim3(kGSThreads), rsmem, st,
                      dim3((unsigned)rg.cs, 1, 1), (const __half*)x, gamma, beta, (__half*)y, rg, eps);
    }
    if (variant && variant[0] == 'r') return set_error(SDEO_EINVAL, "groupnorm_f16: the sample does not fit a cluster (resident variant forced)");
  }
  GSGeom g;
  size_t smem = 0;
  int G = 0;
  // default for samples beyond a cluster: the two-launch grid (statistics + apply, norm.cu); the streamed single-launch
  // kernel below (spin-waits on slots of co-resident CTAs) runs only when asked for by SDEO_GN_F16_VARIANT=stream
  if (two_pass || !want_stream || gs_plan(n, hw, c, groups, gs_sm_count(), lag_env, &g, &smem, &G))
    return groupnorm_f16_two_pass(x, gamma, beta, y, n, hw, c, groups, eps, with_silu, workspace, workspace_bytes, stream);
  const size_t nparts_ = (size_t)(g.chunks < g.ngroups * G ? g.chunks : g.ngroups * G);
  const size_t slot_bytes = (size_t)n * (nparts_ + 1 + (nparts_ + kGSTeam - 1) / kGSTeam) * groups * sizeof(unsigned long long);
  if (workspace_bytes < slot_bytes)
    return set_error(SDEO_EINVAL, "groupnorm_f16: workspace too small (sdeo_groupnorm_f16_workspace_bytes)");
  if (cudaMemsetAsync(workspace, 0xFF, slot_bytes, st) != cudaSuccess) {  // every partial slot: "not written"
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
  auto fn = mode == 0 ? gn_stream_kernel<0> : (mode == 1 ? gn_stream_kernel<1> : gn_stream_kernel<2>);
  return launch_k("groupnorm_f16 (streamed)", fn, dim3((unsigned)G), dim3(kGSAll), smem, st, one, (const __half*)x, gamma,
                  beta, (__half*)y, (unsigned long long*)workspace, g, eps, hints);
}
