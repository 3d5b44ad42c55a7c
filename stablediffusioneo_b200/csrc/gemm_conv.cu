// Implicit-GEMM convolution / linear for sm_100a: TMA-fed, tcgen05.mma with TMEM accumulators.
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * W[cout, tap, cin]
//   M = output pixels, tiled as a (bn x bh x bw) box of up to 128 pixels fetched by ONE 4-D TMA box per
//       (tap, 64-channel chunk): the tap shift is a coordinate offset, the zero padding is TMA out-of-bounds
//       fill, the stride-2 case is the tensor map's elementStrides, torch.cat([x1,x2],1) is two tensor maps.
//   N = output channels, tile BN (multiple of 16, <= 256, chosen at run time).
//   K = taps * channels in chunks of 64 bf16 (one 128-byte swizzle atom per row).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
//
// Epilogue, two phases:
//   1. TMEM -> registers -> fp32 tile in shared memory (the drained pipeline stages are reused), thread = row.
//   2. the tile is walked in (row, 8-column) items with consecutive threads on consecutive columns, so that bias /
//      time-embedding / residual loads and the output stores are coalesced 16-32 byte vectors with many in flight:
//      bias + emb -> SiLU -> scale -> + residual -> bf16 / fp32 (+ bf16 twin) | GEGLU | QKV scatter.
// Split-K (small-M, weight-streaming layers): the S K-slices of one output tile run as ONE THREAD-BLOCK CLUSTER
// (1,1,S). After the mainloop every CTA holds its partial tile in its own shared memory; after a cluster barrier CTA r
// reduces rows [r*128/S, (r+1)*128/S) over all S partials through distributed shared memory (ld.shared::cluster), in
// rank order (deterministic), and runs the epilogue for those rows. No global workspace, no atomics.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"
#include <stdlib.h>

namespace sdeo {

constexpr int kConvThreads = 384;  // warp 0 TMA, warp 1 MMA, warps 2..5 TMEM drain; all 12 warps run epilogue phase 2
// TMA producers: lane 0 of warps 0, 9, 10, 11. One thread pays ~200 cycles per mbarrier.try_wait (even on a completed phase)
// plus ~45 + 3 cycles/KB per cp.async.bulk.tensor it issues (tools/exp_issue.cu) -- 450..600 cycles per K step, more than the
// MMAs of the step take (2 N cycles for a 128 x N x 64 step). K step i belongs to producer i % nprod, so the waits and
// issues of consecutive steps overlap. nprod <= ring depth (a parity wait cannot tell barrier phases two apart: with the
// in-order consumer that bound keeps every producer less than two phases ahead of any slot -- tests/test_host_cpu.py
// models it), and the depth is a multiple of nprod, so a ring slot is always refilled by the same thread.
constexpr int kProducers = 4;
constexpr int kHelperThreads = 96;  // warps 6..8: column vectors, folded-LayerNorm row statistics, residual prefetch
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB
constexpr int kMaxCluster = 8;              // portable cluster size limit
constexpr int kGnfWarps = 7;                // folded GroupNorm: warps 2..8 normalise the A tiles
constexpr int kGnfThreads = kGnfWarps * 32;

struct ConvKParams {
  // K loop
  int kw, pad, stride;
  // sub-pixel phase of a nearest-x2-upsample + 3x3 convolution (sdeo_conv_args::up2_phase): a 2x2 filter over the LOW
  // resolution input whose leading padding differs per axis (pad_h / pad_w = 1 - phase bit) and whose output pixel (h, w)
  // lands at (2h + out_off_h, 2w + out_off_w) of the [N, Ho_full, Wo_full] output. Ordinary convs: pad_h = pad_w = pad,
  // out_mul = 1, offsets 0, Ho_full / Wo_full = Ho / Wo.
  int pad_h, pad_w, out_mul, out_off_h, out_off_w, Ho_full, Wo_full;
  int c1_chunks, chunks_per_tap;
  int total_chunks, chunks_per_split, splits;
  // M tiling
  int bn_, bh, bw, rows_valid;
  int tiles_h, tiles_w;
  int N, Ho, Wo;
  // N tiling
  int BN, cout, stages, tmem_cols;
  // epilogue
  int epi_mode, act, y_fp32;
  const float* bias;
  const float* emb;
  const int* emb_step;  // table mode: every sample adds row *emb_step of emb (device-side DDIM step counter)
  const void* residual;
  int ldr, residual_f32;
  float scale;
  void* y;
  int ldy;
  __nv_bfloat16* y2;
  int ldy2;
  // qkv
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  __nv_bfloat16* vt;
  int heads, dhead, tokens, ldv, qkv_first;
  long long* dbg;  // optional per-CTA phase timestamps (SDEO_CONV_DEBUG), 16 slots per CTA
  int res_smem_off;        // > 0: the residual tile is fetched into shared memory at this byte offset by one TMA box of
  int res_tx;              //      res_tx bytes during the mainloop
  // STATS == 2 (producer of a LayerNorm input): per output row (sum, sum of squares) over this N tile's columns
  float2* row_stats;      // [n_tile][row_stats_ld]
  int row_stats_ld;
  // LayerNorm folded into this GEMM (its A operand is the RAW bf16 x, its weight carries gamma): the epilogue turns
  // acc into rstd[row] * (acc - mean[row] * csum[n]) (+ bias', which carries beta @ W). mean / rstd come from the
  // producer's row statistics: ln_parts partials per row.
  const float2* ln_stats;  // [ln_parts][ln_ld]
  int ln_parts, ln_ld;
  float ln_invc, ln_eps;
  const float* ln_csum;    // [rows_packed] column sums of the packed (bf16) weight, packed row order
  float2* gn_stats;  // STATS == 1: per-(M tile, K-slice rank) per-channel (sum, sum of squares) of the final outputs
  // HALO mode (3x3, stride 1, one sample per tile): ONE A tile per 64-channel chunk -- the (bh+2) x (bw+2) halo box of the
  // output tile, pixel pitch hpitch = bw+2 -- serves all nine taps: tap (ky,kx) is the same tile read from row offset
  // ky*hpitch + kx (a 128-byte-aligned descriptor start; the 128B swizzle is a function of the absolute shared-memory
  // address, so TMA's layout and the MMA's reads agree for any row offset -- measured, tools/exp_rowshift.cu). MMA row r
  // is halo position (r / hpitch, r % hpitch): rows with r % hpitch >= bw are the halo columns, computed and discarded.
  // A and B (one weight tile per tap) travel through separate rings. K order: chunk-major, tap-minor.
  int halo, hpitch, a_stages, a_stage_bytes;
  // PAIR mode (cta_group::2): CTAs (2i, 2i+1) of grid.x form a pair inside the (2,1,S) cluster and run ONE M=256 MMA per
  // K step: each stages its own 128 rows of A and its half (BN/2 rows) of the weight tile. m_tiles = real M tiles (grid.x is
  // rounded up to even; the odd one out is a null tile whose loads fall outside the tensor and whose rows are all invalid).
  int pair, m_tiles;
  int nprod;  // TMA producer threads in use (<= kProducers); stages % nprod == 0
  // epilogue phase 2 geometry, computed on the host (integer divisions by run-time values cost ~100 cycles each on the
  // epilogue's critical path): rows per K-slice rank, (row, 8- or 16-column) items per row, rows / leftover items one
  // pass of the 384 threads covers, and the multiplier that turns threadIdx.x / cols_items into a multiply-shift
  int rows_per, cols_items, step_rows, step_cols, ci_magic;
  // GroupNorm (+ SiLU) folded into the A operand path (sdeo_conv_args::gnf_*): warps 2..8 fold the producers' partial
  // statistics into a per-(sample of the tile, channel) table (a, b) in shared memory while the first tiles are in flight,
  // then turn every A tile that lands into silu(a * x + b) IN PLACE (positions outside the input stay zero: the
  // convolution pads the normalised tensor) and hand it to the MMA issuer through a second barrier set (xf_bar / xf_a).
  int gnf, gnf_silu, gnf_off;            // on / SiLU / byte offset of the table in shared memory
  const float2* gnf_st1;
  const float2* gnf_st2;
  int gnf_parts1, gnf_parts2, gnf_c1, gnf_c2, gnf_groups, gnf_cpg;
  const float* gnf_gamma;
  const float* gnf_beta;
  float gnf_eps, gnf_inv;                // 1 / (H * W * channels per group)
  int Hin, Win;
  int l2_prefetch;  // producers pull the CTA's whole weight slice into L2 before the grid dependency resolves (SDEO_L2_PREFETCH=1: on)
  int probe;  // MMA issuer probes the next stage's barrier while it issues the current step (SDEO_NO_PROBE=1: off)
};

#define SDEO_DBG(slot)                                                                               \
  do {                                                                                               \
    if (p.dbg) p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); \
  } while (0)

struct RowInfo {
  bool valid;
  int batch;      // sample index
  long long pix;  // linear output pixel index
};

// ---- cluster / distributed shared memory ----
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into a peer's, completing on the PEER's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void st_smem_f4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void load8_f32(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8_bf16(const float* x) {
  uint4 o;
  o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
  o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
  return o;
}

// Compile-time epilogue variants (the hot instantiations carry no runtime flag tests or scalar tails).
enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_TWIN = 2 };
enum { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2 };

// NORMAL epilogue, fast path: cout % 16 == 0 (every N tile full) and all row pitches vector-aligned (checked on the host).
// cv: this item's 8 entries of the CTA's column vector in shared memory (bias, plus the time-embedding row when the whole
// tile adds the same one); batch >= 0: the embedding row differs between the rows of the tile and comes from global memory.
template <int OUT, int RES>
__device__ __forceinline__ void epi_normal_fast(const ConvKParams& p, long long pix, int batch, int n, float* v,
                                                const uint4& raw0, const uint4& raw1, const float* cv) {
  {
    const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
    v[0] += c0.x; v[1] += c0.y; v[2] += c0.z; v[3] += c0.w; v[4] += c1.x; v[5] += c1.y; v[6] += c1.z; v[7] += c1.w;
  }
  if (batch >= 0) {
    float e[8];
    load8_f32(p.emb + (long long)batch * p.cout + n, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += e[j];
  }
  if (p.act == SDEO_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
  } else if (p.act == SDEO_ACT_QUICK_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = quick_gelu_f(v[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= p.scale;
  if (RES == RES_F32) {
    v[0] += __uint_as_float(raw0.x); v[1] += __uint_as_float(raw0.y); v[2] += __uint_as_float(raw0.z);
    v[3] += __uint_as_float(raw0.w); v[4] += __uint_as_float(raw1.x); v[5] += __uint_as_float(raw1.y);
    v[6] += __uint_as_float(raw1.z); v[7] += __uint_as_float(raw1.w);
  } else if (RES == RES_BF16) {
    float2 a = unpack_bf16x2(raw0.x), b = unpack_bf16x2(raw0.y), c = unpack_bf16x2(raw0.z), d = unpack_bf16x2(raw0.w);
    v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
  }
  if (OUT == OUT_BF16) {
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.ldy + n) = pack8_bf16(v);
  } else {
    float* yp = reinterpret_cast<float*>(p.y) + pix * p.ldy + n;
    if ((p.ldy & 7) == 0) {
      // one 256-bit store per lane: whole 32-byte sectors (two 16-byte stores per lane leave every sector half-written
      // per instruction and halved the measured store rate)
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                   "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                   : "memory");
    } else {
      *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (OUT == OUT_F32_TWIN) *reinterpret_cast<uint4*>(p.y2 + pix * p.ldy2 + n) = pack8_bf16(v);
  }
}

// NORMAL epilogue of one (row, 8 columns) item: v holds the fp32 accumulators of columns n .. n+7.
__device__ __noinline__ void epi_normal_item(const ConvKParams& p, const RowInfo& ri, int n, float* v, bool res_pref,
                                                const uint4& raw0, const uint4& raw1) {
  if (n >= p.cout) return;
  const bool full = (n + 8 <= p.cout);
  if (full) {
    if (p.bias) {
      float b[8];
      load8_f32(p.bias + n, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += b[j];
    }
    if (p.emb) {
      float e[8];
      load8_f32(p.emb + (long long)ri.batch * p.cout + n, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += e[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (n + j < p.cout) {
        if (p.bias) v[j] += __ldg(p.bias + n + j);
        if (p.emb) v[j] += __ldg(p.emb + (long long)ri.batch * p.cout + n + j);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = v[j];
    if (p.act == SDEO_ACT_SILU) t = silu_f(t);
    else if (p.act == SDEO_ACT_QUICK_GELU) t = quick_gelu_f(t);
    v[j] = t * p.scale;
  }
  if (res_pref) {  // residual vector was prefetched by the caller (8 columns, aligned)
    if (p.residual_f32) {
      v[0] += __uint_as_float(raw0.x); v[1] += __uint_as_float(raw0.y); v[2] += __uint_as_float(raw0.z);
      v[3] += __uint_as_float(raw0.w); v[4] += __uint_as_float(raw1.x); v[5] += __uint_as_float(raw1.y);
      v[6] += __uint_as_float(raw1.z); v[7] += __uint_as_float(raw1.w);
    } else {
      float2 a = unpack_bf16x2(raw0.x), b = unpack_bf16x2(raw0.y), c = unpack_bf16x2(raw0.z), d = unpack_bf16x2(raw0.w);
      v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
    }
  } else if (p.residual) {
    if (p.residual_f32) {
      const float* rp = reinterpret_cast<const float*>(p.residual) + ri.pix * p.ldr + n;
      if (full && ((p.ldr & 3) == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(rp), b = *reinterpret_cast<const float4*>(rp + 4);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n + j < p.cout) v[j] += rp[j];
      }
    } else {
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + ri.pix * p.ldr + n;
      if (full && ((p.ldr & 7) == 0)) {
        float r[8];
        load8_bf16(rp, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n + j < p.cout) v[j] += __bfloat162float(rp[j]);
      }
    }
  }
  if (p.y_fp32) {
    float* yp = reinterpret_cast<float*>(p.y) + ri.pix * p.ldy + n;
    if (full && ((p.ldy & 3) == 0)) {
      *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n + j < p.cout) yp[j] = v[j];
    }
    if (p.y2) {  // bf16 twin of an fp32 residual-stream tensor, for consumers that read it through TMA
      __nv_bfloat16* y2p = p.y2 + ri.pix * p.ldy2 + n;
      if (full && ((p.ldy2 & 7) == 0)) {
        *reinterpret_cast<uint4*>(y2p) = pack8_bf16(v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n + j < p.cout) y2p[j] = __float2bfloat16(v[j]);
      }
    }
  } else {
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + ri.pix * p.ldy + n;
    if (full && ((p.ldy & 7) == 0)) {
      *reinterpret_cast<uint4*>(yp) = pack8_bf16(v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n + j < p.cout) yp[j] = __float2bfloat16(v[j]);
    }
  }
}

// QKV epilogue of one item: scatter into head-major q/k and transposed v.
__device__ __forceinline__ void epi_qkv_item(const ConvKParams& p, const RowInfo& ri, int n, float* v, const float* cv) {
  if (n >= p.cout) return;
  const int C = p.heads * p.dhead;
  const int which = n / C + p.qkv_first;
  const int nc = n % C;
  const int head = nc / p.dhead;
  const int dd = nc % p.dhead;
  const int b = (int)(ri.pix / p.tokens);
  const int tok = (int)(ri.pix % p.tokens);
  {
    const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
    v[0] += c0.x; v[1] += c0.y; v[2] += c0.z; v[3] += c0.w; v[4] += c1.x; v[5] += c1.y; v[6] += c1.z; v[7] += c1.w;
  }
  const long long bh = (long long)b * p.heads + head;
  if (which < 2) {
    __nv_bfloat16* dst = (which == 0 ? p.q : p.k) + (bh * p.tokens + tok) * p.dhead + dd;
    *reinterpret_cast<uint4*>(dst) = pack8_bf16(v);
  }
  // which == 2 (v): written transposed by qkv_store_vt() with a token-major item mapping
}

// V^T part of the QKV epilogue: item = (8 consecutive token rows, one v column) -> ONE 16-byte store into
// vt[b*heads + head][dd][tok .. tok+7]. Lanes run along the columns so the shared-memory reads are conflict-free.
__device__ __forceinline__ void qkv_store_vt(const ConvKParams& p, const float* tile, int LD, const int* row_pix,
                                             const float2* ln_vec, int n_base, int tid, int nthreads, const float* colv,
                                             const float* csumv) {
  const int C = p.heads * p.dhead;
  int vb = (2 - p.qkv_first) * C, ve = (3 - p.qkv_first) * C;  // global column range holding v
  if (vb < n_base) vb = n_base;
  if (ve > n_base + p.BN) ve = n_base + p.BN;
  if (ve > p.cout) ve = p.cout;
  const int ncols = ve - vb;
  if (ncols <= 0 || !p.vt) return;
  const int groups = (p.rows_valid + 7) / 8;
  for (int it = tid; it < groups * ncols; it += nthreads) {
    const int c = it % ncols, g = it / ncols;
    const int n = vb + c;
    const int nc = n - (2 - p.qkv_first) * C;
    const int head = nc / p.dhead, dd = nc % p.dhead;
    const float bias = colv[n - n_base];
    const float csum = ln_vec ? csumv[n - n_base] : 0.f;
    float x[8];
    int pix[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = g * 8 + i;
      pix[i] = r < p.rows_valid ? row_pix[r] : -1;
      float a = tile[(size_t)r * LD + (n - n_base)];
      if (ln_vec) {  // folded LayerNorm (see ConvKParams::ln_stats)
        const float2 mr = ln_vec[r];
        a = mr.y * (a - mr.x * csum);
      }
      x[i] = a + bias;
    }
    const int tok0 = pix[0] >= 0 ? pix[0] % p.tokens : 0;
    const bool vec = pix[0] >= 0 && pix[7] == pix[0] + 7 && tok0 + 7 < p.tokens && (tok0 & 7) == 0;
    if (vec) {
      const long long bh = (long long)(pix[0] / p.tokens) * p.heads + head;
      *reinterpret_cast<uint4*>(p.vt + (bh * p.dhead + dd) * p.ldv + tok0) = pack8_bf16(x);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (pix[i] < 0) continue;
        const long long bh = (long long)(pix[i] / p.tokens) * p.heads + head;
        p.vt[(bh * p.dhead + dd) * p.ldv + pix[i] % p.tokens] = __float2bfloat16(x[i]);
      }
    }
  }
}

// GEGLU epilogue of one item: y[:, n_out..+7] = (x + bx) * gelu(gate + bg); nb_x / nb_g index the packed bias.
__device__ __forceinline__ void epi_geglu_item(const ConvKParams& p, const RowInfo& ri, int n_out, const float* cv_x,
                                               const float* cv_g, const float* vx, const float* vg) {
  float bx[8], bg[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { bx[j] = cv_x[j]; bg[j] = cv_g[j]; }
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = (vx[j] + bx[j]) * gelu_erf_f(vg[j] + bg[j]);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + ri.pix * p.ldy + n_out;
  *reinterpret_cast<uint4*>(yp) = pack8_bf16(x);
}

// TMA loads signalling an mbarrier given by its 32-bit shared-memory address: this CTA's own barrier, or (PAIR mode, via
// the cta_group::2 form) the pair leader's barrier as a shared::cluster address.
__device__ __forceinline__ void tma2d_to(void* dst, const CUtensorMap* m, uint32_t bar, bool pair, int c0, int c1) {
  if (pair) {
    tma_load_2d_pair(dst, m, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ void tma4d_to(void* dst, const CUtensorMap* m, uint32_t bar, bool pair, int c0, int c1, int c2,
                                         int c3) {
  if (pair) {
    tma_load_4d_pair(dst, m, bar, c0, c1, c2, c3);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
__device__ __forceinline__ void commit_to(uint64_t* bar, bool pair, uint16_t pair_mask) {
  if (pair) tc_commit_pair(bar, pair_mask);
  else tc_commit(bar);
}
__device__ __forceinline__ void mma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, bool pair) {
  if (pair) tc_mma_bf16_pair(d, a, b, idesc, acc);
  else tc_mma_bf16(d, a, b, idesc, acc);
}

// Output pixel (n, h, w) of tile row `row`, or false for a padding row / halo column / pixel outside the tensor.
__device__ __forceinline__ bool tile_row_coords(const ConvKParams& p, int row, int n0, int h0, int w0, int* nn, int* hh,
                                                int* ww) {
  int nl, hl, wl;
  if (p.halo) {
    nl = 0; hl = row / p.hpitch; wl = row % p.hpitch;
    if (wl >= p.bw || hl >= p.bh) return false;
  } else {
    const int per_img = p.bh * p.bw;
    nl = row / per_img;
    const int rem = row % per_img;
    hl = rem / p.bw; wl = rem % p.bw;
  }
  *nn = n0 + nl; *hh = h0 + hl; *ww = w0 + wl;
  return (row < p.rows_valid) && (*nn < p.N) && (*hh < p.Ho) && (*ww < p.Wo);
}

// ---------------------------------------------------------------------------------------------------------------
// Folded GroupNorm (sdeo_conv_args::gnf_*), run by warps 2..8 during the mainloop: fold the producers' partial statistics
// into the (a, b) table, then normalise every A tile that lands, in place. A separate (not inlined) function: inlined, its
// register pressure and code size changed the allocation of the whole kernel body (every conv launch, folded GroupNorm
// or not, ran 3-20% slower).
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void gnf_transform_warps(const ConvKParams& p, uint8_t* smem, uint8_t* tiles, float2* ln_vec,
                                                 uint64_t* full_bar, uint64_t* full_a, uint64_t* xf_bar, uint64_t* xf_a,
                                                 int k_begin, int k_end, int nchunks, int stage_bytes, int n0, int h0, int w0) {
  const int lane = threadIdx.x & 31;
  {
    // ===================== folded GroupNorm: statistics -> (a, b) table, then every A tile in place =====================
    const int t = (int)threadIdx.x - 64;
    const int Cp = p.chunks_per_tap * 64;                 // one entry per K position of a tap (padded channels)
    // table row of one sample: entry of channel c at float2 index (c / 8) * 10 + c % 8 -- 8-channel vectors 80 bytes apart,
    // so that the eight vectors a quarter-warp reads lie in different banks
    const int Cpp = (Cp / 8) * 10;
    float2* ab = reinterpret_cast<float2*>(smem + p.gnf_off);   // [bn_][Cpp]: first (sum, sum of squares), then (a, b)
    float2* gst = ab + (size_t)p.bn_ * Cpp;                     // [bn_][groups] (mean, rstd)
    const int C = p.gnf_c1 + p.gnf_c2;
    const int n_entries = p.bn_ * Cp;
    // tile row -> input position table (shares the folded-LayerNorm row vector's space: the two never meet)
    int* row_hw = reinterpret_cast<int*>(ln_vec);           // tap-by-tap: (sample << 16 | h << 8 | w) inside the tile, -1 = none
    uint8_t* halo_ok = reinterpret_cast<uint8_t*>(ln_vec);  // HALO: halo position lies inside the input
    const int halo_rows = (p.bh + 2) * p.hpitch;
    if (p.halo) {
      for (int r = t; r < halo_rows; r += kGnfThreads) {
        const int hy = r / p.hpitch, hx = r - hy * p.hpitch;
        const int hh = h0 - 1 + hy, ww = w0 - 1 + hx;
        halo_ok[r] = (hh >= 0 && hh < p.Hin && ww >= 0 && ww < p.Win && n0 < p.N) ? 1 : 0;
      }
    } else {
      const int per_img = p.bh * p.bw;
      for (int r = t; r < kBM; r += kGnfThreads) {
        const int nl = r / per_img, rem = r - nl * per_img;
        const int hl = rem / p.bw, wl = rem - hl * p.bw;
        row_hw[r] = (r < p.rows_valid && n0 + nl < p.N) ? ((nl << 16) | (hl << 8) | wl) : -1;
      }
    }
    // 1. per-channel (sum, sum of squares): the producers' partial slots, added in slot order (deterministic). The fold sits
    // in front of the first MMA and every batch of loads is an L2 round trip: four table entries x eight slots in flight
    for (int idx0 = t; idx0 < n_entries; idx0 += 4 * kGnfThreads) {
      const float2* src[4];
      int parts[4], ld[4], slot[4];
      float sum[4], sq[4];
      int max_parts = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = idx0 + e * kGnfThreads;
        src[e] = nullptr; parts[e] = 0; ld[e] = 0; slot[e] = -1; sum[e] = 0.f; sq[e] = 0.f;
        if (idx < n_entries) {
          const int nl = idx / Cp, c = idx - nl * Cp;
          const int n = n0 + nl;
          slot[e] = nl * Cpp + (c >> 3) * 10 + (c & 7);
          if (c < C && n < p.N) {
            const bool first = c < p.gnf_c1;
            parts[e] = first ? p.gnf_parts1 : p.gnf_parts2;
            ld[e] = first ? p.gnf_c1 : p.gnf_c2;
            src[e] = first ? p.gnf_st1 + (size_t)n * parts[e] * ld[e] + c : p.gnf_st2 + (size_t)n * parts[e] * ld[e] + (c - p.gnf_c1);
            max_parts = max(max_parts, parts[e]);
          }
        }
      }
      for (int k = 0; k < max_parts; k += 8) {
        float2 v[4][8];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int u = 0; u < 8; ++u)
            v[e][u] = (k + u < parts[e]) ? __ldcg(src[e] + (size_t)(k + u) * ld[e]) : make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int u = 0; u < 8; ++u) { sum[e] += v[e][u].x; sq[e] += v[e][u].y; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (slot[e] >= 0) ab[slot[e]] = make_float2(sum[e], sq[e]);
    }
    bar_sync(1, kGnfThreads);
    // 2. (mean, rstd) of every (sample of the tile, group)
    for (int idx = t; idx < p.bn_ * p.gnf_groups; idx += kGnfThreads) {
      const int nl = idx / p.gnf_groups, g = idx - nl * p.gnf_groups;
      float sum = 0.f, sq = 0.f;
      for (int j = 0, c = g * p.gnf_cpg; j < p.gnf_cpg; ++j, ++c) {
        const float2 v = ab[nl * Cpp + (c >> 3) * 10 + (c & 7)];
        sum += v.x; sq += v.y;
      }
      const float mean = sum * p.gnf_inv;
      float var = sq * p.gnf_inv - mean * mean;
      var = var < 0.f ? 0.f : var;
      gst[idx] = make_float2(mean, rsqrtf(var + p.gnf_eps));
    }
    bar_sync(1, kGnfThreads);
    // 3. y = a * x + b per (sample, channel); channels beyond the tensor (K padding) stay finite: the weights there are zero
    for (int idx = t; idx < n_entries; idx += kGnfThreads) {
      const int nl = idx / Cp, c = idx - nl * Cp;
      float2 v = make_float2(0.f, 0.f);
      if (c < C) {
        const float2 mr = gst[nl * p.gnf_groups + c / p.gnf_cpg];
        v.x = __ldg(p.gnf_gamma + c) * mr.y;
        v.y = __ldg(p.gnf_beta + c) - mr.x * v.x;
      }
      ab[nl * Cpp + (c >> 3) * 10 + (c & 7)] = v;
    }
    bar_sync(1, kGnfThreads);
    if (t == 0) SDEO_DBG(2);
    const bool with_silu = p.gnf_silu != 0;
    long long dbg_x = 0;
    // One 16-byte unit = 8 channels of one tile row. Thread t owns the LOGICAL 8-channel vector lu = t % 8 of rows
    // t / 8, t / 8 + 28, ...: inside a 128B-swizzled tile that vector sits at physical position lu ^ (row % 8), so a warp
    // still touches four whole 128-byte rows per access (no bank conflicts), and with one sample per tile the thread's
    // sixteen (a, b) values are loaded ONCE per tile into registers instead of once per unit (the table reads were four
    // times the tile's own bytes and bank-conflicted: measured 3400 cycles per 16 KB tile before, ...). Rows are handled kU
    // at a time: all loads of a batch are issued before the first value is used.
    constexpr int kU = 4;
    constexpr int kRowStep = kGnfThreads / 8;   // 28
    const uint32_t lu = (uint32_t)t & 7u;
    const int r0 = t >> 3;
    const bool one_sample = p.bn_ == 1;
    auto xform_tile = [&](uint32_t base, int rows, const float2* abk, auto&& row_of) {
      // row_of(row) -> table row offset (in float2) of the row's sample, or -1: the row stays as it is (zero padding)
      float4 ab0, ab1, ab2, ab3;
      if (one_sample) {
        const float4* abp = reinterpret_cast<const float4*>(abk + lu * 10);
        ab0 = abp[0]; ab1 = abp[1]; ab2 = abp[2]; ab3 = abp[3];
      }
      for (int rb = r0; rb < rows; rb += kU * kRowStep) {
        uint4 raw[kU];
        int off[kU];
        uint32_t addr[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int r = rb + u * kRowStep;
          off[u] = (r < rows) ? row_of(r) : -1;
          addr[u] = base + (uint32_t)r * 128u + ((lu ^ ((uint32_t)r & 7u)) << 4);
          if (off[u] >= 0) asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[u].x), "=r"(raw[u].y), "=r"(raw[u].z), "=r"(raw[u].w) : "r"(addr[u]) : "memory");
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          if (off[u] < 0) continue;
          if (!one_sample) {
            const float4* abp = reinterpret_cast<const float4*>(abk + off[u] + lu * 10);
            ab0 = abp[0]; ab1 = abp[1]; ab2 = abp[2]; ab3 = abp[3];
          }
          const float2 x0 = unpack_bf16x2(raw[u].x), x1 = unpack_bf16x2(raw[u].y), x2 = unpack_bf16x2(raw[u].z), x3 = unpack_bf16x2(raw[u].w);
          float y[8];
          y[0] = fmaf(x0.x, ab0.x, ab0.y); y[1] = fmaf(x0.y, ab0.z, ab0.w);
          y[2] = fmaf(x1.x, ab1.x, ab1.y); y[3] = fmaf(x1.y, ab1.z, ab1.w);
          y[4] = fmaf(x2.x, ab2.x, ab2.y); y[5] = fmaf(x2.y, ab2.z, ab2.w);
          y[6] = fmaf(x3.x, ab3.x, ab3.y); y[7] = fmaf(x3.y, ab3.z, ab3.w);
          if (with_silu) {
            // silu(y) = h + h * tanh(h), h = y / 2: ONE special-function op per element (tanh.approx, relative error 2^-11 --
            // below the bf16 rounding of the result) where y / (1 + exp(-y)) takes two
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float hlf = 0.5f * y[j];
              float th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hlf));
              y[j] = fmaf(hlf, th, hlf);
            }
          }
          st_smem_f4(addr[u], pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
      }
    };
    if (p.halo) {
      const int c_first = k_begin / 9, c_last = (k_end - 1) / 9;
      for (int c = c_first; c <= c_last; ++c) {
        const int na = c - c_first, sa = na % p.a_stages;
        // ONE thread polls the barrier and releases the other six warps through a named barrier: mbarrier.try_wait costs
        // ~200 cycles per warp instruction even on a completed phase and the barrier unit serialises them (seven polling
        // warps: measured ~1500 cycles per tile before the first byte was touched)
        if (t == 0) mbar_wait(&full_a[sa], (uint32_t)((na / p.a_stages) & 1));
        bar_sync(1, kGnfThreads);
        const long long t_x0 = (p.dbg && t == 0) ? clock64() : 0;
        xform_tile(smem_u32(tiles) + (uint32_t)(sa * p.a_stage_bytes), halo_rows, ab + c * 80,
                   [&](int row) -> int { return halo_ok[row] ? 0 : -1; });
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the MMA's operand reads (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_a[sa]);
        if (p.dbg && t == 0) dbg_x += clock64() - t_x0;
      }
    } else {
      int s = 0;
      uint32_t ph = 0;
      int tap = k_begin / p.chunks_per_tap, within = k_begin % p.chunks_per_tap;
      int ky = tap / p.kw, kx = tap % p.kw;
      for (int i = 0; i < nchunks; ++i) {
        if (t == 0) mbar_wait(&full_bar[s], ph);
        bar_sync(1, kGnfThreads);
        const long long t_x0 = (p.dbg && t == 0) ? clock64() : 0;
        const int hb = h0 * p.stride + ky - p.pad_h, wb = w0 * p.stride + kx - p.pad_w;
        xform_tile(smem_u32(tiles) + (uint32_t)(s * stage_bytes), p.rows_valid, ab + within * 80, [&](int row) -> int {
          const int info = row_hw[row];
          const int hh = hb + ((info >> 8) & 255) * p.stride, ww = wb + (info & 255) * p.stride;
          return (info >= 0 && hh >= 0 && hh < p.Hin && ww >= 0 && ww < p.Win) ? (info >> 16) * Cpp : -1;
        });
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[s]);
        if (p.dbg && t == 0) dbg_x += clock64() - t_x0;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++within == p.chunks_per_tap) {
          within = 0;
          if (++kx == p.kw) { kx = 0; ++ky; }
        }
      }
    }
    if (p.dbg && t == 0) p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 15] = dbg_x;
    }
}

// MODE: SDEO_EPI_*; OUT / RES: see enums above; FAST: vector-aligned NORMAL epilogue (else the generic item path).
// STATS (NORMAL + FAST + fp32 output only): the epilogue also reduces the FINAL output values of this CTA's rows to
// per-channel (sum, sum of squares) and writes them to p.gn_stats[(M tile * S + K-slice rank)][cout] -- the GroupNorm
// that consumes this tensor folds those partials instead of re-reading the tensor for its statistics.
// (STATS == 2: per-row statistics for a LayerNorm consumer instead, see ConvKParams::row_stats.)
// LNF: a LayerNorm is folded into this GEMM (ConvKParams::ln_stats); kept out of the other instantiations' code.
// PAIR: CTA pairs (cta_group::2). A compile-time variant, not a run-time flag: a kernel that contains cta_group::2
// instructions can only be launched with an even cluster width (measured: "cluster misconfiguration" otherwise).
template <int MODE, int OUT, int RES, bool FAST, int STATS, bool LNF, bool PAIR>
__device__ __forceinline__ void conv_gemm_body(const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                                               const CUtensorMap& tmR, const ConvKParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  // layout: [barriers: 1 KB][pipeline stages, reused as the fp32 epilogue tile]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* tmem_full_bar = empty_bar + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint64_t* recv_bar = full_bar + 48;  // split-K: completes when all S partial slices of this CTA's rows have arrived
  uint64_t* res_bar = full_bar + 49;   // completes when the residual tile (one TMA box) has landed
  uint64_t* full_a = full_bar + 40;    // HALO mode: the A (halo tile) ring, up to 4 stages
  uint64_t* empty_a = full_bar + 44;
  uint64_t* xf_bar = full_bar + 50;    // folded GroupNorm: stage s holds the NORMALISED A tile (tap-by-tap ring, <= 8 stages)
  uint64_t* xf_a = full_bar + 58;      // folded GroupNorm, HALO mode: halo tile slot sa is normalised (<= 4 slots)
  int* row_pix = reinterpret_cast<int*>(smem + 512);  // [128] output pixel index per tile row
  float2* ln_vec = reinterpret_cast<float2*>(smem + 1024);  // [128] (mean, rstd) of the folded LayerNorm per tile row
  float* colv = reinterpret_cast<float*>(smem + 2048);   // [BN] bias (+ the tile's time-embedding row) per column of this N tile
  float* csumv = reinterpret_cast<float*>(smem + 3072);  // [BN] folded-LayerNorm column sums
  uint8_t* tiles = smem + 4096;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int split = blockIdx.z;  // K slice; cluster = (1,1,S), or (2,1,S) in PAIR mode: cluster rank = px + 2 * split
  constexpr bool pair = PAIR;
  const uint32_t px = pair ? (cluster_ctarank() & 1u) : 0u;  // position inside the CTA pair; 0 = leader (issues the MMAs)
  const bool leader = px == 0;
  const uint32_t leader_rank = pair ? (uint32_t)(2 * split) : 0u;
  const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
  const int bn_cta = pair ? p.BN / 2 : p.BN;            // weight rows this CTA stages per K step
  const int b_bytes = bn_cta * 128;
  const int stage_bytes = kATileBytes + b_bytes;
  const int tx_mult = pair ? 2 : 1;                      // the leader's barrier counts both CTAs' bytes
  const int iw = m_tile % p.tiles_w;
  const int ih = (m_tile / p.tiles_w) % p.tiles_h;
  const int in_ = m_tile / (p.tiles_w * p.tiles_h);
  const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in_ * p.bn_;

  const int k_begin = split * p.chunks_per_split;
  int k_end = k_begin + p.chunks_per_split;
  if (k_end > p.total_chunks) k_end = p.total_chunks;
  const int nchunks = k_end - k_begin;

  if (threadIdx.x == 0) SDEO_DBG(0);
  const int trc = trace_start(1ULL | ((unsigned long long)MODE << 40) | ((unsigned long long)p.splits << 44) |
                              ((unsigned long long)p.BN << 48) | ((unsigned long long)(p.halo & 1) << 60));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if (p.chunks_per_tap > p.c1_chunks) tma_prefetch_desc(&tmA2);
    if (p.res_smem_off) tma_prefetch_desc(&tmR);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&empty_a[s], 1);
    }
    if (p.gnf) {  // one arrival per transforming warp (warps 2..8)
      for (int s = 0; s < (p.halo ? p.a_stages : p.stages); ++s) mbar_init(p.halo ? &xf_a[s] : &xf_bar[s], kGnfWarps);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(recv_bar, 1);
    mbar_init(res_bar, 1);
    fence_mbar_init();
    if (p.splits > 1) {
      // bytes this CTA will receive: one slice of its own rows from each of the S K-slice ranks (armed here, long
      // before any peer can push: pushes start after a cluster barrier this CTA has not arrived at yet)
      const int rp = (p.rows_valid + p.splits - 1) / p.splits;
      int mine = min(p.rows_valid, (split + 1) * rp) - split * rp;
      if (mine < 0) mine = 0;
      mbar_expect_tx(recv_bar, (uint32_t)(p.splits * mine * (p.BN + 4) * 4));
    }
  }
  if (warp == 1) {
    if (pair) {
      tmem_alloc_pair(tmem_ptr_smem, (uint32_t)p.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) {  // the partner's barriers must be initialised before a TMA completion or a multicast commit can reach them
    cluster_arrive();
    cluster_wait();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) SDEO_DBG(1);
  // PDL: the next kernel of the stream may start its prologue / weight prefetch from here on. (Not earlier: a
  // successor CTA co-resident on this SM must not win the TMEM allocation while this CTA still needs its own -- it
  // would hold the columns until this grid completes, i.e. forever.)
  griddep_launch_dependents();
  // PDL: activations, residual, emb and the output buffers belong to earlier kernels of the stream until the grid
  // dependency resolves. Only the TMA thread goes on without waiting: it first streams WEIGHT tiles (constants).
  // producer index of this thread (-1: not a producer)
  // (single-thread loops sit inside `if (elect_one())` under warp-uniform branches: ptxas then issues their TMA / MMA
  //  instructions straight from uniform registers; under a `lane == 0` test it wraps each one in an election loop with
  //  vector -> uniform register moves, several times the cost -- tools/exp_issue.cu)
  const int prod = warp == 0 ? 0 : (warp >= 12 - (kProducers - 1) ? warp - (12 - kProducers) : -1);   // warp-uniform
  const int np = p.nprod;
  const bool prod_warp = prod >= 0 && prod < np;
  if (prod_warp) {
    if (!elect_one()) griddep_wait();   // (the elected lane waits inside its producer loop, after the weight prefetch)
  } else {
    griddep_wait();
  }

  const int LD = p.BN + 4;  // fp32 tile row pitch in floats: 16-byte aligned rows, conflict-free 16 B row writes
  float* tile = reinterpret_cast<float*>(tiles);
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    // output pixel of every tile row (-1: padding row / outside the image), read by epilogue phase 2
    const int row = (int)threadIdx.x - 128;
    int nn, hh, ww;
    const bool ok = tile_row_coords(p, row, n0, h0, w0, &nn, &hh, &ww);
    row_pix[row] = ok ? (nn * p.Ho_full + hh * p.out_mul + p.out_off_h) * p.Wo_full + ww * p.out_mul + p.out_off_w : -1;
  }

  // HALO mode ring geometry: [A ring: a_stages x a_stage_bytes][B ring: stages x BN*128]
  uint8_t* b_ring = tiles + (size_t)p.a_stages * p.a_stage_bytes;
  // barrier the TMA completions go to: this CTA's own, or the pair leader's
  auto full_addr = [&](uint64_t* bar) -> uint32_t {
    return pair ? mapa_u32(smem_u32(bar), leader_rank) : smem_u32(bar);
  };
  if (prod_warp && p.halo) {
    // ===================== TMA producers, HALO mode =====================
    // K step i (global index g = k_begin + i): chunk c = g / 9, tap t = g % 9. One halo A tile per chunk, one weight tile
    // per step; the weight tile of (c, t) sits at packed K offset (t * chunks_per_tap + c) * 64. The producer that owns
    // the first step of a chunk inside this CTA's K range also loads the chunk's halo tile.
    if (elect_one()) {
      const int nb0 = n_tile * p.BN + (int)px * bn_cta;
      const uint32_t a_tx = (uint32_t)((p.bh + 2) * p.hpitch) * 128u * (uint32_t)tx_mult;
      const uint32_t b_tx = (uint32_t)(b_bytes * tx_mult);
      const int npre = nchunks < p.stages ? nchunks : p.stages;
      const int c_first = k_begin / 9;
      {
        int cc = (k_begin + prod) / 9, tt = (k_begin + prod) % 9;
        for (int i = prod; i < npre; i += np) {  // weight tiles of the first ring pass: before the grid dependency resolves
          if (leader) mbar_expect_tx(&full_bar[i], b_tx);
          tma2d_to(b_ring + (size_t)i * b_bytes, &tmB, full_addr(&full_bar[i]), pair, (tt * p.chunks_per_tap + cc) * kBK, nb0);
          tt += np;
          while (tt >= 9) { tt -= 9; ++cc; }
        }
        // ... and the rest of this CTA's weight slice into L2: inside a denoising step every layer's weights come from
        // HBM (2.4 GB per step), and a 4..12-stage ring does not cover DRAM latency; the wait for the previous kernel
        // (several microseconds under programmatic dependent launch) does
        if (p.l2_prefetch) {
          for (int i = npre + prod; i < nchunks; i += np) {
            const int g = k_begin + i;
            tma_prefetch_l2_2d(&tmB, ((g % 9) * p.chunks_per_tap + g / 9) * kBK, nb0);
          }
        }
      }
      griddep_wait();
      if (prod == 0) trace_mark(trc, 2);
      if (prod == np - 1 && p.res_smem_off) {
        // the residual tile of the epilogue: ONE unswizzled box (BN channels x the tile's pixel box, with the halo
        // tiling's pixel pitch), dense [tile row][BN] in shared memory; always this CTA's own barrier
        mbar_expect_tx(res_bar, (uint32_t)p.res_tx);
        tma_load_4d(smem + p.res_smem_off, &tmR, res_bar, n_tile * p.BN, w0, h0, n0);
      }
      int c = (k_begin + prod) / 9, t = (k_begin + prod) % 9;
      int sb = prod % p.stages;
      uint32_t phb = (uint32_t)((prod / p.stages) & 1);
      for (int i = prod; i < nchunks; i += np) {
        if (i == 0 || t == 0) {  // first step of a chunk inside this CTA's K range: its halo tile
          const int na = c - c_first;           // halo tiles before this one
          const int sa = na % p.a_stages;
          if (na >= p.a_stages) mbar_wait(&empty_a[sa], (uint32_t)(((na / p.a_stages) & 1) ^ 1));
          if (leader) mbar_expect_tx(&full_a[sa], a_tx);
          uint8_t* a_dst = tiles + (size_t)sa * p.a_stage_bytes;
          if (c < p.c1_chunks)
            tma4d_to(a_dst, &tmA1, full_addr(&full_a[sa]), pair, c * kBK, w0 - 1, h0 - 1, n0);
          else
            tma4d_to(a_dst, &tmA2, full_addr(&full_a[sa]), pair, (c - p.c1_chunks) * kBK, w0 - 1, h0 - 1, n0);
        }
        if (i >= npre) {
          mbar_wait(&empty_bar[sb], phb ^ 1u);
          if (leader) mbar_expect_tx(&full_bar[sb], b_tx);
          tma2d_to(b_ring + (size_t)sb * b_bytes, &tmB, full_addr(&full_bar[sb]), pair, (t * p.chunks_per_tap + c) * kBK, nb0);
        }
        sb += np;
        while (sb >= p.stages) { sb -= p.stages; phb ^= 1u; }
        t += np;
        while (t >= 9) { t -= 9; ++c; }
      }
    }
  } else if (warp == 1 && p.halo) {
    // ===================== MMA issuer, HALO mode (PAIR mode: the leader CTA only) =====================
    // ONE thread runs the whole loop: a warp-wide loop with an elected issuer inside pays ~150 cycles per K step for the
    // election, the reconvergence and the vector -> uniform register moves (tools/exp_issue.cu: 343 vs 192 cycles per
    // 128x64x64 step), more than the MMAs of a narrow N tile take.
    if (leader && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(pair ? 2 * kBM : kBM, (uint32_t)p.BN);
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(tiles));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(b_ring));
    const uint64_t a_step = (uint64_t)(p.a_stage_bytes >> 4), b_step = (uint64_t)(b_bytes >> 4);
    int sb = 0, sa = 0;
    uint32_t phb = 0, pha = 0;
    uint64_t a_desc = a_desc0, b_desc = b_desc0;
    int t = k_begin % 9;
    int ky = t / 3, kx = t % 3;
    long long dbg_wait = 0;
    uint32_t ready = 0;   // the weight tile of step i was probed (landed) while step i - 1 was being issued
    for (int i = 0; i < nchunks; ++i) {
      const long long t_w0 = p.dbg ? clock64() : 0;
      if (i == 0 || t == 0) {
        mbar_wait(p.gnf ? &xf_a[sa] : &full_a[sa], pha);
      }
      if (!ready) mbar_wait(&full_bar[sb], phb);
      if (p.dbg) dbg_wait += clock64() - t_w0;
      tc_fence_after();
      const bool last_of_chunk = (t == 8) || (i == nchunks - 1);
      int sbn = sb + 1;
      uint32_t phbn = phb;
      if (sbn == p.stages) { sbn = 0; phbn ^= 1u; }
      {
        // tap shift = row offset inside the halo tile: (ky * pitch + kx) rows of 128 bytes = 8 descriptor units each
        const uint64_t a_tap = a_desc + (uint64_t)((ky * p.hpitch + kx) * 8);
        if (p.probe) {
          ready = tc_mma_step_probe<PAIR>(tmem_base, a_tap, b_desc, idesc, i > 0 ? 1u : 0u, &empty_bar[sb], pair_mask,
                                          &full_bar[sbn], phbn);
          if (i + 1 == nchunks) ready = 0;
        } else {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_any(tmem_base, a_tap + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u, pair);
          commit_to(&empty_bar[sb], pair, pair_mask);
        }
        if (last_of_chunk) commit_to(&empty_a[sa], pair, pair_mask);
      }
      b_desc += b_step;
      sb = sbn; phb = phbn;
      if (sb == 0) b_desc = b_desc0;
      if (last_of_chunk) {
        a_desc += a_step;
        if (++sa == p.a_stages) { sa = 0; pha ^= 1u; a_desc = a_desc0; }
      }
      if (++kx == 3) { kx = 0; ++ky; }
      if (++t == 9) { t = 0; ky = 0; }
    }
    commit_to(tmem_full_bar, pair, pair_mask);
    SDEO_DBG(3);
    if (p.dbg) p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 9] = dbg_wait;
    }
    __syncwarp();
  } else if (prod_warp) {
    // ===================== TMA producers =====================
    // (all ring / tap / chunk indices advance incrementally: no integer division on the per-chunk path)
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)(p.rows_valid * 128 + b_bytes) * (uint32_t)tx_mult;
      const int nb0 = n_tile * p.BN + (int)px * bn_cta;
      // the first ring pass needs no empty-slot wait; its weight tiles are requested before the grid dependency
      // resolves, so the weight stream of this layer overlaps the tail of the previous kernel
      const int npre = nchunks < p.stages ? nchunks : p.stages;
      for (int i = prod; i < npre; i += np) {
        if (leader) mbar_expect_tx(&full_bar[i], tx_bytes);
        tma2d_to(tiles + (size_t)i * stage_bytes + kATileBytes, &tmB, full_addr(&full_bar[i]), pair, (k_begin + i) * kBK, nb0);
      }
      // ... and the rest of this CTA's weight slice into L2 (see the HALO producer)
      if (p.l2_prefetch) {
        for (int i = npre + prod; i < nchunks; i += np) tma_prefetch_l2_2d(&tmB, (k_begin + i) * kBK, nb0);
      }
      griddep_wait();
      if (prod == 0) trace_mark(trc, 2);
      if (prod == np - 1 && p.res_smem_off) {
        // the residual tile of the epilogue: ONE unswizzled box (BN channels x the tile's pixel box, with the halo
        // tiling's pixel pitch), dense [tile row][BN] in shared memory; always this CTA's own barrier
        mbar_expect_tx(res_bar, (uint32_t)p.res_tx);
        tma_load_4d(smem + p.res_smem_off, &tmR, res_bar, n_tile * p.BN, w0, h0, n0);
      }
      int s = prod % p.stages;
      uint32_t ph = (uint32_t)((prod / p.stages) & 1);
      const int g0 = k_begin + prod;
      int tap = g0 / p.chunks_per_tap, within = g0 % p.chunks_per_tap;
      int ky = tap / p.kw, kx = tap % p.kw;
      long long dbg_wait = 0;
      for (int i = prod; i < nchunks; i += np) {
        uint8_t* a_dst = tiles + (size_t)s * stage_bytes;
        if (i >= npre) {
          const long long t_w0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (p.dbg) dbg_wait += clock64() - t_w0;
          if (leader) mbar_expect_tx(&full_bar[s], tx_bytes);
        }
        const int wc = w0 * p.stride + kx - p.pad_w;
        const int hc = h0 * p.stride + ky - p.pad_h;
        const uint32_t fb = full_addr(&full_bar[s]);
        if (within < p.c1_chunks)
          tma4d_to(a_dst, &tmA1, fb, pair, within * kBK, wc, hc, n0);
        else
          tma4d_to(a_dst, &tmA2, fb, pair, (within - p.c1_chunks) * kBK, wc, hc, n0);
        if (i >= npre) tma2d_to(a_dst + kATileBytes, &tmB, fb, pair, (k_begin + i) * kBK, nb0);
        within += np;
        while (within >= p.chunks_per_tap) {
          within -= p.chunks_per_tap;
          if (++kx == p.kw) { kx = 0; ++ky; }
        }
        s += np;
        while (s >= p.stages) { s -= p.stages; ph ^= 1u; }
      }
      if (p.dbg && prod == 0) {
        long long* d = p.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
        d[10] = dbg_wait;
        d[11] = clock64();   // producer 0 done (all its loads issued)
        d[12] = p.stages;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread feeds the tensor pipe: everything between two chunks' MMAs is on the critical path (the pipe idles
    // while this thread probes the barrier), so the loop carries precomputed descriptors and no divisions.
    if (leader && elect_one()) {  // (PAIR mode: the leader CTA issues the M=256 MMAs for both)
      const uint32_t idesc = umma_idesc_bf16(pair ? 2 * kBM : kBM, (uint32_t)p.BN);
      const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(tiles));
      const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(tiles) + kATileBytes);
      const uint64_t desc_step = (uint64_t)(stage_bytes >> 4);  // start-address field advances by one stage
      int s = 0;
      uint32_t ph = 0;
      uint64_t a_desc = a_desc0, b_desc = b_desc0;
      long long dbg_wait = 0;
      uint32_t ready = 0;   // the full barrier of step i was probed (complete) while step i - 1 was being issued
      uint64_t* const data_bar = p.gnf ? xf_bar : full_bar;   // folded GroupNorm: wait for the normalised tile
      for (int i = 0; i < nchunks; ++i) {
        if (!ready) {
          const long long t_w0 = p.dbg ? clock64() : 0;
          mbar_wait(&data_bar[s], ph);
          if (p.dbg) dbg_wait += clock64() - t_w0;
        }
        tc_fence_after();
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == p.stages) { sn = 0; phn ^= 1u; }
        // four MMAs on the 32-byte K slices of the chunk (+2 per slice in the descriptor's (addr >> 4) field), the commit
        // that frees this stage (in both CTAs of a pair) once they have read it, and the probe of the next stage
        if (p.probe) {
          ready = tc_mma_step_probe<PAIR>(tmem_base, a_desc, b_desc, idesc, i > 0 ? 1u : 0u, &empty_bar[s], pair_mask,
                                          &data_bar[sn], phn);
          if (i + 1 == nchunks) ready = 0;
        } else {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_any(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u, pair);
          commit_to(&empty_bar[s], pair, pair_mask);
        }
        a_desc += desc_step;
        b_desc += desc_step;
        if (sn == 0) { a_desc = a_desc0; b_desc = b_desc0; }
        s = sn; ph = phn;
      }
      commit_to(tmem_full_bar, pair, pair_mask);  // accumulator complete (all MMAs done => every stage consumed)
      SDEO_DBG(3);
      if (p.dbg) p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 9] = dbg_wait;
    }
    __syncwarp();
  }

  // the time-embedding row is the same for every row of the tile (one sample per tile, or table mode): it joins the bias
  // in the column vector; otherwise the epilogue reads it per row from global memory
  const bool emb_in_colv = p.emb && (p.emb_step || p.bn_ == 1);
  if (warp >= 6 && warp < 9) {
    // the epilogue's per-column vectors (bias + time-embedding row, folded-LayerNorm column sums): global memory ->
    // shared memory during the mainloop (read from global memory inside the epilogue they cost an L2 round trip per item
    // batch -- measured: the largest share of epilogue phase 2)
    const int nb = n_tile * p.BN;
    // (the null tile of an odd PAIR grid lies beyond the last sample: clamp its row -- it read 4 * cout bytes past the end of
    //  emb, an illegal address whenever emb happened to end a mapped segment; its results are discarded either way)
    const long long erow = emb_in_colv ? (p.emb_step ? __ldg(p.emb_step) : (n0 < p.N ? n0 : p.N - 1)) : 0;
    for (int c = (int)threadIdx.x - 192; c < p.BN; c += kHelperThreads) {
      const int n = nb + c;
      float v = 0.f, cs = 0.f;
      if (n < p.cout) {
        if (p.bias) v = __ldg(p.bias + n);
        if (emb_in_colv) v += __ldg(p.emb + erow * p.cout + n);
        if (LNF) cs = __ldg(p.ln_csum + n);
      }
      colv[c] = v;
      if (LNF) csumv[c] = cs;
    }
  }

  if (LNF && warp >= 6 && warp < 9) {
    // ===================== folded LayerNorm: (mean, rstd) of every tile row (warps 6..11, during the mainloop) =========
    const int per_img = p.bh * p.bw;
    for (int row = (int)threadIdx.x - 192; row < kBM; row += kHelperThreads) {
      const int nl = row / per_img, rem = row % per_img;
      const int hl = rem / p.bw, wl = rem % p.bw;
      const int nn = n0 + nl, hh = h0 + hl, ww = w0 + wl;
      float2 mr = make_float2(0.f, 0.f);
      if (row < p.rows_valid && nn < p.N && hh < p.Ho && ww < p.Wo) {
        const long long pix = ((long long)nn * p.Ho + hh) * p.Wo + ww;
        float sm_ = 0.f, sq_ = 0.f;
        for (int k0 = 0; k0 < p.ln_parts; k0 += 8) {  // up to 8 independent loads in flight, added in part order
          float2 a[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            a[k] = (k0 + k < p.ln_parts) ? __ldcg(p.ln_stats + (size_t)(k0 + k) * p.ln_ld + pix) : make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 8; ++k) { sm_ += a[k].x; sq_ += a[k].y; }
        }
        const float mean = sm_ * p.ln_invc;
        float var = sq_ * p.ln_invc - mean * mean;
        var = var < 0.f ? 0.f : var;
        mr = make_float2(mean, rsqrtf(var + p.ln_eps));
      }
      ln_vec[row] = mr;
    }
  }

  if (p.gnf && warp >= 2 && warp < 2 + kGnfWarps)
    gnf_transform_warps(p, smem, tiles, ln_vec, full_bar, full_a, xf_bar, xf_a, k_begin, k_end, nchunks, stage_bytes, n0, h0, w0);

  // ===================== epilogue phase 1: TMEM -> fp32 tile in shared memory, all 12 warps =====================
  // A warp may read the TMEM lane quarter (warp % 4); the three warps of a quarter split the columns in 32-wide
  // chunks; rows land in this CTA's own tile (the drained pipeline stages).
  // Split-K (S > 1, one cluster per output tile): CTA r reduces rows [r*rows_per, (r+1)*rows_per). After its tile is
  // complete every CTA sends, for each owner r, the contiguous rows_per x LD block of its tile with ONE asynchronous
  // bulk copy through distributed shared memory (cp.async.bulk.shared::cluster) into slot [own rank] of the owner's
  // receive area [S][rows_per][LD]; the copy completes on the owner's mbarrier, which doubles as the "all partials are
  // here" signal. The first cluster barrier (arrive right after the accumulator is complete, wait after phase 1) only
  // guards the receive areas, which alias the peers' pipeline stages; the second (arrive after the slices are in, wait
  // at teardown) keeps every CTA's shared memory alive until its outgoing copies have been read.
  __syncwarp();
  const int S = p.splits;
  const int rows_per = p.rows_per;
  float* recv = tile + (size_t)p.rows_valid * LD;  // S > 1 only: right behind the valid rows of the tile
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  if (threadIdx.x == 64) SDEO_DBG(4);
  if (S > 1) cluster_arrive();
  {
    const int quarter = warp & 3, third = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t dst_row = smem_u32(tile) + (uint32_t)(row * LD * 4);
    const bool live = (S == 1) || (row < p.rows_valid);  // split-K: the receive area starts right behind the valid rows
    for (int c = third * 32; c < p.BN; c += 96) {
      if (c + 32 <= p.BN) {
        uint32_t r[32];
        tmem_ld32(taddr_row + (uint32_t)c, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int g = 0; g < 8; ++g) st_smem_f4(dst_row + (uint32_t)((c + 4 * g) * 4), r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        }
      } else {
        uint32_t r[16];
        tmem_ld16(taddr_row + (uint32_t)c, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int g = 0; g < 4; ++g) st_smem_f4(dst_row + (uint32_t)((c + 4 * g) * 4), r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        }
      }
    }
    tc_fence_before();
  }
  if (threadIdx.x == 64) SDEO_DBG(5);
  if (S > 1) {
    fence_proxy_async_smem();  // the tile was written by ordinary stores; the bulk copies read it through the async proxy
    __syncthreads();
    cluster_wait();            // every peer's MMAs are done: the receive areas are free
    if (threadIdx.x < (unsigned)S) {
      const int r = (int)threadIdx.x;
      int rows_r = min(p.rows_valid, (r + 1) * rows_per) - r * rows_per;
      if (rows_r > 0) {
        const uint32_t src = smem_u32(tile) + (uint32_t)(r * rows_per * LD * 4);
        const uint32_t peer = pair ? (uint32_t)(2 * r) + px : (uint32_t)r;  // cluster rank of K slice r of this M tile
        const uint32_t dst = dsmem_addr(smem_u32(recv) + (uint32_t)(split * rows_per * LD * 4), peer);
        dsmem_bulk_copy(dst, src, (uint32_t)(rows_r * LD * 4), dsmem_addr(smem_u32(recv_bar), peer));
      }
    }
    mbar_wait(recv_bar, 0);
    cluster_arrive();          // (waited for at teardown)
  } else {
    __syncthreads();
  }

  if (threadIdx.x == 64) SDEO_DBG(6);
  {
    // ===================== epilogue phase 2: coalesced (row, 8-column) items, all 12 warps =====================
    // Items are processed U at a time per thread: every load of a batch (tile values from local / distributed
    // shared memory, residual vectors from global memory) is issued before the first value is used. The item
    // index advances without divisions (row / column stepping by the constant thread count).
    constexpr int U = 2;  // measured: U = 4 is slower (register pressure, longer dependent chains in split mode)
    constexpr bool geglu = (MODE == SDEO_EPI_GEGLU);
    const int r_begin = split * rows_per;
    const int r_end = min(p.rows_valid, r_begin + rows_per);
    const int cols_items = p.cols_items;  // items per row (BN / 8; GEGLU: BN / 16)
    const int n_base = n_tile * p.BN;
    const uint32_t slice_bytes = (uint32_t)(rows_per * LD) * 4u;  // S > 1: one received slice per K-slice rank
    const int hw_out = p.Ho_full * p.Wo_full;
    const int half = p.BN / 2;
    const bool emb_global = p.emb && !emb_in_colv;   // (several samples per tile without a step table)
    if (MODE == SDEO_EPI_NORMAL && FAST && RES != RES_NONE && p.res_smem_off) mbar_wait(res_bar, 0);
    if (threadIdx.x == 64) SDEO_DBG(13);
    int dbg_iter = 0;
    // STATS: every thread keeps ONE column item (ci) and strides over rows, so that it can accumulate column sums in
    // registers; the (at most cols_items - 1) threads beyond the last full row group stay idle.
    const int step_rows = p.step_rows, step_cols = STATS ? 0 : p.step_cols;
    const int row_of_tid = (int)(((uint32_t)threadIdx.x * (uint32_t)p.ci_magic) >> 16);   // threadIdx.x / cols_items
    int row = r_begin + row_of_tid;
    int ci = (int)threadIdx.x - row_of_tid * cols_items;
    if (STATS && (int)threadIdx.x >= step_rows * cols_items) row = r_end;
    float st_s[STATS == 1 ? 8 : 1], st_q[STATS == 1 ? 8 : 1];
    if (STATS == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { st_s[j] = 0.f; st_q[j] = 0.f; }
    }
    // STATS == 2 reduces every row over the cols_items lanes that hold it with warp shuffles: all 384 threads then run
    // the same number of iterations (items beyond the last row are dummies)
    int iters_left = (max(r_end - r_begin, 0) + step_rows * U - 1) / (step_rows * U);
    while (STATS == 2 ? (iters_left-- > 0) : (row < r_end)) {
      if (threadIdx.x == 64 && dbg_iter++ == 1) SDEO_DBG(14);
      float v[U][8], g[U][8];
      uint4 raw0[U], raw1[U];
      int pixs[U], cols[U], rws[U];
      uint32_t offs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pixs[u] = -1;
        rws[u] = row < r_end ? row : r_begin;
        cols[u] = ci * 8;
        // S == 1: the tile is indexed by the tile row; S > 1: slices hold this CTA's rows only (row - r_begin)
        offs[u] = (uint32_t)(((size_t)(row < r_end ? row - (S > 1 ? r_begin : 0) : 0) * LD + cols[u]) * sizeof(float));
        if (row < r_end) {
          pixs[u] = row_pix[row];
          if (MODE == SDEO_EPI_NORMAL && FAST && RES != RES_NONE && pixs[u] >= 0 && p.res_smem_off) {
            constexpr int kElem = (RES == RES_F32) ? 4 : 2;
            const uint4* rp = reinterpret_cast<const uint4*>(smem + p.res_smem_off + (size_t)(row * p.BN + cols[u]) * kElem);
            raw0[u] = rp[0];
            if (RES == RES_F32) raw1[u] = rp[1];
          } else if (MODE == SDEO_EPI_NORMAL && FAST && RES != RES_NONE && pixs[u] >= 0) {
            if (RES == RES_F32) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.residual) +
                                                               (long long)pixs[u] * p.ldr + n_base + cols[u]);
              raw0[u] = rp[0];
              raw1[u] = rp[1];
            } else {
              raw0[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) +
                                                        (long long)pixs[u] * p.ldr + n_base + cols[u]);
            }
          }
        }
        // advance to this thread's next item
        ci += step_cols;
        row += step_rows;
        if (ci >= cols_items) { ci -= cols_items; ++row; }
      }
      if (threadIdx.x == 64 && dbg_iter == 1) SDEO_DBG(9);
      if (S == 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float* src = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(tile) + offs[u]);
          const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
          v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
          if (geglu) {
            const float4 c2 = *reinterpret_cast<const float4*>(src + half), d2 = *reinterpret_cast<const float4*>(src + half + 4);
            g[u][0] = c2.x; g[u][1] = c2.y; g[u][2] = c2.z; g[u][3] = c2.w; g[u][4] = d2.x; g[u][5] = d2.y; g[u][6] = d2.z; g[u][7] = d2.w;
          }
        }
      } else {
        // the S partial vectors of an item sit in this CTA's own shared memory (pushed by the peers): summed in rank
        // order, so the result does not depend on timing
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[u][j] = 0.f; g[u][j] = 0.f; }
          const uint8_t* src0 = reinterpret_cast<const uint8_t*>(recv) + offs[u];
          for (int s = 0; s < S; ++s) {
            const float* src = reinterpret_cast<const float*>(src0 + (size_t)s * slice_bytes);
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            v[u][0] += a.x; v[u][1] += a.y; v[u][2] += a.z; v[u][3] += a.w;
            v[u][4] += b.x; v[u][5] += b.y; v[u][6] += b.z; v[u][7] += b.w;
            if (geglu) {
              const float4 c2 = *reinterpret_cast<const float4*>(src + half), d2 = *reinterpret_cast<const float4*>(src + half + 4);
              g[u][0] += c2.x; g[u][1] += c2.y; g[u][2] += c2.z; g[u][3] += c2.w;
              g[u][4] += d2.x; g[u][5] += d2.y; g[u][6] += d2.z; g[u][7] += d2.w;
            }
          }
        }
      }
      if (threadIdx.x == 64 && dbg_iter == 1) SDEO_DBG(10);
      if (LNF) {  // folded LayerNorm: acc -> rstd * (acc - mean * csum)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (pixs[u] < 0) continue;
          const float2 mr = ln_vec[rws[u]];
          const float* cs = csumv + cols[u];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = mr.y * (v[u][j] - mr.x * cs[j]);
          if (geglu) {
            const float* cg = csumv + half + cols[u];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[u][j] = mr.y * (g[u][j] - mr.x * cg[j]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (STATS == 2) {
          // every lane takes part in the shuffles; lanes without a valid item contribute zeros
          float rs = 0.f, rq = 0.f;
          if (pixs[u] >= 0) {
            epi_normal_fast<OUT, RES>(p, (long long)pixs[u], emb_global ? pixs[u] / hw_out : -1, n_base + cols[u], v[u], raw0[u], raw1[u], colv + cols[u]);
#pragma unroll
            for (int j = 0; j < 8; ++j) { rs += v[u][j]; rq += v[u][j] * v[u][j]; }
          }
          for (int off = cols_items >> 1; off > 0; off >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, off);
            rq += __shfl_xor_sync(0xffffffffu, rq, off);
          }
          if (pixs[u] >= 0 && ci == 0) p.row_stats[(size_t)n_tile * p.row_stats_ld + pixs[u]] = make_float2(rs, rq);
          continue;
        }
        if (pixs[u] < 0) continue;
        const int col = cols[u];
        if (MODE == SDEO_EPI_GEGLU) {
          RowInfo ri; ri.valid = true; ri.pix = pixs[u]; ri.batch = 0;
          epi_geglu_item(p, ri, n_tile * half + col, colv + col, colv + half + col, v[u], g[u]);
        } else if (MODE == SDEO_EPI_QKV) {
          RowInfo ri; ri.valid = true; ri.pix = pixs[u]; ri.batch = 0;
          epi_qkv_item(p, ri, n_base + col, v[u], colv + col);
        } else if (FAST) {
          epi_normal_fast<OUT, RES>(p, (long long)pixs[u], emb_global ? pixs[u] / hw_out : -1, n_base + col, v[u], raw0[u], raw1[u], colv + col);
          if (STATS == 1) {  // v[u] now holds the values that were stored
#pragma unroll
            for (int j = 0; j < 8; ++j) { st_s[j] += v[u][j]; st_q[j] += v[u][j] * v[u][j]; }
          }
        } else {
          RowInfo ri; ri.valid = true; ri.pix = pixs[u]; ri.batch = p.emb ? (p.emb_step ? __ldg(p.emb_step) : pixs[u] / hw_out) : 0;
          epi_normal_item(p, ri, n_base + col, v[u], false, raw0[u], raw1[u]);
        }
        if (threadIdx.x == 64 && dbg_iter == 1 && u == 0) SDEO_DBG(11);
      }
      if (threadIdx.x == 64 && dbg_iter == 1) SDEO_DBG(12);
    }
    if (MODE == SDEO_EPI_QKV) qkv_store_vt(p, tile, LD, row_pix, LNF ? ln_vec : nullptr, n_base, (int)threadIdx.x, kConvThreads, colv, csumv);  // S == 1
    if (STATS == 1) {
      // column sums of this CTA's rows: registers -> shared memory (one 16-float record per thread) -> one thread per
      // channel adds the row groups in fixed order (deterministic) -> global partial [M tile * S + rank][channel]
      if (S > 1) cluster_wait();  // the scratch below reuses the tile, which peers read until this barrier completes
      __syncthreads();
      float* scratch = reinterpret_cast<float*>(tiles);
      if ((int)threadIdx.x < step_rows * cols_items) {
        float* dst = scratch + (size_t)threadIdx.x * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) { dst[j] = st_s[j]; dst[8 + j] = st_q[j]; }
      }
      __syncthreads();
      // 2*BN values (sum and sum of squares per channel); kparts thread groups split the row groups of a value, a
      // second pass adds the kparts partial sums: both in fixed order
      const int nval = 2 * p.BN;
      const int kparts = kConvThreads / nval > 0 ? kConvThreads / nval : 1;
      float* part2 = scratch + (size_t)kConvThreads * 16;  // [kparts][nval]
      const int t = (int)threadIdx.x;
      for (int idx = t; idx < nval * kparts; idx += kConvThreads) {  // (BN = 256: 512 values for 384 threads)
        const int val = idx % nval, part = idx / nval;
        const int c = val % p.BN, sq = val / p.BN;
        const float* src = scratch + (size_t)(c >> 3) * 16 + sq * 8 + (c & 7);
        float acc = 0.f;
        for (int rg = part; rg < step_rows; rg += kparts) acc += src[(size_t)rg * cols_items * 16];
        part2[part * nval + val] = acc;
      }
      __syncthreads();
      for (int val = t; val < nval; val += kConvThreads) {
        const int c = val % p.BN, sq = val / p.BN;
        if (n_base + c < p.cout) {
          float acc = 0.f;
          for (int k = 0; k < kparts; ++k) acc += part2[k * nval + val];
          // slot = sample * (parts per sample) + part. One sample per tile: part = (spatial tile, K-slice rank). A tile
          // that holds several whole samples (tiny feature maps) is only handled as a split-K cluster whose per-rank
          // row ranges do not straddle samples (checked on the host): part = position of the range inside the sample.
          int sample = n0, part = (ih * p.tiles_w + iw) * S + split, pps = p.tiles_h * p.tiles_w * S;
          if (p.bn_ > 1) {
            const int per_img = p.bh * p.bw;
            sample = n0 + r_begin / per_img;
            part = (r_begin % per_img) / rows_per;
            pps = per_img / rows_per;
          }
          // (one sample per tile: a rank without rows still owns a slot and writes zeros into it)
          // (a null tile of PAIR mode lies beyond the last sample and owns no slot)
          if (sample < p.N && (p.bn_ == 1 || r_begin < p.rows_valid))
            reinterpret_cast<float*>(p.gn_stats)[(((size_t)sample * pps + part) * p.cout + n_base + c) * 2 + sq] = acc;
        }
      }
    }
  }

  if (threadIdx.x == 64) SDEO_DBG(7);
  // ---- teardown (split-K: peers may still be reading this CTA's tile until the second cluster barrier completes) ----
  tc_fence_before();
  if (p.splits > 1 && STATS != 1) cluster_wait();
  if (pair && p.splits == 1) {  // both CTAs of the pair are done with their accumulators before the pair-wide dealloc
    cluster_arrive();
    cluster_wait();
  }
  __syncthreads();
  trace_mark(trc, 3);
  if (threadIdx.x == 64) SDEO_DBG(8);
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// The two entry points of the body above: one CTA per SM with the full register budget (125 registers; measured 2.7%
// faster on the batch-1 step than the capped build), and the occ2 form capped at 80 registers so that two CTAs of a
// half-shared-memory plan fit one SM (one CTA's epilogue overlaps the other's mainloop on multi-wave grids).
template <int MODE, int OUT, int RES, bool FAST, int STATS, bool LNF, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ ConvKParams p) {
  conv_gemm_body<MODE, OUT, RES, FAST, STATS, LNF, PAIR>(tmA1, tmA2, tmB, tmR, p);
}
template <int MODE, int OUT, int RES, bool FAST, int STATS, bool LNF, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 2)
conv_gemm_kernel_occ2(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ ConvKParams p) {
  conv_gemm_body<MODE, OUT, RES, FAST, STATS, LNF, PAIR>(tmA1, tmA2, tmB, tmR, p);
}

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout,
                                        int c1, int c2, int ksize, int rows_packed, int geglu_bn) {
  const int c1c = (c1 + 63) / 64, c2c = (c2 + 63) / 64;
  const int cpt = c1c + c2c;
  const int taps = ksize * ksize;
  const long long kp = (long long)taps * cpt * 64;
  const long long total = (long long)rows_packed * kp;
  const int cin = c1 + c2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / kp);
    const long long kk = idx % kp;
    const int tap = (int)(kk / (cpt * 64));
    const int within = (int)((kk / 64) % cpt);
    const int j = (int)(kk % 64);
    int src_row = r;
    if (geglu_bn > 0) {
      const int half = geglu_bn / 2, inner = cout / 2;
      const int t = r / geglu_bn, rr = r % geglu_bn;
      src_row = (rr < half) ? (t * half + rr) : (inner + t * half + (rr - half));
    }
    int ci = -1;
    if (within < c1c) {
      const int c = within * 64 + j;
      if (c < c1) ci = c;
    } else {
      const int c = (within - c1c) * 64 + j;
      if (c < c2) ci = c1 + c;
    }
    float v = 0.f;
    if (src_row < cout && ci >= 0) v = w[((long long)src_row * cin + ci) * taps + tap];
    out[idx] = __float2bfloat16(v);
  }
}

__global__ void pack_geglu_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int n2, int geglu_bn) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n2) return;
  const int half = geglu_bn / 2, inner = n2 / 2;
  const int t = r / geglu_bn, rr = r % geglu_bn;
  const int src = (rr < half) ? (t * half + rr) : (inner + t * half + (rr - half));
  out[r] = b[src];
}

// ---------------------------------------------------------------------------------------------
// host side: planner + launcher
// ---------------------------------------------------------------------------------------------
struct ConvPlan {
  int Ho, Wo;
  int bn_, bh, bw, tiles_n, tiles_h, tiles_w;
  int rows_packed, BN, n_tiles;
  int c1c, c2c, cpt, total_chunks, splits, cps;
  int stages, tmem_cols;
  int res_smem_off;
  size_t smem_bytes;
  int halo, hpitch, a_stages, a_stage_bytes, rows_valid;
  int pair;  // CTA pairs (cta_group::2): grid.x = M tiles rounded up to even, cluster (2,1,S)
  int nprod;
  int occ2;  // the plan leaves room for two CTAs per SM (<= 112 KB of shared memory): one CTA's epilogue overlaps the other's mainloop
  int gnf_off, gnf_bytes;  // folded GroupNorm: the (a, b) table + group statistics behind the pipeline / residual tile
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static int pick_bn_impl(int rows_packed, int epi_mode) {
  if (epi_mode == SDEO_EPI_GEGLU) {
    // need BN % 32 == 0 and (rows/2) % (BN/2) == 0. The tile is fixed by the weight packing (SDEO_GEGLU_BN caps it:
    // tuning aid, read when the weights are packed).
    static const int cand[] = {256, 192, 160, 128, 96, 64, 32};
    int cap = 128;   // 128 (64 output columns): the 67 KB staging tile leaves room for two CTAs per SM, and the GELU-heavy
                     // epilogue of one overlaps the mainloop of the other (measured +22% at M = 73728, +4..12% at M <= 3072)
    if (const char* e = getenv("SDEO_GEGLU_BN")) cap = atoi(e);
    for (int bn : cand)
      if (bn <= cap && rows_packed % bn == 0 && (rows_packed / 2) % (bn / 2) == 0) return bn;
    return 0;
  }
  if (rows_packed <= 256) return rows_packed;
  static const int cand[] = {256, 192, 160, 128, 96, 80, 64, 48, 32, 16};
  for (int bn : cand)
    if (rows_packed % bn == 0) return bn;
  return 16;
}

// CTA budget of one launch (sdeo_conv_set_cta_budget): 0 = the whole GPU. The step engine halves it while two
// independent branches (UNet encoder || ControlNet body) run on two streams: at one 200 KB CTA per SM two 100-CTA
// kernels cannot co-run, two <= 74-CTA kernels can, and these layers are latency-bound, not throughput-bound.
static int g_cta_budget = 0;
static inline int cta_limit() { return g_cta_budget > 0 && g_cta_budget < 148 ? g_cta_budget : 148; }

// HALO mode is available for 3x3 stride-1 "same" convolutions whose M tiles hold one sample each (feature maps of at
// least ~64 pixels; smaller ones are weight-streaming layers where the A operand does not matter).
// force_halo: -1 = heuristic (SDEO_HALO=0/1 overrides), 0 = off, 1 = on (fails if not available).
// force_pair: -1 = heuristic (SDEO_PAIR=0/1 overrides), 0 = off, 1 = on (fails with fewer than two M tiles).
static bool make_plan(const sdeo_conv_args* a, ConvPlan* pl, int force_bn = 0, int force_splits = 0, int force_halo = -1,
                      int force_pair = -1, int force_occ2 = -1) {
  const bool up2 = a->up2_phase != 0;
  if (up2) {  // 2x2 phase filter of upsample + conv3x3: no other geometry options, no fused statistics / LayerNorm / GroupNorm
    if (a->ksize != 2 || a->stride != 1 || a->pad_hi != 0 || a->up2_phase < 1 || a->up2_phase > 4 || a->ln_stats || a->gnf_stats1 ||
        a->epi_mode != SDEO_EPI_NORMAL)
      return false;
  }
  if (!(a->ksize == 1 || a->ksize == 3 || up2)) return false;
  if (!(a->stride == 1 || a->stride == 2)) return false;
  // symmetric "same" padding, or the VAE encoder's Downsample: 3x3 stride 2 over F.pad(x, (0,1,0,1)) = no leading padding,
  // one trailing zero row / column (TMA out-of-bounds fill supplies it like every other padding pixel)
  const bool tail_pad = a->ksize == 3 && a->stride == 2 && a->pad == 0 && a->pad_hi == 1;
  if (!tail_pad && !up2 && (a->pad != (a->ksize == 3 ? 1 : 0) || a->pad_hi != 0)) return false;
  if (a->c1 <= 0 || (a->ld1 % 8) != 0 || (a->x2 && (a->ld2 % 8) != 0)) return false;
  if (a->x2 && (a->c1 % 64) != 0) return false;
  pl->Ho = (a->h + 2 * a->pad + a->pad_hi - a->ksize) / a->stride + 1;
  pl->Wo = (a->w + 2 * a->pad + a->pad_hi - a->ksize) / a->stride + 1;
  if (up2) { pl->Ho = a->h; pl->Wo = a->w; }   // one output pixel per input pixel and phase
  // ---- M tile box: minimise tile count, then prefer wide boxes ----
  int best_tiles = INT32_MAX, bbn = 1, bbh = 1, bbw = 1;
  const int maxw = pl->Wo < kBM ? pl->Wo : kBM;
  for (int bw = maxw; bw >= 1; --bw) {
    const int maxh = (kBM / bw) < pl->Ho ? (kBM / bw) : pl->Ho;
    for (int bh = maxh; bh >= 1; --bh) {
      int bn = 1;
      if (bh == pl->Ho && bw == pl->Wo) {
        bn = kBM / (bh * bw);
        if (bn > a->n) bn = a->n;
        if (bn < 1) bn = 1;
      }
      if (bw * a->stride > 256 || bh * a->stride > 256) continue;
      const int tiles = ((a->n + bn - 1) / bn) * ((pl->Ho + bh - 1) / bh) * ((pl->Wo + bw - 1) / bw);
      if (tiles < best_tiles) {
        best_tiles = tiles; bbn = bn; bbh = bh; bbw = bw;
      }
    }
  }
  // ---- HALO tile: (bh x bw) output pixels laid out with pitch bw+2; rows used = (bh-1)*(bw+2) + bw <= 128 ----
  pl->halo = 0; pl->hpitch = 0; pl->a_stages = 0; pl->a_stage_bytes = 0;
  {
    int want = force_halo;
    if (want < 0) {
      if (const char* e = getenv("SDEO_HALO")) want = atoi(e) ? 1 : 0;
    }
    const bool eligible = a->ksize == 3 && a->stride == 1 && a->pad == 1 && a->pad_hi == 0 && bbn == 1;
    if (want == 1 && !eligible) return false;
    if (eligible && want != 0) {
      int h_tiles = INT32_MAX, hbh = 0, hbw = 0;
      long long h_area = 0;
      for (int bw = (pl->Wo < 126 ? pl->Wo : 126); bw >= 4; --bw) {
        int bh = (kBM - bw) / (bw + 2) + 1;
        if (bh > pl->Ho) bh = pl->Ho;
        if (bh < 1) continue;
        const int tiles = a->n * ((pl->Ho + bh - 1) / bh) * ((pl->Wo + bw - 1) / bw);
        const long long area = (long long)(bh + 2) * (bw + 2);   // halo pixels fetched per tile and chunk
        if (tiles < h_tiles || (tiles == h_tiles && area < h_area)) { h_tiles = tiles; hbh = bh; hbw = bw; h_area = area; }
      }
      // the heuristic accepts up to 1/3 more M tiles than the tap-by-tap tiling (the A operand shrinks ~5x)
      // (a folded GroupNorm normalises every A tile it stages: once per chunk here, once per TAP and chunk otherwise)
      if (hbw > 0 && (want == 1 || 3LL * h_tiles <= 4LL * best_tiles || (a->gnf_stats1 && 2LL * h_tiles <= 4LL * best_tiles))) {
        pl->halo = 1; pl->hpitch = hbw + 2;
        bbn = 1; bbh = hbh; bbw = hbw;
        best_tiles = h_tiles;
        // the MMA reads 128 rows from row offset up to 2*pitch+2: the stage covers that even where the box is smaller
        int rows = (hbh + 2) * (hbw + 2);
        const int touched = 2 * (hbw + 2) + 2 + kBM;
        if (rows < touched) rows = touched;
        pl->a_stage_bytes = round_up(rows * 128, 1024);
      } else if (want == 1) {
        return false;
      }
    }
  }
  pl->bn_ = bbn; pl->bh = bbh; pl->bw = bbw;
  pl->rows_valid = pl->halo ? (bbh - 1) * pl->hpitch + bbw : bbn * bbh * bbw;
  pl->tiles_n = (a->n + bbn - 1) / bbn;
  pl->tiles_h = (pl->Ho + bbh - 1) / bbh;
  pl->tiles_w = (pl->Wo + bbw - 1) / bbw;
  // ---- K ----
  pl->c1c = (a->c1 + 63) / 64;
  pl->c2c = a->x2 ? (a->c2 + 63) / 64 : 0;
  pl->cpt = pl->c1c + pl->c2c;
  pl->total_chunks = a->ksize * a->ksize * pl->cpt;
  // ---- N tile ----
  pl->rows_packed = round_up(a->cout, 16);
  pl->BN = pick_bn_impl(pl->rows_packed, a->epi_mode);
  if (a->epi_mode != SDEO_EPI_GEGLU && pl->rows_packed > 128 && best_tiles <= 2) {
    // weight-streaming layers with very few M tiles: narrower N tiles put more SMs on the weight stream
    // (cluster split-K is capped at kMaxCluster K-slices per tile)
    static const int cand[] = {256, 192, 160, 128, 96, 80, 64};
    for (int bn : cand) {
      if (bn > pl->BN || pl->rows_packed % bn != 0) continue;
      pl->BN = bn;
      if (best_tiles * (pl->rows_packed / bn) * kMaxCluster >= 120 || pl->total_chunks < 32) break;
    }
  }
  if (force_bn > 0 && a->epi_mode != SDEO_EPI_GEGLU) pl->BN = force_bn;
  if (const char* e = getenv("SDEO_FORCE_BN")) {  // tuning aid; GEGLU tiles are fixed by the weight packing
    const int f = atoi(e);
    if (a->epi_mode != SDEO_EPI_GEGLU && f >= 16 && f <= 256 && f % 16 == 0 && pl->rows_packed % f == 0) pl->BN = f;
  }
  if (a->row_stats && a->epi_mode == SDEO_EPI_NORMAL && pl->BN != 64 && pl->BN != 128 && pl->BN != 256) {
    if (force_bn > 0) return false;  // (the autotuner skips this candidate)
    int pick = 0;
    for (int bn : {256, 128, 64})
      if (!pick && pl->rows_packed % bn == 0) pick = bn;
    if (!pick) return false;
    pl->BN = pick;
  }
  if (pl->BN <= 0 || pl->BN > 256 || (pl->BN % 16) != 0) return false;
  pl->n_tiles = pl->rows_packed / pl->BN;
  // ---- CTA pairs ----
  {
    int want = force_pair;
    if (want < 0) {
      if (const char* e = getenv("SDEO_PAIR")) want = atoi(e) ? 1 : 0;
    }
    if (want == 1 && best_tiles < 2) return false;
    // (folded GroupNorm: every CTA normalises the A tiles it stages after waiting on its OWN barrier; in a pair the
    //  partner's loads complete on the leader's barrier, which a remote CTA cannot wait on)
    if (want == 1 && a->gnf_stats1) return false;
    pl->pair = (best_tiles >= 2 && want != 0 && !a->gnf_stats1) ? 1 : 0;
    // a forced K-slice count beyond what fits a cluster next to the pair wins over the pair HEURISTIC
    int fs = force_splits;
    if (const char* e = getenv("SDEO_FORCE_SPLITS")) fs = atoi(e);
    if (pl->pair && want < 0 && fs > kMaxCluster / 2) pl->pair = 0;
    if (pl->pair) best_tiles = (best_tiles + 1) & ~1;
  }
  const int max_splits = pl->pair ? kMaxCluster / 2 : kMaxCluster;
  // ---- split-K over a thread-block cluster ----
  const int base = best_tiles * pl->n_tiles;
  int splits = 1;
  if ((base <= 40 || (pl->total_chunks >= 80 && base <= 74)) && a->epi_mode != SDEO_EPI_QKV) {  // (the V^T scatter reads the local tile only) measured: with >= 48 tiles the cluster reduction costs more than the extra SMs give back
    splits = cta_limit() / base;  // one CTA per SM: never spill into a second wave
    const int max_by_k = pl->total_chunks / 4;  // at least 4 K chunks per slice
    if (splits > max_by_k) splits = max_by_k;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (force_splits > 0 && a->epi_mode != SDEO_EPI_QKV) {
    if (force_splits > max_splits) return false;  // (the autotuner skips this candidate)
    splits = force_splits > pl->total_chunks ? pl->total_chunks : force_splits;
  }
  if (const char* e = getenv("SDEO_FORCE_SPLITS")) {  // tuning aid (tools/bench_conv.py)
    const int f = atoi(e);
    if (f > max_splits) return false;
    if (f >= 1 && a->epi_mode != SDEO_EPI_QKV) splits = f > pl->total_chunks ? pl->total_chunks : f;
  }
  pl->cps = (pl->total_chunks + splits - 1) / splits;
  pl->splits = (pl->total_chunks + pl->cps - 1) / pl->cps;  // every slice gets >= 1 chunk
  // ---- smem / tmem ----
  // Two CTAs per SM (occ2): grids of several waves whose tiles pay a long epilogue (TMEM drain + ~30 B/clk of stores) --
  // a second resident CTA runs its mainloop meanwhile. Needs <= 112 KB per CTA and <= 256 TMEM columns (always true).
  // force_occ2: -1 = heuristic (SDEO_OCC2=0/1 overrides), 0 / 1.
  int want_occ2 = force_occ2;
  if (want_occ2 < 0) {
    if (const char* e = getenv("SDEO_OCC2")) want_occ2 = atoi(e) ? 1 : 0;
  }
  {
    const long long ctas = (long long)best_tiles * pl->n_tiles * pl->splits;
    if (want_occ2 < 0) want_occ2 = (ctas >= 2 * 148 && pl->splits == 1) ? 1 : 0;
    if (pl->splits > 1) want_occ2 = 0;   // (split-K clusters keep one CTA per SM)
  }
  pl->occ2 = want_occ2;
  const int kSmemMax = pl->occ2 ? 112 * 1024 : 227 * 1024, kFixed = 5120;  // 1 KB alignment slack + 4 KB barriers / row tables / column vectors
  const int b_stage = (pl->pair ? pl->BN / 2 : pl->BN) * 128;   // weight rows one CTA stages per K step
  const int stage_bytes = pl->halo ? b_stage : kATileBytes + b_stage;   // HALO: the B ring's stage
  const int rows_valid = pl->rows_valid;
  // fp32 epilogue tile, aliases the pipeline stages. Split-K: the tile (valid rows only) plus the receive area
  // [S][ceil(rows/S)][LD] the peers' bulk copies land in.
  auto tile_bytes_for = [&](int sp) {
    const int rp = (rows_valid + sp - 1) / sp;
    return (sp > 1 ? rows_valid + sp * rp : kBM) * (pl->BN + 4) * 4;
  };
  if (pl->splits > 1 && kFixed + tile_bytes_for(pl->splits) > kSmemMax) {
    if (force_splits > 0 || getenv("SDEO_FORCE_SPLITS")) return false;  // (the autotuner skips this candidate)
    pl->splits = 1;
    pl->cps = pl->total_chunks;
  }
  const int tile_bytes = tile_bytes_for(pl->splits);
  // HALO: two A stages (three when at least six weight stages still fit)
  auto a_ring_for = [&](int extra) {
    if (!pl->halo) return 0;
    const int chunks = (pl->cps + 8) / 9 + 1;   // halo tiles a K slice can touch
    int ast = chunks < 2 ? chunks : 2;
    if (chunks >= 3 && (kSmemMax - kFixed - extra - 3 * pl->a_stage_bytes) / stage_bytes >= 6) ast = 3;
    return ast;
  };
  pl->gnf_off = 0; pl->gnf_bytes = 0;
  if (a->gnf_stats1) {
    if (a->ln_stats || a->gnf_groups <= 0 || (a->c1 + (a->x2 ? a->c2 : 0)) % a->gnf_groups != 0 || !a->gnf_gamma || !a->gnf_beta ||
        (a->x2 && !a->gnf_stats2) || pl->bn_ > 128 || pl->bh > 255 || pl->bw > 255)
      return false;
    pl->gnf_bytes = round_up(pl->bn_ * (pl->cpt * 80 + a->gnf_groups) * 8, 128);   // (a, b) rows padded 8 -> 10 entries
  }
  const int gnf_bytes = pl->gnf_bytes;
  auto stages_for = [&](int extra0) {  // pipeline depth that fits next to `extra` bytes of residual buffer (+ the folded GroupNorm table)
    const int extra = extra0 + gnf_bytes;
    if (kFixed + tile_bytes + extra > kSmemMax) return 0;
    const int a_ring = a_ring_for(extra) * pl->a_stage_bytes;
    int st = (kSmemMax - kFixed - extra - a_ring) / stage_bytes;
    if (st > (pl->halo ? 12 : 8)) st = pl->halo ? 12 : 8;
    if (st > pl->cps) st = pl->cps < 2 ? 2 : pl->cps;
    return st;
  };
  int stages = stages_for(0);
  if (pl->occ2 && (stages < 3 || (pl->halo && a_ring_for(gnf_bytes) < 2))) {
    // too shallow a ring (or no room for the epilogue tile) at half the shared memory
    if (force_occ2 == 1) return false;
    return make_plan(a, pl, force_bn, force_splits, force_halo, force_pair, 0);
  }
  if (stages < 2) return false;
  // residual tile prefetched into shared memory by the otherwise idle warps, if enough pipeline stages still fit
  pl->res_smem_off = 0;
  int res_bytes = 0;
  // (only the vector-aligned NORMAL epilogue reads the prefetched tile; the same conditions make the TMA box legal)
  bool res_fast = a->residual && !up2 && a->epi_mode == SDEO_EPI_NORMAL && (a->cout % 16 == 0) &&
                  (a->residual_f32 ? (a->ldr % 4 == 0) : (a->ldr % 8 == 0)) && ((reinterpret_cast<uintptr_t>(a->residual) & 15) == 0);
  res_fast = res_fast && (a->y_fp32 ? (a->ldy % 4 == 0) : (a->ldy % 8 == 0)) && (!(a->y_fp32 && a->y2) || (a->ldy2 % 8 == 0));
  if (res_fast && !getenv("SDEO_NO_RES_PREFETCH")) {
    const int res_rows = pl->halo ? pl->hpitch * pl->bh : kBM;
    res_bytes = (res_rows > kBM ? res_rows : kBM) * pl->BN * (a->residual_f32 ? 4 : 2);
    const int st2 = stages_for(res_bytes + 128);
    const int need = pl->cps < 4 ? (pl->cps < 2 ? 2 : pl->cps) : 4;   // (measured: a 3-stage ring costs more than the residual prefetch saves)
    if (st2 >= need) stages = st2; else res_bytes = 0;
  }
  pl->nprod = stages < kProducers ? stages : kProducers;
  stages -= stages % pl->nprod;
  pl->stages = stages;
  pl->a_stages = a_ring_for(res_bytes + gnf_bytes);
  size_t body = (size_t)stages * stage_bytes + (size_t)pl->a_stages * pl->a_stage_bytes;
  if (body < (size_t)tile_bytes) body = tile_bytes;
  body = (body + 127) & ~(size_t)127;   // (the residual tile behind it is a TMA destination)
  if (res_bytes) pl->res_smem_off = 4096 + (int)body;
  if (gnf_bytes) pl->gnf_off = 4096 + (int)body + ((res_bytes + 127) & ~127);
  pl->smem_bytes = kFixed + body + ((res_bytes + 127) & ~127) + gnf_bytes;
  int tc = 32;
  while (tc < pl->BN) tc *= 2;  // fp32 accumulator columns (power of two >= 32)
  pl->tmem_cols = tc;
  return pl->smem_bytes <= (size_t)kSmemMax;
}

static inline int cfg_tiles(const ConvPlan& pl) { return pl.tiles_n * pl.tiles_h * pl.tiles_w * pl.n_tiles; }

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_conv)

extern "C" size_t sdeo_conv_counter_bytes(void) { return 0; }

extern "C" size_t sdeo_conv_workspace_bytes(const sdeo_conv_args* a) {
  (void)a;  // split-K partial tiles: at most one 128 x (256+4) fp32 tile per resident CTA (one CTA per SM, 148 SMs)
  return (size_t)160 * kBM * (256 + 4) * sizeof(float);
}

extern "C" int32_t sdeo_packed_rows(int32_t cout) { return round_up(cout, 16); }
extern "C" int32_t sdeo_packed_k(int32_t c1, int32_t c2, int32_t ksize) {
  return ksize * ksize * (((c1 + 63) / 64) + ((c2 + 63) / 64)) * 64;
}
extern "C" int32_t sdeo_pick_bn(int32_t rows_packed, int32_t epi_mode, int32_t dhead) {
  (void)dhead;
  return pick_bn_impl(rows_packed, epi_mode);
}

extern "C" int sdeo_pack_conv_weight(const float* w, int32_t cout, int32_t c1, int32_t c2, int32_t ksize,
                                     int32_t geglu_bn, void* w_packed, void* stream) {
  if (!w || !w_packed || cout <= 0 || c1 <= 0 || c2 < 0) return set_error(SDEO_EINVAL, "pack_conv_weight: bad args");
  if (c2 > 0 && (c1 % 64) != 0) return set_error(SDEO_EINVAL, "pack_conv_weight: c1 must be a multiple of 64 when c2 > 0");
  const int rows = round_up(cout, 16);
  if (geglu_bn > 0 && (rows % geglu_bn != 0 || (cout / 2) % (geglu_bn / 2) != 0))
    return set_error(SDEO_EINVAL, "pack_conv_weight: geglu tile does not divide rows");
  const long long total = (long long)rows * sdeo_packed_k(c1, c2, ksize);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_conv_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)w_packed, cout, c1, c2, ksize,
                                                                    rows, geglu_bn);
  return check_launch("pack_conv_weight");
}

extern "C" int sdeo_pack_geglu_bias(const float* b, int32_t n2, int32_t geglu_bn, float* b_packed, void* stream) {
  if (!b || !b_packed || n2 <= 0 || geglu_bn <= 0 || n2 % geglu_bn != 0) return set_error(SDEO_EINVAL, "pack_geglu_bias: bad args");
  pack_geglu_bias_kernel<<<(n2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b, b_packed, n2, geglu_bn);
  return check_launch("pack_geglu_bias");
}

static int launch_conv(const sdeo_conv_args* a, const ConvPlan& pl, void* stream);

// GroupNorm partial statistics (sdeo_conv_args::gn_stats): produced by the vector-aligned NORMAL epilogue with an fp32
// output when no M tile spans two samples. Returns the number of partial slots per sample (M tiles per sample x K
// slices) under plan `pl`, 0 if this call does not produce them.
// Row statistics (sdeo_conv_args::row_stats) need the vector-aligned NORMAL epilogue with an fp32 output and an N tile of
// 64 / 128 / 256 columns (the lanes that hold one row must form an aligned power-of-two group of a warp).
static bool row_stats_ok(const sdeo_conv_args* a, const ConvPlan& pl) {
  if (!a->row_stats || a->epi_mode != SDEO_EPI_NORMAL || !a->y_fp32 || a->up2_phase) return false;
  bool fast = (a->cout % 16 == 0) && (a->ldy % 4 == 0);
  if (a->y2) fast = fast && (a->ldy2 % 8 == 0);
  if (a->residual) fast = fast && (a->residual_f32 ? (a->ldr % 4 == 0) : (a->ldr % 8 == 0));
  return fast && (pl.BN == 64 || pl.BN == 128 || pl.BN == 256);
}

static int stats_parts(const sdeo_conv_args* a, const ConvPlan& pl) {
  if (!a->gn_stats || a->row_stats || a->epi_mode != SDEO_EPI_NORMAL || a->up2_phase) return 0;
  bool fast = (a->cout % 16 == 0) && (a->y_fp32 ? (a->ldy % 4 == 0) : (a->ldy % 8 == 0));
  if (a->y2) fast = fast && (a->ldy2 % 8 == 0);
  if (a->residual) fast = fast && (a->residual_f32 ? (a->ldr % 4 == 0) : (a->ldr % 8 == 0));
  if (!fast) return 0;
  if (pl.bn_ == 1) return pl.tiles_h * pl.tiles_w * pl.splits;
  // several whole samples per tile: only as a split-K cluster whose per-rank row ranges stay inside one sample
  const int per_img = pl.bh * pl.bw, rows_valid = pl.bn_ * per_img;
  const int rows_per = (rows_valid + pl.splits - 1) / pl.splits;
  if (pl.splits > 1 && per_img % rows_per == 0) return per_img / rows_per;
  return 0;
}

// ---- per-shape autotuning of (N tile, K slices) -------------------------------------------------------------
// The best tile / split-K choice depends on the layer shape in ways the heuristic does not capture (measured: 2 K
// slices speed up the 3x3 convs at M=3072 by 20-30% and slow the GEGLU linears down 2x). With autotuning enabled
// the first eager call of a shape times the candidates on the caller's stream and caches the winner; calls made
// during CUDA-graph capture only read the cache.
#include <map>
#include <array>
#include <mutex>
namespace {
typedef std::array<int, 16> TuneKey;
struct Tuned { int first, second, halo, pair, occ2; };   // N tile, K slices, HALO mode, CTA pairs, two CTAs per SM
std::map<TuneKey, Tuned> g_tuned;
std::mutex g_tune_mu;
int g_autotune = 0;

TuneKey tune_key(const sdeo_conv_args* a) {
  TuneKey k = {a->n, a->h, a->w, a->c1, a->x2 ? a->c2 : 0, a->cout, a->ksize, a->stride | (a->pad_hi << 4) | (a->up2_phase << 8), a->epi_mode, a->y_fp32,
               a->residual ? (a->residual_f32 ? 2 : 1) : 0, a->y2 ? 1 : 0, a->emb ? 1 : 0, a->act, a->dhead,
               (a->gn_stats ? 1 : 0) | (a->row_stats ? 2 : 0) | (a->ln_stats ? 4 : 0) | (cta_limit() << 3) |
                   (a->gnf_stats1 ? (1 << 12) | (a->gnf_silu ? 1 << 13 : 0) : 0)};
  return k;
}

// true if some (BN, 1) candidate of this shape fits the CTA budget (else the budget cannot be honoured at all)
bool any_within_budget(const sdeo_conv_args* a, const ConvPlan& base) {
  static const int bns[] = {256, 192, 160, 128, 96, 80, 64};
  if (base.tiles_n * base.tiles_h * base.tiles_w * base.n_tiles <= cta_limit()) return true;
  if (a->epi_mode == SDEO_EPI_GEGLU) return false;
  for (int bn : bns) {
    ConvPlan pl;
    if (base.rows_packed % bn != 0 || !make_plan(a, &pl, bn, 1) || pl.BN != bn) continue;
    if (pl.tiles_n * pl.tiles_h * pl.tiles_w * pl.n_tiles <= cta_limit()) return true;
  }
  return false;
}

bool tune_shape(const sdeo_conv_args* a, void* stream, Tuned* best) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return false;
  ConvPlan base;
  if (!make_plan(a, &base)) return false;
  static const int bns[] = {0, 256, 192, 160, 128, 96, 80, 64};
  static const int ss[] = {1, 2, 3, 4, 6, 8};
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return false;
  static const size_t kFlushBytes = (size_t)256 << 20;  // > 126 MB L2
  static void* flush = nullptr;
  static bool flush_tried = false;
  if (!flush_tried) {
    flush_tried = true;
    if (getenv("SDEO_TUNE_WARM") || cudaMalloc(&flush, kFlushBytes) != cudaSuccess) { flush = nullptr; (void)cudaGetLastError(); }
  }
  float best_ms = 1e30f;
  *best = Tuned{base.BN, base.splits, base.halo, base.pair, base.occ2};
  const bool no_halo_tune = getenv("SDEO_HALO") != nullptr;   // forced on / off: tune within that mode only
  const bool no_pair_tune = getenv("SDEO_PAIR") != nullptr;
  const bool no_occ_tune = getenv("SDEO_OCC2") != nullptr;
  for (int occ2 = 0; occ2 < 2; ++occ2)
  for (int pair = 0; pair < 2; ++pair)
  for (int halo = 0; halo < 2; ++halo)
  for (int bn : bns) {
    if (bn == 0) bn = base.BN;
    else if (a->epi_mode == SDEO_EPI_GEGLU || bn == base.BN || base.rows_packed % bn != 0) continue;
    for (int sp : ss) {
      if (a->epi_mode == SDEO_EPI_QKV && sp > 1) continue;
      ConvPlan pl;
      if ((no_halo_tune && halo != base.halo) || (no_pair_tune && pair != base.pair)) continue;
      if (no_occ_tune && occ2 != base.occ2) continue;
      if (!make_plan(a, &pl, bn, sp, no_halo_tune ? -1 : halo, no_pair_tune ? -1 : pair, no_occ_tune ? -1 : occ2) || pl.BN != bn ||
          pl.splits != sp || pl.halo != halo || pl.pair != pair || pl.occ2 != occ2)
        continue;
      if (occ2 && (long long)pl.tiles_n * pl.tiles_h * pl.tiles_w * pl.n_tiles <= 148) continue;   // nothing to co-schedule
      int mt = pl.tiles_n * pl.tiles_h * pl.tiles_w;
      if (pl.pair) mt = (mt + 1) & ~1;
      const int ctas = mt * pl.n_tiles * pl.splits;
      if (sp > 1 && ctas > cta_limit()) continue;  // K slices must not spill into a second wave
      if (g_cta_budget > 0 && ctas > cta_limit() && any_within_budget(a, base)) continue;  // honour the CTA budget
      if (launch_conv(a, pl, stream) != 0) {
        (void)cudaGetLastError();
        if (getenv("SDEO_TUNE_VERBOSE"))
          fprintf(stderr, "sdeo tune: candidate bn %d splits %d halo %d pair %d occ2 %d failed to launch: %s\n", bn, sp, halo, pair, occ2,
                  sdeo_last_error());
        continue;
      }
      if (getenv("SDEO_TUNE_VERBOSE")) {   // debugging aid: run every candidate to completion and name the one that faults
        const cudaError_t ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
          fprintf(stderr, "sdeo tune: candidate bn %d splits %d halo %d pair %d occ2 %d (n %d h %d w %d c1 %d cout %d k %d) FAULTED: %s\n", bn, sp,
                  halo, pair, occ2, a->n, a->h, a->w, a->c1, a->cout, a->ksize, cudaGetErrorString(ce));
          return false;
        }
      }
      // Timed COLD: inside a denoising step every layer's weights come from HBM (2.4 GB are streamed per step, the L2
      // holds 126 MB), so the L2 is flushed before each timed launch. Back-to-back launches of one layer would measure
      // L2-resident weights and favour configurations with few CTAs on the weight stream.
      float ms = 1e30f;  // best of 5 (the minimum is robust against interference from other streams)
      bool ok = true;
      for (int r = 0; r < 5 && ok; ++r) {
        if (flush) cudaMemsetAsync(flush, r, kFlushBytes, st);
        cudaEventRecord(e0, st);
        ok = launch_conv(a, pl, stream) == 0;
        cudaEventRecord(e1, st);
        float t = 0.f;
        ok = ok && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&t, e0, e1) == cudaSuccess;
        if (t < ms) ms = t;
      }
      if (!ok) { (void)cudaGetLastError(); continue; }
      // A plan that cannot leave the GroupNorm statistics the caller asked for (a tile that spans samples without a K split
      // on sample boundaries, stats_parts) sends the consumer to the standalone cluster GroupNorm: ~8 us per launch in the
      // step instead of ~4 for the statistics-fed apply. Charge that difference to the candidate.
      if (a->gn_stats && stats_parts(a, pl) == 0) ms += 0.004f;
      if (ms < best_ms) { best_ms = ms; *best = Tuned{bn, sp, halo, pair, occ2}; }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return true;
}
}  // namespace

extern "C" int sdeo_conv_set_cta_budget(int max_ctas) {
  g_cta_budget = max_ctas;
  return SDEO_OK;
}

extern "C" int sdeo_conv_autotune(int enable) {
  g_autotune = enable;
  return SDEO_OK;
}

// Resolves the plan sdeo_conv2d uses for these args: the autotuned (N tile, K slices) if the shape has been tuned
// (tuning it first when `stream` is given and autotuning is on), the heuristic otherwise.
static bool resolve_plan(const sdeo_conv_args* a, void* stream, bool may_tune, ConvPlan* pl) {
  int force_bn = 0, force_s = 0, force_halo = -1, force_pair = -1, force_occ2 = -1;
  if (g_autotune && !getenv("SDEO_FORCE_BN") && !getenv("SDEO_FORCE_SPLITS")) {
    std::lock_guard<std::mutex> lock(g_tune_mu);
    const TuneKey key = tune_key(a);
    auto it = g_tuned.find(key);
    if (it == g_tuned.end() && may_tune) {
      Tuned best;
      if (tune_shape(a, stream, &best)) it = g_tuned.emplace(key, best).first;
    }
    if (it != g_tuned.end()) {
      force_bn = it->second.first; force_s = it->second.second;
      if (!getenv("SDEO_HALO")) force_halo = it->second.halo;
      if (!getenv("SDEO_PAIR")) force_pair = it->second.pair;
      if (!getenv("SDEO_OCC2")) force_occ2 = it->second.occ2;
    }
  }
  return make_plan(a, pl, force_bn, force_s, force_halo, force_pair, force_occ2);
}

extern "C" int sdeo_conv2d(const sdeo_conv_args* a, void* stream) {
  if (!a || !a->x1 || !a->w_packed) return set_error(SDEO_EINVAL, "conv2d: null argument");
  ConvPlan pl;
  if (!resolve_plan(a, stream, true, &pl)) return set_error(SDEO_EINVAL, "conv2d: unsupported geometry");
  return launch_conv(a, pl, stream);
}

extern "C" int sdeo_conv_row_stats_parts(const sdeo_conv_args* a, int32_t* max_parts, int32_t* parts) {
  if (!a) return set_error(SDEO_EINVAL, "conv_row_stats_parts: null argument");
  ConvPlan pl;
  sdeo_conv_args b = *a;
  if (!b.row_stats) b.row_stats = (float*)(uintptr_t)16;  // "would be produced if a buffer were given"
  if (!resolve_plan(&b, nullptr, false, &pl)) return set_error(SDEO_EINVAL, "conv_row_stats_parts: unsupported geometry");
  if (max_parts) *max_parts = (pl.rows_packed + 63) / 64;
  if (parts) *parts = row_stats_ok(&b, pl) ? pl.n_tiles : 0;
  return SDEO_OK;
}

extern "C" int sdeo_conv_gn_stats_slots(const sdeo_conv_args* a, int32_t* max_slots_total, int32_t* parts_per_sample) {
  if (!a) return set_error(SDEO_EINVAL, "conv_gn_stats_slots: null argument");
  ConvPlan pl;
  if (!resolve_plan(a, nullptr, false, &pl)) return set_error(SDEO_EINVAL, "conv_gn_stats_slots: unsupported geometry");
  if (max_slots_total) {
    // upper bound over every plan the autotuner may still pick for this shape (the HALO tiling has its own tile count)
    int tiles = pl.tiles_n * pl.tiles_h * pl.tiles_w;
    for (int halo = 0; halo < 2; ++halo) {
      ConvPlan alt;
      if (make_plan(a, &alt, 0, 0, halo) && alt.tiles_n * alt.tiles_h * alt.tiles_w > tiles)
        tiles = alt.tiles_n * alt.tiles_h * alt.tiles_w;
    }
    // (K slices only exist where all CTAs of the launch fit one wave: see make_plan and the autotuner's candidate filter)
    *max_slots_total = tiles * ((tiles > 148 && !getenv("SDEO_FORCE_SPLITS")) ? 1 : kMaxCluster);
  }
  if (parts_per_sample) {
    sdeo_conv_args b = *a;
    if (!b.gn_stats) b.gn_stats = (float*)(uintptr_t)16;  // "would be produced if a buffer were given"
    *parts_per_sample = stats_parts(&b, pl);
  }
  return SDEO_OK;
}

extern "C" int sdeo_conv_plan_describe(const sdeo_conv_args* a, int32_t halo, int32_t* out, int32_t n_out) {
  if (!a || !out || n_out < 16) return set_error(SDEO_EINVAL, "conv_plan_describe: bad args");
  ConvPlan pl;
  const bool ok = halo < 0 ? resolve_plan(a, nullptr, false, &pl) : make_plan(a, &pl, 0, 0, halo);
  if (!ok) return set_error(SDEO_EINVAL, "conv_plan_describe: unsupported geometry");
  const int32_t v[16] = {pl.BN, pl.splits, pl.halo | (pl.pair << 1) | (pl.occ2 << 2) | (pl.nprod << 4), pl.bn_, pl.bh, pl.bw, pl.tiles_n * pl.tiles_h * pl.tiles_w, pl.n_tiles,
                         pl.stages, pl.a_stages, pl.a_stage_bytes, (int32_t)pl.smem_bytes, pl.rows_valid, pl.tmem_cols,
                         pl.hpitch, pl.cps};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return SDEO_OK;
}

static int launch_conv(const sdeo_conv_args* a, const ConvPlan& pl, void* stream) {
  if (a->epi_mode == SDEO_EPI_QKV) {
    if (!a->vt && !a->q && !a->k) return set_error(SDEO_EINVAL, "conv2d: qkv outputs missing");
    if ((a->dhead % 8) != 0 || a->heads <= 0 || a->tokens <= 0) return set_error(SDEO_EINVAL, "conv2d: bad qkv geometry");
    if ((a->cout % 8) != 0) return set_error(SDEO_EINVAL, "conv2d: qkv cout must be a multiple of 8");
  } else if (!a->y) {
    return set_error(SDEO_EINVAL, "conv2d: null output");
  }
  if (a->epi_mode == SDEO_EPI_GEGLU && ((a->cout % 32) != 0 || (a->ldy % 8) != 0))
    return set_error(SDEO_EINVAL, "conv2d: bad geglu geometry");

  CUtensorMap tmA1, tmA2, tmB;
  const uint32_t st = (uint32_t)a->stride;
  {
    uint64_t dims[4] = {(uint64_t)a->c1, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
    uint64_t strides[3] = {(uint64_t)a->ld1 * 2, (uint64_t)a->w * a->ld1 * 2, (uint64_t)a->h * a->w * a->ld1 * 2};
    uint32_t box[4] = {64, (uint32_t)pl.bw * st, (uint32_t)pl.bh * st, (uint32_t)pl.bn_};
    if (pl.halo) { box[1] = (uint32_t)pl.bw + 2; box[2] = (uint32_t)pl.bh + 2; }
    uint32_t es[4] = {1, st, st, 1};
    int rc = encode_tmap_bf16(&tmA1, a->x1, 4, dims, strides, box, es);
    if (rc) return rc;
    tmA2 = tmA1;
  }
  if (a->x2) {
    uint64_t dims[4] = {(uint64_t)a->c2, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
    uint64_t strides[3] = {(uint64_t)a->ld2 * 2, (uint64_t)a->w * a->ld2 * 2, (uint64_t)a->h * a->w * a->ld2 * 2};
    uint32_t box[4] = {64, (uint32_t)pl.bw * st, (uint32_t)pl.bh * st, (uint32_t)pl.bn_};
    if (pl.halo) { box[1] = (uint32_t)pl.bw + 2; box[2] = (uint32_t)pl.bh + 2; }
    uint32_t es[4] = {1, st, st, 1};
    int rc = encode_tmap_bf16(&tmA2, a->x2, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    const uint64_t kp = (uint64_t)a->ksize * a->ksize * pl.cpt * 64;
    uint64_t dims[2] = {kp, (uint64_t)pl.rows_packed};
    uint64_t strides[1] = {kp * 2};
    uint32_t box[2] = {64, (uint32_t)(pl.pair ? pl.BN / 2 : pl.BN)};
    uint32_t es[2] = {1, 1};
    int rc = encode_tmap_bf16(&tmB, a->w_packed, 2, dims, strides, box, es);
    if (rc) return rc;
  }

  ConvKParams p;
  p.kw = a->ksize; p.pad = a->pad; p.stride = a->stride;
  p.pad_h = p.pad_w = a->pad; p.out_mul = 1; p.out_off_h = p.out_off_w = 0; p.Ho_full = pl.Ho; p.Wo_full = pl.Wo;
  if (a->up2_phase) {
    const int ph_h = (a->up2_phase - 1) >> 1, ph_w = (a->up2_phase - 1) & 1;
    p.pad_h = 1 - ph_h; p.pad_w = 1 - ph_w; p.pad = 0;
    p.out_mul = 2; p.out_off_h = ph_h; p.out_off_w = ph_w; p.Ho_full = 2 * pl.Ho; p.Wo_full = 2 * pl.Wo;
  }
  p.c1_chunks = pl.c1c; p.chunks_per_tap = pl.cpt;
  p.total_chunks = pl.total_chunks; p.chunks_per_split = pl.cps; p.splits = pl.splits;
  p.bn_ = pl.bn_; p.bh = pl.bh; p.bw = pl.bw; p.rows_valid = pl.rows_valid;
  p.halo = pl.halo; p.hpitch = pl.hpitch; p.a_stages = pl.a_stages; p.a_stage_bytes = pl.a_stage_bytes;
  p.pair = pl.pair; p.m_tiles = pl.tiles_n * pl.tiles_h * pl.tiles_w; p.nprod = pl.nprod;
  p.probe = getenv("SDEO_NO_PROBE") ? 0 : 1;
  p.l2_prefetch = getenv("SDEO_L2_PREFETCH") ? 1 : 0;   // (measured neutral inside the step graph: opt-in)
  p.rows_per = (pl.rows_valid + pl.splits - 1) / pl.splits;
  p.cols_items = a->epi_mode == SDEO_EPI_GEGLU ? pl.BN / 16 : pl.BN / 8;
  p.step_rows = kConvThreads / p.cols_items;
  p.step_cols = kConvThreads % p.cols_items;
  p.ci_magic = (65536 + p.cols_items - 1) / p.cols_items;   // exact for threadIdx.x < 384, cols_items <= 32
  p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w;
  p.N = a->n; p.Ho = pl.Ho; p.Wo = pl.Wo;
  p.BN = pl.BN; p.cout = a->cout; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
  p.epi_mode = a->epi_mode; p.act = a->act; p.y_fp32 = a->y_fp32;
  p.bias = a->bias; p.emb = a->emb; p.emb_step = a->emb ? a->emb_step : nullptr; p.residual = a->residual; p.ldr = a->ldr;
  p.scale = a->scale; p.y = a->y; p.ldy = a->ldy;
  p.residual_f32 = a->residual_f32;
  p.y2 = (a->y_fp32 && a->epi_mode == SDEO_EPI_NORMAL) ? (__nv_bfloat16*)a->y2 : nullptr;
  p.ldy2 = a->ldy2;
  p.q = (__nv_bfloat16*)a->q; p.k = (__nv_bfloat16*)a->k; p.vt = (__nv_bfloat16*)a->vt;
  p.heads = a->heads; p.dhead = a->dhead; p.tokens = a->tokens; p.ldv = a->ldv; p.qkv_first = a->qkv_first;
  p.dbg = nullptr;
  if (const char* e = getenv("SDEO_CONV_DEBUG")) p.dbg = (long long*)strtoull(e, nullptr, 16);
  p.res_smem_off = pl.res_smem_off;
  p.res_tx = 0;
  CUtensorMap tmR = tmB;
  if (pl.res_smem_off) {
    const int eb = a->residual_f32 ? 4 : 2;
    uint64_t dims[4] = {(uint64_t)a->cout, (uint64_t)pl.Wo, (uint64_t)pl.Ho, (uint64_t)a->n};
    uint64_t strides[3] = {(uint64_t)a->ldr * eb, (uint64_t)pl.Wo * a->ldr * eb, (uint64_t)pl.Ho * pl.Wo * a->ldr * eb};
    uint32_t box[4] = {(uint32_t)pl.BN, (uint32_t)(pl.halo ? pl.hpitch : pl.bw), (uint32_t)pl.bh, (uint32_t)pl.bn_};
    int rc = encode_tmap_plain(&tmR, a->residual, eb, 4, dims, strides, box);
    if (rc) return rc;
    p.res_tx = (int)(box[0] * box[1] * box[2] * box[3]) * eb;
  }
  p.gn_stats = nullptr;
  p.row_stats = nullptr; p.row_stats_ld = 0;
  p.gnf = a->gnf_stats1 ? 1 : 0; p.gnf_silu = a->gnf_silu; p.gnf_off = pl.gnf_off;
  p.gnf_st1 = (const float2*)a->gnf_stats1; p.gnf_st2 = (const float2*)a->gnf_stats2;
  p.gnf_parts1 = a->gnf_parts1; p.gnf_parts2 = a->gnf_parts2; p.gnf_c1 = a->c1; p.gnf_c2 = a->x2 ? a->c2 : 0;
  p.gnf_groups = a->gnf_groups; p.gnf_cpg = a->gnf_groups > 0 ? (a->c1 + (a->x2 ? a->c2 : 0)) / a->gnf_groups : 0;
  p.gnf_gamma = a->gnf_gamma; p.gnf_beta = a->gnf_beta; p.gnf_eps = a->gnf_eps;
  p.gnf_inv = p.gnf ? 1.0f / ((float)a->h * (float)a->w * (float)p.gnf_cpg) : 0.f;
  p.Hin = a->h; p.Win = a->w;
  if (p.gnf && (pl.pair || !pl.gnf_off || a->gnf_parts1 <= 0 || (a->x2 && a->gnf_parts2 <= 0)))
    return set_error(SDEO_EINVAL, "conv2d: folded GroupNorm needs statistics for every source and a plan without CTA pairs");
  p.ln_stats = (const float2*)a->ln_stats; p.ln_parts = a->ln_parts; p.ln_ld = a->ln_ld;
  p.ln_invc = a->ln_c > 0 ? 1.0f / (float)a->ln_c : 0.f; p.ln_eps = a->ln_eps; p.ln_csum = a->ln_csum;
  if (p.ln_stats && (!p.ln_csum || p.ln_parts <= 0 || a->ln_c <= 0))
    return set_error(SDEO_EINVAL, "conv2d: folded LayerNorm needs ln_csum, ln_parts and ln_c");

  // ---- pick the kernel instantiation ----
#define KSEL(...)                                                                                                  \
  (pl.occ2 ? (pl.pair ? conv_gemm_kernel_occ2<__VA_ARGS__, true> : conv_gemm_kernel_occ2<__VA_ARGS__, false>)     \
           : (pl.pair ? conv_gemm_kernel<__VA_ARGS__, true> : conv_gemm_kernel<__VA_ARGS__, false>))
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvKParams);
  KernelFn fn = nullptr;
  if (a->epi_mode == SDEO_EPI_GEGLU) {
    fn = p.ln_stats ? KSEL(SDEO_EPI_GEGLU, OUT_BF16, RES_NONE, true, 0, true)
                    : KSEL(SDEO_EPI_GEGLU, OUT_BF16, RES_NONE, true, 0, false);
  } else if (a->epi_mode == SDEO_EPI_QKV) {
    fn = p.ln_stats ? KSEL(SDEO_EPI_QKV, OUT_BF16, RES_NONE, true, 0, true)
                    : KSEL(SDEO_EPI_QKV, OUT_BF16, RES_NONE, true, 0, false);
  } else {
    const int out_kind = !a->y_fp32 ? OUT_BF16 : (p.y2 ? OUT_F32_TWIN : OUT_F32);
    const int res_kind = !a->residual ? RES_NONE : (a->residual_f32 ? RES_F32 : RES_BF16);
    // (rows are packed in multiples of 16: with cout % 16 == 8 the last N tile has 8 columns beyond cout, which only the
    // generic item path masks -- e.g. the VAE encoder's 8-channel conv_out / quant_conv)
    bool fast = (a->cout % 16 == 0);
    fast = fast && (out_kind == OUT_BF16 ? (a->ldy % 8 == 0) : (a->ldy % 4 == 0));
    if (out_kind == OUT_F32_TWIN) fast = fast && (a->ldy2 % 8 == 0);
    if (res_kind == RES_BF16) fast = fast && (a->ldr % 8 == 0);
    if (res_kind == RES_F32) fast = fast && (a->ldr % 4 == 0);
    if (!fast) {
      if (p.ln_stats) return set_error(SDEO_EINVAL, "conv2d: folded LayerNorm needs the vector-aligned epilogue");
      fn = KSEL(SDEO_EPI_NORMAL, OUT_BF16, RES_NONE, false, 0, false);
    } else if (p.ln_stats) {
      // (plain-epilogue consumers of a folded LayerNorm: bf16 or fp32 output, no residual; used by the tests)
      if (res_kind != RES_NONE || out_kind == OUT_F32_TWIN)
        return set_error(SDEO_EINVAL, "conv2d: folded LayerNorm supports the QKV, GEGLU and residual-free plain epilogues");
      fn = out_kind == OUT_F32 ? KSEL(SDEO_EPI_NORMAL, OUT_F32, RES_NONE, true, 0, true)
                               : KSEL(SDEO_EPI_NORMAL, OUT_BF16, RES_NONE, true, 0, true);
    } else if (row_stats_ok(a, pl)) {
      p.row_stats = (float2*)a->row_stats;
      p.row_stats_ld = a->row_stats_ld;
#define SDEO_PICK(O, R) if (out_kind == O && res_kind == R) fn = KSEL(SDEO_EPI_NORMAL, O, R, true, 2, false);
      SDEO_PICK(OUT_F32, RES_NONE) SDEO_PICK(OUT_F32, RES_BF16) SDEO_PICK(OUT_F32, RES_F32)
      SDEO_PICK(OUT_F32_TWIN, RES_NONE) SDEO_PICK(OUT_F32_TWIN, RES_BF16) SDEO_PICK(OUT_F32_TWIN, RES_F32)
#undef SDEO_PICK
    } else if (stats_parts(a, pl) > 0) {
      p.gn_stats = (float2*)a->gn_stats;
#define SDEO_PICK(O, R) if (out_kind == O && res_kind == R) fn = KSEL(SDEO_EPI_NORMAL, O, R, true, 1, false);
      SDEO_PICK(OUT_BF16, RES_NONE) SDEO_PICK(OUT_BF16, RES_BF16) SDEO_PICK(OUT_BF16, RES_F32)
      SDEO_PICK(OUT_F32, RES_NONE) SDEO_PICK(OUT_F32, RES_BF16) SDEO_PICK(OUT_F32, RES_F32)
      SDEO_PICK(OUT_F32_TWIN, RES_NONE) SDEO_PICK(OUT_F32_TWIN, RES_BF16) SDEO_PICK(OUT_F32_TWIN, RES_F32)
#undef SDEO_PICK
    } else {
#define SDEO_PICK(O, R) if (out_kind == O && res_kind == R) fn = KSEL(SDEO_EPI_NORMAL, O, R, true, 0, false);
      SDEO_PICK(OUT_BF16, RES_NONE) SDEO_PICK(OUT_BF16, RES_BF16) SDEO_PICK(OUT_BF16, RES_F32)
      SDEO_PICK(OUT_F32, RES_NONE) SDEO_PICK(OUT_F32, RES_BF16) SDEO_PICK(OUT_F32, RES_F32)
      SDEO_PICK(OUT_F32_TWIN, RES_NONE) SDEO_PICK(OUT_F32_TWIN, RES_BF16) SDEO_PICK(OUT_F32_TWIN, RES_F32)
#undef SDEO_PICK
    }
  }
#undef KSEL
  if (!fn) return set_error(SDEO_EINVAL, "conv2d: no kernel instantiation");
  {
    // opt in to > 48 KB of dynamic shared memory, once per instantiation
    static KernelFn configured[256];
    static int n_configured = 0;
    bool seen = false;
    for (int i = 0; i < n_configured; ++i) seen = seen || (configured[i] == fn);
    if (!seen) {
      cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
      // (occ2 plans: two 112 KB CTAs per SM need the whole shared-memory carve-out)
      (void)cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (n_configured < 256) configured[n_configured++] = fn;
    }
  }
  unsigned gx = (unsigned)(pl.tiles_n * pl.tiles_h * pl.tiles_w);
  if (pl.pair) gx = (gx + 1u) & ~1u;
  return launch_k("conv2d", fn, dim3(gx, (unsigned)pl.n_tiles, (unsigned)pl.splits), dim3(kConvThreads), pl.smem_bytes,
                  (cudaStream_t)stream, dim3(pl.pair ? 2u : 1u, 1, (unsigned)pl.splits), tmA1, tmA2, tmB, tmR, p);
}
