// Flash-style attention for sm_100a: S = Q K^T and O += P V on tcgen05 tensor cores with both accumulators in
// TMEM, online softmax in registers (one thread per query row), P staged through 128B-swizzled shared memory.
//
// Layouts (produced by the QKV epilogue of gemm_conv.cu): q,k [B*heads, n, d] bf16; vt [B*heads, d, ldv] bf16
// (V transposed, so every MMA operand is K-major); o [B, nq, heads*d] bf16.
// One CTA per (128-query tile, batch*head). Warp 0: TMA producer. Warp 1: TMEM allocator + MMA issuer.
// Warps 2..9: softmax / O-rescale / epilogue, two threads per query row (64 of the 128 S columns each).
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"

namespace sdeo {

constexpr int kAttThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two threads per query row)
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

struct AttParams {
  int nq, nkv, d, heads;
  int causal;     // 1: query i attends keys 0..i only (CLIP text encoder)
  int nkc;        // ceil(d / 64): 64-wide K chunks of the QK^T contraction
  int dk16;       // round_up(d, 16): contraction length actually multiplied
  int dv16;       // round_up(d + 1, 16): N of the PV MMA -- V^T row d is all ones, so O[:, d] accumulates the row sums of P
  int kv_stages;  // 1 or 2
  int tmem_cols;
  float scale_log2;
  __nv_bfloat16* o;
};

__global__ void __launch_bounds__(kAttThreads, 2)  // two CTAs per SM (d = 40: 106 KB of shared memory each): 192 tiles fit one wave
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int q_bytes = p.nkc * kTileQ * 128;
  const int k_bytes = p.nkc * kTileKV * 128;
  const int v_chunk_bytes = p.dv16 * 128;     // [dv16 rows][64 kv] bf16
  const int v_bytes = 2 * v_chunk_bytes;
  const int kv_stage_bytes = k_bytes + v_bytes;
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + q_bytes;
  uint8_t* sP = sKV + (size_t)p.kv_stages * kv_stage_bytes;  // 2 chunks [128][64] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileQ * 128);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_done = bars + 7;
  uint64_t* s_read = bars + 9;        // the softmax warps hold S(j) in registers: the MMA warp may overwrite the S columns
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);   // (bars + 9: s_read)
  float* s_xchg = reinterpret_cast<float*>(bars + 16);  // [2][128] partner exchange (row max / row sum)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int n_kv_tiles = (p.nkv + kTileKV - 1) / kTileKV;
  const int trc = trace_start(2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 1);
    mbar_init(o_done, 1);
    mbar_init(s_read, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base;          // 128 columns
  const uint32_t tmem_O = tmem_base + 128;    // dv16 columns
  griddep_launch_dependents();  // PDL (after this CTA's own TMEM allocation, see gemm_conv.cu)
  griddep_wait();  trace_mark(trc, 2);  // q/k/v are written by the preceding projection kernels; o may still be read by an earlier one

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, (uint32_t)q_bytes);
      for (int c = 0; c < p.nkc; ++c) tma_load_3d(sQ + c * kTileQ * 128, &tmQ, q_full, c * 64, q_tile * kTileQ, bh);
      for (int j = 0; j < n_kv_tiles; ++j) {
        const int s = j % p.kv_stages;
        const uint32_t ph = (uint32_t)(j / p.kv_stages) & 1u;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_expect_tx(&kv_full[s], (uint32_t)kv_stage_bytes);
        uint8_t* kdst = sKV + (size_t)s * kv_stage_bytes;
        uint8_t* vdst = kdst + k_bytes;
        for (int c = 0; c < p.nkc; ++c) tma_load_3d(kdst + c * kTileKV * 128, &tmK, &kv_full[s], c * 64, j * kTileKV, bh);
        for (int c = 0; c < 2; ++c) tma_load_3d(vdst + c * v_chunk_bytes, &tmV, &kv_full[s], j * kTileKV + c * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      const uint32_t idesc_pv = umma_idesc_bf16(128, (uint32_t)p.dv16);
      // S(j) = Q K_j^T : M=128 (queries), N=kv16, K=dk16
      auto issue_s = [&](int j) {
        const int s = j % p.kv_stages;
        const uint32_t ph = (uint32_t)(j / p.kv_stages) & 1u;
        const int kv_cols = min(kTileKV, p.nkv - j * kTileKV);
        const int kv16 = (kv_cols + 15) & ~15;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + (size_t)s * kv_stage_bytes);
        const uint32_t idesc_s = umma_idesc_bf16(128, (uint32_t)kv16);
        for (int kk = 0; kk < p.dk16 / 16; ++kk) {
          const int c = kk >> 2, ki = kk & 3;
          const uint64_t a_desc = umma_desc_k_sw128(q_addr + (uint32_t)(c * kTileQ * 128)) + (uint64_t)(2 * ki);
          const uint64_t b_desc = umma_desc_k_sw128(k_addr + (uint32_t)(c * kTileKV * 128)) + (uint64_t)(2 * ki);
          tc_mma_bf16(tmem_S, a_desc, b_desc, idesc_s, kk > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
      };
      issue_s(0);
      for (int j = 0; j < n_kv_tiles; ++j) {
        const int s = j % p.kv_stages;
        const int kv_cols = min(kTileKV, p.nkv - j * kTileKV);
        const int kv16 = (kv_cols + 15) & ~15;
        const uint32_t v_addr = smem_u32(sKV + (size_t)s * kv_stage_bytes) + (uint32_t)k_bytes;
        // the softmax warps have read S(j) into registers and written P(j) once p_full(j) completes: S(j+1) may
        // overwrite the S columns now and runs on the tensor pipe ahead of PV(j), so that the next tile's row
        // maxima overlap this tile's PV (needs the second K/V stage; with one stage K_{j+1} waits for PV(j))
        // S(j+1) is issued as soon as the softmax warps have S(j) in REGISTERS (s_read), not when P(j) is written: the
        // QK^T of the next tile then runs under this tile's exponentials instead of between two softmax passes
        if (j + 1 < n_kv_tiles && p.kv_stages > 1) {
          mbar_wait(s_read, (uint32_t)j & 1u);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
        // ---- O += P V : M=128, N=dv16, K=kv16 ----
        for (int kk = 0; kk < kv16 / 16; ++kk) {
          const int c = kk >> 2, ki = kk & 3;
          const uint64_t a_desc = umma_desc_k_sw128(p_addr + (uint32_t)(c * kTileQ * 128)) + (uint64_t)(2 * ki);
          const uint64_t b_desc = umma_desc_k_sw128(v_addr + (uint32_t)(c * v_chunk_bytes)) + (uint64_t)(2 * ki);
          tc_mma_bf16(tmem_O, a_desc, b_desc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[s]);
        tc_commit(o_done);
        if (j + 1 < n_kv_tiles && p.kv_stages == 1) issue_s(j + 1);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: TWO threads per query row =====================
    // warps 2..5 take columns [0, 64) of the 128-wide S tile, warps 6..9 columns [64, 128); the partners of a row
    // exchange their partial row maximum (per tile) and row sum (once, at the end) through shared memory.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int st = threadIdx.x - 64;  // 0..255
    float m_run = -INFINITY;  // running max of s * scale_log2 (identical in both partners)
    uint8_t* p_row = sP + half * (kTileQ * 128) + row * 128;  // this thread's 64 columns = one 128-byte chunk row
    const int n_oc = p.dv16 / 16;                 // 16-column chunks of O
    const int oc_begin = half == 0 ? 0 : (n_oc + 1) / 2, oc_end = half == 0 ? (n_oc + 1) / 2 : n_oc;

    for (int j = 0; j < n_kv_tiles; ++j) {
      const int kv_cols = min(kTileKV, p.nkv - j * kTileKV);
      int kv_lim = kv_cols;  // valid key columns of this tile for this query row
      if (p.causal) kv_lim = min(kv_cols, max(0, q_tile * kTileQ + row + 1 - j * kTileKV));
      const int ncols = min(64, max(0, kv_lim - half * 64));  // valid columns among this thread's 64
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      if (ncols > 0) tmem_ld32(tmem_S + lane_off + (uint32_t)(half * 64), r0);
      if (ncols > 32) tmem_ld32(tmem_S + lane_off + (uint32_t)(half * 64 + 32), r1);
      tmem_ld_wait();
      float m_loc = -INFINITY;
      const bool full_tile = __all_sync(0xffffffffu, ncols == 64);  // (all but the last tile; never under the causal mask's edge)
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) m_loc = fmax3(m_loc, __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
#pragma unroll
        for (int i = 0; i < 32; i += 2) m_loc = fmax3(m_loc, __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < ncols) m_loc = fmaxf(m_loc, __uint_as_float(r0[i]));
          if (32 + i < ncols) m_loc = fmaxf(m_loc, __uint_as_float(r1[i]));
        }
      }
      // V^T row d of this tile := 1.0 (both 64-key chunks): the PV MMA then accumulates sum_k P[q, k] in O[:, d]. (V(j) has
      // landed: S(j) needed the same barrier; PV(j) starts only after p_full(j), which follows the fence + barrier below.)
      if (st < 16) {
        uint8_t* vrow = sKV + (size_t)(j % p.kv_stages) * kv_stage_bytes + k_bytes + (st >> 3) * v_chunk_bytes + p.d * 128 + (st & 7) * 16;
        *reinterpret_cast<uint4*>(vrow) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      }
      s_xchg[half * kTileQ + row] = m_loc;
      tc_fence_before();   // (this thread's TMEM loads of S(j) are complete: ordered before the MMA warp's next write)
      bar_sync(2, 256);
      if (st == 0) mbar_arrive(s_read);
      const float m_tile = fmaxf(m_loc, s_xchg[(half ^ 1) * kTileQ + row]);
      const float m_new = fmaxf(m_run, m_tile * p.scale_log2);
      const float alpha = fast_exp2(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      // previous PV must have finished reading P (and updating O) before either is overwritten
      if (j > 0) {
        mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      if (full_tile) {  // no column masks
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int cidx = u * 8 + i;
            const float sv = __uint_as_float(cidx < 32 ? r0[cidx & 31] : r1[cidx & 31]);
            pv[i] = fast_exp2(fmaf(sv, p.scale_log2, -m_new));
          }
          uint4 v;
          v.x = pack_bf16x2(pv[0], pv[1]); v.y = pack_bf16x2(pv[2], pv[3]);
          v.z = pack_bf16x2(pv[4], pv[5]); v.w = pack_bf16x2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(p_row + ((u ^ (row & 7)) << 4)) = v;
        }
      } else if (half * 64 < ((kv_cols + 15) & ~15)) {  // the PV contraction reads this chunk
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int cidx = u * 8 + i;
            const float sv = __uint_as_float(cidx < 32 ? r0[cidx & 31] : r1[cidx & 31]);
            const float e = fast_exp2(sv * p.scale_log2 - m_new);
            pv[i] = (cidx < ncols) ? e : 0.f;
          }
          uint4 v;
          v.x = pack_bf16x2(pv[0], pv[1]); v.y = pack_bf16x2(pv[2], pv[3]);
          v.z = pack_bf16x2(pv[4], pv[5]); v.w = pack_bf16x2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(p_row + ((u ^ (row & 7)) << 4)) = v;
        }
      }
      m_run = m_new;
      // rescale this thread's share of the O columns (skipped when no row of the warp moved its max)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        for (int c = oc_begin; c < oc_end; ++c) {
          uint32_t r[16];
          tmem_ld16(tmem_O + lane_off + (uint32_t)(c * 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st16(tmem_O + lane_off + (uint32_t)(c * 16), r);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      bar_sync(1, 256);
      if (st == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 [B, nq, heads*d] ----
    mbar_wait(o_done, (uint32_t)(n_kv_tiles - 1) & 1u);
    tc_fence_after();
    float inv_l;
    {  // the row sum sits in O column d
      uint32_t r[16];
      tmem_ld16(tmem_O + lane_off + (uint32_t)((p.d >> 4) * 16), r);
      tmem_ld_wait();
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.d & 15)) l = __uint_as_float(r[i]);
      inv_l = 1.0f / l;
    }
    const int qi = q_tile * kTileQ + row;
    const int b = bh / p.heads, head = bh % p.heads;
    __nv_bfloat16* orow = p.o + ((size_t)b * p.nq + qi) * ((size_t)p.heads * p.d) + (size_t)head * p.d;
    for (int c = oc_begin; c < oc_end; ++c) {
      uint32_t r[16];
      tmem_ld16(tmem_O + lane_off + (uint32_t)(c * 16), r);
      tmem_ld_wait();
      if (qi < p.nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = c * 16 + g * 8;
          if (col < p.d) {  // d % 8 == 0
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + col) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  trace_mark(trc, 3);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_attention)

static int attention_impl(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads, int32_t nq,
                          int32_t nkv, int32_t d, int32_t ldv, float scale, int causal, void* stream);

extern "C" int sdeo_attention(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads,
                              int32_t nq, int32_t nkv, int32_t d, int32_t ldv, float scale, void* stream) {
  return attention_impl(q, k, vt, o, batch, heads, nq, nkv, d, ldv, scale, 0, stream);
}

extern "C" int sdeo_attention_causal(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads,
                                     int32_t n, int32_t d, int32_t ldv, float scale, void* stream) {
  return attention_impl(q, k, vt, o, batch, heads, n, n, d, ldv, scale, 1, stream);
}

static int attention_impl(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads, int32_t nq,
                          int32_t nkv, int32_t d, int32_t ldv, float scale, int causal, void* stream) {
  if (!q || !k || !vt || !o) return set_error(SDEO_EINVAL, "attention: null argument");
  if (batch <= 0 || heads <= 0 || nq <= 0 || nkv <= 0 || d <= 0 || d % 8 != 0 || d > 192 || ldv < nkv || ldv % 8 != 0)
    return set_error(SDEO_EINVAL, "attention: unsupported geometry (need d % 8 == 0, d <= 192, ldv % 8 == 0)");
  const int BH = batch * heads;
  if (BH > 65535) return set_error(SDEO_EINVAL, "attention: batch*heads too large");
  AttParams p;
  p.nq = nq; p.nkv = nkv; p.d = d; p.heads = heads;
  p.causal = causal;
  p.nkc = (d + 63) / 64;
  p.dk16 = (d + 15) & ~15;
  p.dv16 = (d + 1 + 15) & ~15;   // one spare V^T row (index d) for the row sums
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = (__nv_bfloat16*)o;
  const int q_bytes = p.nkc * kTileQ * 128;
  const int kv_stage = p.nkc * kTileKV * 128 + 2 * p.dv16 * 128;
  const int fixed = q_bytes + 2 * kTileQ * 128 + 1024 + 128 + 2 * kTileQ * 4;
  p.kv_stages = (fixed + 2 * kv_stage <= 220 * 1024 && nkv > kTileKV) ? 2 : 1;
  const size_t smem = (size_t)fixed + (size_t)p.kv_stages * kv_stage;
  int tc = 32;
  while (tc < 128 + p.dv16) tc *= 2;
  p.tmem_cols = tc;

  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)nq, (uint64_t)BH};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)nq * d * 2};
    uint32_t box[3] = {64, kTileQ, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap_bf16(&tmQ, q, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)nkv, (uint64_t)BH};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)nkv * d * 2};
    uint32_t box[3] = {64, kTileKV, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap_bf16(&tmK, k, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)nkv, (uint64_t)d, (uint64_t)BH};
    uint64_t strides[2] = {(uint64_t)ldv * 2, (uint64_t)d * ldv * 2};
    uint32_t box[3] = {64, (uint32_t)p.dv16, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap_bf16(&tmV, vt, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(SDEO_ECUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)((nq + kTileQ - 1) / kTileQ), (unsigned)BH);
  return launch_k("attention", attention_kernel, grid, dim3(kAttThreads), smem, (cudaStream_t)stream, dim3(1, 1, 1), tmQ, tmK,
                  tmV, p);
}
