// fp32 ("precise") mode support kernels. The contractions of fp32 mode still run on the bf16 tcgen05 GEMM: every fp32
// operand is split into bf16 terms (v = hi + mid + lo, each term the bf16 rounding of what the previous ones left) and
// the product is rebuilt from the leading cross terms by CONCATENATING the terms along K:
//   x.w ~= hi.hi + lo.hi + hi.lo          activation [hi | mid | hi], weight [hi | hi | mid]   (K x 3, ~2^-17 per operand)
// (a 6-term pattern adds mid.mid, lo.hi, hi.lo for full fp32 operand precision). Accumulation is fp32 in TMEM as before.
// Everything that is not a contraction -- GroupNorm, LayerNorm, softmax/attention, GEGLU, SiLU, the sinusoidal
// embedding -- is computed here in plain fp32 on fp32 tensors (no bf16 tensor anywhere between two GEMMs).
// These kernels are for the 1e-4 parity mode, not for the bench path: simple, one pass per op.
#include "common.cuh"
#include "host_util.h"
#include "../../include/sdeo.h"

namespace sdeo {

__device__ __forceinline__ void precise_prologue() {
  const int trc = trace_start(5);
  griddep_launch_dependents();
  griddep_wait();
  trace_mark(trc, 2);
  trace_mark(trc, 3);
}

// ---- operand split -------------------------------------------------------------------------------
// x: fp32, element (row r, channel ch) at x[r * ldx + ch] (nchw_hw == 0) or, for an NCHW source with nchw_hw = H*W,
// at x[((r / hw) * c + ch) * hw + r % hw]. y: bf16 [rows, terms * cp], block t holds level ((pattern >> 2t) & 3) of
// every channel (0 = hi, 1 = mid, 2 = lo); channels c..cp-1 are zero.
__global__ void split_terms_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int c,
                                   int cp, long long ldx, int nchw_hw, int terms, unsigned pattern) {
  precise_prologue();
  const long long total = rows * cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cp;
    const int ch = (int)(i % cp);
    float v = 0.f;
    if (ch < c) v = nchw_hw ? x[((r / nchw_hw) * c + ch) * (long long)nchw_hw + r % nchw_hw] : x[r * ldx + ch];
    __nv_bfloat16 lv[3];
    lv[0] = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(lv[0]);
    lv[1] = __float2bfloat16(r1);
    lv[2] = __float2bfloat16(r1 - __bfloat162float(lv[1]));
    __nv_bfloat16* row = y + r * (long long)terms * cp;
    for (int t = 0; t < terms; ++t) row[(long long)t * cp + ch] = lv[(pattern >> (2 * t)) & 3u];
  }
}

// Weight side: w fp32 [cout, cin, kk] (PyTorch filter layout) -> fp32 [cout, terms * cp, kk] whose values are exactly
// bf16-representable (the regular weight packer then converts them without rounding). cin0..cin0+cin-1 selects a channel
// range of the source filter (the two halves of a fused concat are split separately); src_cin is its full channel count.
__global__ void split_terms_weight_kernel(const float* __restrict__ w, float* __restrict__ y, int cout, int src_cin,
                                          int cin0, int cin, int cp, int kk, int terms, unsigned pattern) {
  precise_prologue();
  const long long total = (long long)cout * cp * kk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk);
    const int ch = (int)((i / kk) % cp);
    const long long o = i / ((long long)kk * cp);
    float v = 0.f;
    if (ch < cin) v = w[(o * src_cin + cin0 + ch) * kk + tap];
    float lv[3];
    lv[0] = __bfloat162float(__float2bfloat16(v));
    const float r1 = v - lv[0];
    lv[1] = __bfloat162float(__float2bfloat16(r1));
    lv[2] = __bfloat162float(__float2bfloat16(r1 - lv[1]));
    for (int t = 0; t < terms; ++t) y[((o * terms + t) * cp + ch) * kk + tap] = lv[(pattern >> (2 * t)) & 3u];
  }
}

// ---- GroupNorm (+SiLU), fp32 in / fp32 out, optional fused concat ---------------------------------
// grid (groups, n); statistics in double (two plain passes over the group's hw x cpg elements).
__global__ void groupnorm_f32_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                     int hw, int c1, int c2, int groups, float eps, int with_silu) {
  precise_prologue();
  const int c = c1 + c2, cpg = c / groups;
  const int g = blockIdx.x, b = blockIdx.y;
  const long long count = (long long)hw * cpg;
  auto load = [&](long long e) -> float {
    const long long p = e / cpg;
    const int ch = g * cpg + (int)(e % cpg);
    return ch < c1 ? x1[((long long)b * hw + p) * c1 + ch] : x2[((long long)b * hw + p) * c2 + (ch - c1)];
  };
  double s = 0.0, ss = 0.0;
  for (long long e = threadIdx.x; e < count; e += blockDim.x) {
    const double v = load(e);
    s += v;
    ss += v * v;
  }
  __shared__ double red[2][32];
  __shared__ float stat[2];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tss = 0.0;
    for (int i = 0; i < nw; ++i) { ts += red[0][i]; tss += red[1][i]; }
    const double mean = ts / (double)count;
    double var = tss / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (long long e = threadIdx.x; e < count; e += blockDim.x) {
    const long long p = e / cpg;
    const int ch = g * cpg + (int)(e % cpg);
    float v = (load(e) - mean) * rstd * gamma[ch] + beta[ch];
    if (with_silu) v = v / (1.0f + expf(-v));
    y[((long long)b * hw + p) * c + ch] = v;
  }
}

// ---- LayerNorm, fp32 in / fp32 out: one warp per row, two passes ------------------------------------
__global__ void layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, int rows, int c, float eps) {
  precise_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * c;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += xr[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)c;
  float ss = 0.f;
  for (int i = lane; i < c; i += 32) {
    const float d = xr[i] - mean;
    ss += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = 1.0f / sqrtf(ss / (float)c + eps);
  float* yr = y + (long long)row * c;
  for (int i = lane; i < c; i += 32) yr[i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}

// ---- attention, fp32 CUDA-core: o = softmax(q k^T * scale) v ----------------------------------------
// q [B, nq, ldq] (head h at columns h*d..), k [B, nkv, ldk], v [B, nkv, ldv], o [B, nq, ldo]. grid (ceil(nq/16), heads, B),
// 128 threads: each warp owns 4 queries; keys are processed in tiles of 32 staged in shared memory (lane j scores key j,
// then every lane accumulates its output dims d_i = lane + 32 i with the probabilities broadcast by shuffles).
constexpr int kPQ = 16, kPK = 32, kPDmax = 160;
__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, float* __restrict__ o, int nq,
                                                            int nkv, int d, int ldq, int ldk, int ldv, int ldo, float scale) {
  precise_prologue();
  extern __shared__ float sm[];
  float* sq = sm;                      // [kPQ][d]
  float* sk = sq + kPQ * d;            // [kPK][d + 1]
  float* sv = sk + kPK * (d + 1);      // [kPK][d]
  const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kPQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kPQ * d; i += blockDim.x) {
    const int r = i / d, col = i % d;
    sq[i] = (q0 + r < nq) ? q[((long long)b * nq + q0 + r) * ldq + h * d + col] : 0.f;
  }
  float m[4], l[4], acc[4][kPDmax / 32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < kPDmax / 32; ++j) acc[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < nkv; k0 += kPK) {
    __syncthreads();
    for (int i = threadIdx.x; i < kPK * d; i += blockDim.x) {
      const int r = i / d, col = i % d;
      const bool ok = k0 + r < nkv;
      sk[r * (d + 1) + col] = ok ? k[((long long)b * nkv + k0 + r) * ldk + h * d + col] : 0.f;
      sv[r * d + col] = ok ? v[((long long)b * nkv + k0 + r) * ldv + h * d + col] : 0.f;
    }
    __syncthreads();
    const bool valid = k0 + lane < nkv;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* qr = sq + (warp * 4 + i) * d;
      const float* kr = sk + lane * (d + 1);
      float s = 0.f;
      for (int e = 0; e < d; ++e) s = fmaf(qr[e], kr[e], s);
      s = valid ? s * scale : -INFINITY;
      float mt = s;
      for (int of = 16; of > 0; of >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, of));
      const float mn = fmaxf(m[i], mt);
      const float p = valid ? expf(s - mn) : 0.f;
      const float corr = expf(m[i] - mn);  // exp(-inf) = 0 on the first tile
      float ps = p;
      for (int of = 16; of > 0; of >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, of);
      l[i] = l[i] * corr + ps;
      m[i] = mn;
#pragma unroll
      for (int j = 0; j < kPDmax / 32; ++j) acc[i][j] *= corr;
      for (int kk = 0; kk < kPK; ++kk) {
        const float pk = __shfl_sync(0xffffffffu, p, kk);
#pragma unroll
        for (int j = 0; j < kPDmax / 32; ++j) {
          const int col = lane + 32 * j;
          if (col < d) acc[i][j] = fmaf(pk, sv[kk * d + col], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + warp * 4 + i;
    if (qi >= nq) continue;
    const float inv = 1.0f / l[i];
#pragma unroll
    for (int j = 0; j < kPDmax / 32; ++j) {
      const int col = lane + 32 * j;
      if (col < d) o[((long long)b * nq + qi) * ldo + h * d + col] = acc[i][j] * inv;
    }
  }
}

// ---- elementwise fp32 ---------------------------------------------------------------------------------
// GEGLU: x [rows, 2*inner] -> y[r, j] = x[r, j] * gelu_erf(x[r, inner + j])   (attention.py:49-56)
__global__ void geglu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int inner) {
  precise_prologue();
  const long long total = rows * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / inner;
    const int j = (int)(i % inner);
    const float a = x[r * 2 * inner + j], g = x[r * 2 * inner + inner + j];
    y[i] = a * (0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)));
  }
}

__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long count) {
  precise_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v / (1.0f + expf(-v));
  }
}

// sinusoidal timestep embedding in fp32: y[i] = [cos(t_i f_j) | sin(t_i f_j)], f_j = exp(-ln(max_period) j / half)
__global__ void timestep_embedding_f32_kernel(const long long* __restrict__ t, float* __restrict__ y, int n, int dim,
                                              float max_period) {
  precise_prologue();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dim) return;
  const int b = i / dim, j = i % dim;
  if (j >= 2 * half) { y[i] = 0.f; return; }
  const int jj = j < half ? j : j - half;
  const float f = expf(-logf(max_period) * (float)jj / (float)half);
  const float a = (float)t[b] * f;
  y[i] = j < half ? cosf(a) : sinf(a);
}

static int grid_of(long long n, int block) {
  long long b = (n + block - 1) / block;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace sdeo

using namespace sdeo;
SDEO_DEFINE_TRACE_SETTER(sdeo_trace_set_precise)

extern "C" int sdeo_split_terms(const float* x, void* y, int64_t rows, int32_t c, int32_t cp, int64_t ldx, int32_t nchw_hw,
                                int32_t terms, uint32_t pattern, void* stream) {
  if (!x || !y || rows <= 0 || c <= 0 || cp < c || (cp % 8) != 0 || terms < 1 || terms > 8 || nchw_hw < 0 ||
      (!nchw_hw && ldx < c))
    return set_error(SDEO_EINVAL, "split_terms: bad args");
  return launch_k("split_terms", split_terms_kernel, dim3(grid_of(rows * cp, 256)), dim3(256), 0, (cudaStream_t)stream,
                  dim3(1, 1, 1), x, (__nv_bfloat16*)y, (long long)rows, c, cp, (long long)ldx, nchw_hw, terms, pattern);
}

extern "C" int sdeo_split_terms_weight(const float* w, float* y, int32_t cout, int32_t src_cin, int32_t cin0, int32_t cin,
                                       int32_t cp, int32_t kk, int32_t terms, uint32_t pattern, void* stream) {
  if (!w || !y || cout <= 0 || cin <= 0 || cin0 < 0 || cin0 + cin > src_cin || cp < cin || (cp % 8) != 0 || kk <= 0 ||
      terms < 1 || terms > 8)
    return set_error(SDEO_EINVAL, "split_terms_weight: bad args");
  return launch_k("split_terms_weight", split_terms_weight_kernel, dim3(grid_of((long long)cout * cp * kk, 256)), dim3(256),
                  0, (cudaStream_t)stream, dim3(1, 1, 1), w, y, cout, src_cin, cin0, cin, cp, kk, terms, pattern);
}

extern "C" int sdeo_groupnorm_f32(const float* x1, const float* x2, const float* gamma, const float* beta, float* y,
                                  int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps,
                                  int32_t with_silu, void* stream) {
  if (!x1 || !gamma || !beta || !y || n <= 0 || n > 65535 || hw <= 0 || c1 <= 0 || c2 < 0 || (c2 > 0 && !x2) || groups <= 0 ||
      ((c1 + c2) % groups) != 0)
    return set_error(SDEO_EINVAL, "groupnorm_f32: bad args");
  return launch_k("groupnorm_f32", groupnorm_f32_kernel, dim3(groups, n), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1),
                  x1, x2, gamma, beta, y, hw, c1, c2, groups, eps, with_silu);
}

extern "C" int sdeo_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int32_t rows, int32_t c,
                                  float eps, void* stream) {
  if (!x || !gamma || !beta || !y || rows <= 0 || c <= 0) return set_error(SDEO_EINVAL, "layernorm_f32: bad args");
  return launch_k("layernorm_f32", layernorm_f32_kernel, dim3((rows + 3) / 4), dim3(128), 0, (cudaStream_t)stream,
                  dim3(1, 1, 1), x, gamma, beta, y, rows, c, eps);
}

extern "C" int sdeo_attention_f32(const float* q, const float* k, const float* v, float* o, int32_t batch, int32_t heads,
                                  int32_t nq, int32_t nkv, int32_t d, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo,
                                  float scale, void* stream) {
  if (!q || !k || !v || !o || batch <= 0 || batch > 65535 || heads <= 0 || heads > 65535 || nq <= 0 || nkv <= 0 || d <= 0 ||
      d > kPDmax || ldq < heads * d || ldk < heads * d || ldv < heads * d || ldo < heads * d)
    return set_error(SDEO_EINVAL, "attention_f32: bad args (d <= 160)");
  const size_t smem = (size_t)(kPQ * d + kPK * (d + 1) + kPK * d) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return set_error(SDEO_ECUDA, "attention_f32: cannot raise the shared-memory limit");
    }
    smem_set = smem;
  }
  return launch_k("attention_f32", attention_f32_kernel, dim3((nq + kPQ - 1) / kPQ, heads, batch), dim3(128), smem,
                  (cudaStream_t)stream, dim3(1, 1, 1), q, k, v, o, nq, nkv, d, ldq, ldk, ldv, ldo, scale);
}

extern "C" int sdeo_geglu_f32(const float* x, float* y, int64_t rows, int32_t inner, void* stream) {
  if (!x || !y || rows <= 0 || inner <= 0) return set_error(SDEO_EINVAL, "geglu_f32: bad args");
  return launch_k("geglu_f32", geglu_f32_kernel, dim3(grid_of(rows * inner, 256)), dim3(256), 0, (cudaStream_t)stream,
                  dim3(1, 1, 1), x, y, (long long)rows, inner);
}

extern "C" int sdeo_silu_f32(const float* x, float* y, int64_t count, void* stream) {
  if (!x || !y || count <= 0) return set_error(SDEO_EINVAL, "silu_f32: bad args");
  return launch_k("silu_f32", silu_f32_kernel, dim3(grid_of(count, 256)), dim3(256), 0, (cudaStream_t)stream, dim3(1, 1, 1), x,
                  y, (long long)count);
}

extern "C" int sdeo_timestep_embedding_f32(const int64_t* t, float* y, int32_t n, int32_t dim, float max_period,
                                           void* stream) {
  if (!t || !y || n <= 0 || dim <= 0) return set_error(SDEO_EINVAL, "timestep_embedding_f32: bad args");
  return launch_k("timestep_embedding_f32", timestep_embedding_f32_kernel, dim3((n * dim + 127) / 128), dim3(128), 0,
                  (cudaStream_t)stream, dim3(1, 1, 1), (const long long*)t, y, n, dim, max_period);
}
