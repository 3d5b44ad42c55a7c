// Experiment (not product code): what bounds the conv mainloop on one SM -- the tensor pipe, the shared-memory port (TMA
// writes + MMA operand reads), the per-SM L2->SM ingest rate or the chip-wide L2 rate? One CTA per SM runs the real
// pipeline shape (TMA producer thread, MMA issuer thread, ring of stages) with either side switched off:
//   tma=1 mma=0 : loads only (the consumer releases a stage as soon as it has landed)
//   tma=0 mma=1 : MMAs only, on whatever is in shared memory (ring of distinct stage addresses)
//   tma=1 mma=1 : the mainloop
// A descriptor start can be shifted by `shift` rows of 128 bytes (the halo tiling reads its taps that way).
// PAIR: cta_group::2, M = 256 per pair, every CTA stages 128 rows of A and N/2 rows of B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_pipe tools/exp_pipe.cu -lcuda && ./exp_pipe
#include "../stablediffusioneo_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

using namespace sdeo;

struct P {
  int tma, mma, N, nsteps, stages, a_rows, shift, same_b, kc, a_row_tiles;
  int lps;      // TMA loads per step: 1 = B only, 2 = A + B, 3 = A only, 4 = A and B each in two halves
  int plain_arrive;  // consumer releases stages with mbarrier.arrive instead of tcgen05.commit (tma=1 mma=0 only)
  int nacc;     // accumulators the K steps alternate between
  int wm;       // wait flavour: 0 = try_wait (library default), 1 = test_wait spin, 2 = try_wait with suspend hint 0, 3 = hint 1 ms
  long long* out;
};


__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity, int mode) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok = 0;
  if (mode == 0) { mbar_wait(bar, parity); return; }
  while (!ok) {
    if (mode == 1)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(a), "r"(parity), "r"(mode == 2 ? 0u : 1000000u) : "memory");
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(128, 1)
pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* done_bar = full_bar + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 34);
  uint8_t* tiles = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t px = PAIR ? (cluster_ctarank() & 1u) : 0u;
  const bool leader = px == 0;
  const int bn_cta = PAIR ? p.N / 2 : p.N;
  const int a_bytes = 128 * 128;  // the stage always reserves a full A tile; a_rows < 128 loads fewer rows into it
  const int b_bytes = bn_cta * 128;
  const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t0 = clock64();
  const uint16_t pair_mask = 3;
  auto full_addr = [&](uint64_t* bar) -> uint32_t { return PAIR ? mapa_u32(smem_u32(bar), 0) : smem_u32(bar); };
  if (warp == 0 && p.tma) {
    if (lane == 0) {
    const uint32_t tx = (uint32_t)((p.lps != 1 ? p.a_rows * 128 : 0) + (p.lps != 3 ? b_bytes : 0)) * (PAIR ? 2u : 1u);
    int s = 0; uint32_t ph = 0;
    long long pw = 0, pe = 0, pt = 0;
    const int arow0 = ((int)blockIdx.x % p.a_row_tiles) * 128;
    const int brow0 = (p.same_b ? 0 : ((int)blockIdx.x / (PAIR ? 2 : 1)) % 4 * p.N) + (int)px * bn_cta;
    for (int i = 0; i < p.nsteps; ++i) {
      const long long c0 = clock64();
      if (i >= p.stages) wait_mode(&empty_bar[s], ph ^ 1u, p.wm);
      const long long c1 = clock64();
      if (leader) mbar_expect_tx(&full_bar[s], tx);
      const long long c2 = clock64();
      uint8_t* dst = tiles + (size_t)s * stage_bytes;
      const int kc = (i % p.kc) * 64;
      if (PAIR) {
        tma_load_2d_pair(dst, &tmA, full_addr(&full_bar[s]), kc, arow0);
        tma_load_2d_pair(dst + a_bytes, &tmB, full_addr(&full_bar[s]), kc, brow0);
      } else if (p.lps == 4) {
        tma_load_2d(dst, &tmA, &full_bar[s], kc, arow0);
        tma_load_2d(dst + p.a_rows * 64, &tmA, &full_bar[s], kc, arow0 + p.a_rows / 2);
        tma_load_2d(dst + a_bytes, &tmB, &full_bar[s], kc, brow0);
        tma_load_2d(dst + a_bytes + b_bytes / 2, &tmB, &full_bar[s], kc, brow0 + bn_cta / 2);
      } else {
        if (p.lps != 1) tma_load_2d(dst, &tmA, &full_bar[s], kc, arow0);
        if (p.lps != 3) tma_load_2d(dst + a_bytes, &tmB, &full_bar[s], kc, brow0);
      }
      const long long c3 = clock64();
      pw += c1 - c0; pe += c2 - c1; pt += c3 - c2;
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (p.out) { p.out[1024 + blockIdx.x * 8 + 0] = pw; p.out[1024 + blockIdx.x * 8 + 1] = pe; p.out[1024 + blockIdx.x * 8 + 2] = pt; }
    }
    __syncwarp();   // lanes 1..31 park here (a spinning try_wait in the producer's own warp would steal its issue slots)
  } else if (warp == 1 && leader) {
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, (uint32_t)p.N);
    int s = 0; uint32_t ph = 0;
    long long cw = 0, cr = 0;
    for (int i = 0; i < p.nsteps; ++i) {
      const long long c0 = clock64();
      if (p.tma) { wait_mode(&full_bar[s], ph, p.wm); tc_fence_after(); }
      const long long c1 = clock64();
      if (elect_one()) {
        const uint32_t st = smem_u32(tiles) + (uint32_t)(s * stage_bytes);
        if (p.mma) {
          const uint64_t a_desc = umma_desc_k_sw128(st + (uint32_t)p.shift * 128u);
          const uint64_t b_desc = umma_desc_k_sw128(st + a_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = tmem_base + (uint32_t)((p.nacc == 2 ? (i & 1) : p.nacc == 4 ? k : 0) * p.N);
            const uint32_t acc = (p.nacc == 2 ? (i > 1 || k) : p.nacc == 4 ? (i > 0) : (i | k)) ? 1u : 0u;
            if (PAIR) tc_mma_bf16_pair(d, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            else tc_mma_bf16(d, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
          }
        }
        if (p.tma) {
          if (PAIR) tc_commit_pair(&empty_bar[s], pair_mask);
          else if (p.plain_arrive) mbar_arrive(&empty_bar[s]);
          else tc_commit(&empty_bar[s]);
        }
      }
      __syncwarp();
      const long long c2 = clock64();
      cw += c1 - c0; cr += c2 - c1;
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (lane == 0) { p.out[1024 + blockIdx.x * 8 + 3] = cw; p.out[1024 + blockIdx.x * 8 + 4] = cr; }
    if (elect_one()) {
      if (PAIR) tc_commit_pair(done_bar, pair_mask);
      else tc_commit(done_bar);
    }
    __syncwarp();
  }
  mbar_wait(done_bar, 0);
  const long long t1 = clock64();
  if (threadIdx.x == 0) p.out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;
static CUtensorMap make_map(void* base, int rows, int kcols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kcols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  const int kc = 16, a_row_tiles = 148, rowsA = a_row_tiles * 128, rowsB = 1024, kcols = kc * 64;
  __nv_bfloat16 *dA, *dB;
  long long* dOut;
  cudaMalloc(&dA, (size_t)rowsA * kcols * 2);
  cudaMalloc(&dB, (size_t)rowsB * kcols * 2);
  cudaMalloc(&dOut, 1024 * 8 * 9);
  {
    std::vector<__nv_bfloat16> h((size_t)rowsA * kcols);
    srand(3);
    for (auto& v : h) v = __float2bfloat16((rand() % 17 - 8) * 0.125f);
    cudaMemcpy(dA, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, h.data(), (size_t)rowsB * kcols * 2, cudaMemcpyHostToDevice);
  }
  const size_t smem_max = 227 * 1024;
  cudaFuncSetAttribute(pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
  cudaFuncSetAttribute(pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
  struct Case { int pair, tma, mma, N, stages, a_rows, shift, grid, same_b, lps, plain, nacc, wm; };
  std::vector<Case> cases;
  for (int wm : {0, 1}) {
    cases.push_back({0, 1, 0, 128, 5, 128, 0, 148, 1, 1, 1, 1, wm});
    cases.push_back({0, 1, 0, 128, 5, 128, 0, 148, 1, 2, 1, 1, wm});
    cases.push_back({0, 1, 0, 256, 4, 128, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({0, 1, 0, 128, 2, 128, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({0, 1, 1, 160, 5, 128, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({0, 1, 1, 256, 4, 128, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({0, 1, 1, 256, 4, 128, 0, 48, 1, 2, 0, 1, wm});
    cases.push_back({0, 1, 1, 160, 5, 16, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({1, 1, 1, 256, 6, 128, 0, 148, 1, 2, 0, 1, wm});
    cases.push_back({1, 1, 1, 256, 6, 16, 0, 148, 1, 2, 0, 1, wm});
  }
  printf("pair tma mma   N stages a_rows shift grid same_b lps plain nacc wm | cycles/step (median, max) | TMA B/clk/SM | MMA floor cyc/step\n");
  for (const Case& c : cases) {
    P p;
    p.tma = c.tma; p.mma = c.mma; p.N = c.N; p.nsteps = 96; p.stages = c.stages; p.a_rows = c.a_rows; p.shift = c.shift;
    p.same_b = c.same_b; p.lps = c.lps; p.plain_arrive = c.plain; p.nacc = c.nacc; p.wm = c.wm; p.kc = kc; p.a_row_tiles = a_row_tiles; p.out = dOut;
    const int bn_cta = c.pair ? c.N / 2 : c.N;
    const int stage_bytes = 128 * 128 + ((bn_cta * 128 + 1023) & ~1023);
    const size_t smem = 2048 + (size_t)c.stages * stage_bytes + 128 * 128 /* shifted reads run past the last stage */;
    if (smem > smem_max) { printf("skip (smem)\n"); continue; }
    CUtensorMap tmA = make_map(dA, rowsA, kcols, c.a_rows), tmB = make_map(dB, rowsB, kcols, bn_cta);
    cudaMemset(dOut, 0, 1024 * 8 * 9);
    for (int rep = 0; rep < 3; ++rep) {
      if (c.pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c.grid & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, pipe_kernel<true>, tmA, tmB, p);
      } else {
        pipe_kernel<false><<<c.grid, 128, smem>>>(tmA, tmB, p);
      }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(c.grid);
    cudaMemcpy(h.data(), dOut, c.grid * 8, cudaMemcpyDeviceToHost);
    std::vector<long long> v;
    for (int i = 0; i < c.grid; ++i) if (h[i] > 0) v.push_back(h[i]);
    std::sort(v.begin(), v.end());
    const double med = v.empty() ? 0 : (double)v[v.size() / 2] / p.nsteps, mx = v.empty() ? 0 : (double)v.back() / p.nsteps;
    const double bytes = c.tma ? (double)((c.lps != 1 ? c.a_rows * 128 : 0) + (c.lps != 3 ? bn_cta * 128 : 0)) : 0;
    const double floor_ = (double)(c.pair ? 256 : 128) * c.N / (256.0 * (c.pair ? 2 : 1)) * 4;
    {
      std::vector<long long> d(8);
      cudaMemcpy(d.data(), dOut + 1024, 64, cudaMemcpyDeviceToHost);
      printf("   CTA 0 per step: producer wait %.0f expect %.0f tma %.0f | consumer wait %.0f rest %.0f\n", (double)d[0] / p.nsteps,
             (double)d[1] / p.nsteps, (double)d[2] / p.nsteps, (double)d[3] / p.nsteps, (double)d[4] / p.nsteps);
    }
    printf("%4d %3d %3d %4d %6d %6d %5d %4d %6d %3d %5d %4d %2d | %8.1f %8.1f | %6.1f | %6.0f\n", c.pair, c.tma, c.mma, c.N, c.stages, c.a_rows,
           c.shift, c.grid, c.same_b, c.lps, c.plain, c.nacc, c.wm, med, mx, bytes / med, floor_);
  }
  return 0;
}
