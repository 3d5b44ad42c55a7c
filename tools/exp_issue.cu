// Experiment (not product code): what one thread pays to ISSUE the pipeline's instructions on sm_100a -- cp.async.bulk.tensor
// (UTMALDG), tcgen05.mma (UTCHMMA), tcgen05.commit, mbarrier.try_wait on a completed phase, mbarrier.arrive.expect_tx -- and
// whether tcgen05.mma issue from two warps (two accumulators) doubles the small-N MMA rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_issue tools/exp_issue.cu -lcuda && ./exp_issue
#include "../stablediffusioneo_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <vector>

using namespace sdeo;

constexpr int kRep = 32;

__global__ void __launch_bounds__(128, 1)
issue_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmL,
             const __grid_constant__ CUtensorMap tm3, long long* out, int N) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // 64 barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 64);
  uint8_t* tiles = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 64; ++s) mbar_init(&bars[s], s == 30 ? 1000000 : 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* o = out + blockIdx.x * 64;
  if (warp == 0) {
    if (lane == 0) {
      // T1: kRep small loads (16 rows x 128 B), one barrier
      tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmL); tma_prefetch_desc(&tm3);
      mbar_expect_tx(&bars[0], kRep * 16 * 128);
      long long t0 = clock64();
      for (int i = 0; i < kRep; ++i) tma_load_2d(tiles + (i % 8) * 2048, &tmS, &bars[0], (i % 16) * 64, 0);
      long long t1 = clock64();
      mbar_wait(&bars[0], 0);
      long long t2 = clock64();
      o[0] = t1 - t0; o[1] = t2 - t0;
      // T2: kRep loads of 128 rows x 128 B
      mbar_expect_tx(&bars[1], kRep * 128 * 128);
      t0 = clock64();
      for (int i = 0; i < kRep; ++i) tma_load_2d(tiles + (i % 8) * 16384, &tmL, &bars[1], (i % 16) * 64, 0);
      t1 = clock64();
      mbar_wait(&bars[1], 0);
      t2 = clock64();
      o[2] = t1 - t0; o[3] = t2 - t0;
      // T2b: kRep/2 loads of 128 rows x 2 chunks (3-D box): same bytes in half the instructions
      mbar_expect_tx(&bars[2], kRep * 128 * 128);
      t0 = clock64();
      for (int i = 0; i < kRep / 2; ++i) tma_load_3d(tiles + (i % 4) * 32768, &tm3, &bars[2], 0, 0, (i % 8) * 2);
      t1 = clock64();
      mbar_wait(&bars[2], 0);
      t2 = clock64();
      o[4] = t1 - t0; o[5] = t2 - t0;
      // T4: try_wait on completed phases, expect_tx
      t0 = clock64();
      for (int i = 0; i < kRep; ++i) mbar_wait(&bars[i & 1], 0);
      t1 = clock64();
      o[6] = t1 - t0;
      t0 = clock64();
      for (int i = 0; i < kRep; ++i) mbar_expect_tx(&bars[8 + i], 16);
      t1 = clock64();
      o[7] = t1 - t0;
      // single load latency (small / large)
      mbar_expect_tx(&bars[3], 16 * 128);
      t0 = clock64();
      tma_load_2d(tiles, &tmS, &bars[3], 0, 0);
      mbar_wait(&bars[3], 0);
      o[8] = clock64() - t0;
      mbar_expect_tx(&bars[4], 128 * 128);
      t0 = clock64();
      tma_load_2d(tiles, &tmL, &bars[4], 64, 0);
      mbar_wait(&bars[4], 0);
      o[9] = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  // T3: MMA issue, one warp then two warps (separate accumulators), 4 x kRep MMAs each
  const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)N);
  for (int nw = 1; nw <= 2; ++nw) {
    if (warp >= 1 && warp <= nw) {
      const uint64_t a0 = umma_desc_k_sw128(smem_u32(tiles) + (warp - 1) * 65536);
      const uint64_t b0 = umma_desc_k_sw128(smem_u32(tiles) + 16384 + (warp - 1) * 65536);
      const uint32_t d = tmem_base + (uint32_t)((warp - 1) * 256);
      uint64_t* done = &bars[40 + nw * 2 + warp];
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < kRep; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(d, a0 + 2 * k + (i & 1) * 2048, b0 + 2 * k + (i & 1) * 2048, idesc, (i | k) ? 1u : 0u);
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) tc_commit(done);
      __syncwarp();
      mbar_wait(done, 0);
      const long long t2 = clock64();
      if (lane == 0) { o[10 + (nw - 1) * 4 + (warp - 1) * 2] = t1 - t0; o[11 + (nw - 1) * 4 + (warp - 1) * 2] = t2 - t0; }
    }
    __syncthreads();
  }
  // T7: pipe-bound MMA rate with the A descriptor start shifted by `sh` rows of 128 bytes (the halo tiling's tap offsets)
  for (int v = 0; v < 4; ++v) {
    if (warp == 1) {
      const int sh = v == 0 ? 0 : (v == 1 ? 1 : (v == 2 ? 27 : 54));
      const uint64_t a0 = umma_desc_k_sw128(smem_u32(tiles) + sh * 128);
      const uint64_t b0 = umma_desc_k_sw128(smem_u32(tiles) + 32768);
      uint64_t* done = &bars[56 + v];
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < kRep; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, a0 + 2 * k + (i & 1) * 4096, b0 + 2 * k + (i & 1) * 4096, idesc, (i | k) ? 1u : 0u);
        }
      }
      __syncwarp();
      if (elect_one()) tc_commit(done);
      __syncwarp();
      mbar_wait(done, 0);
      const long long t2 = clock64();
      if (lane == 0) o[36 + v] = t2 - t0;
    }
    __syncthreads();
  }
  // T6: the consumer's K step = [try_wait on a completed phase] [tcgen05.fence::after_thread_sync] 4 MMAs [commit]
  //   variant 0: MMAs only, 1: + commit, 2: + fence, 3: + fence + try_wait, 4: try_wait + MMAs + commit (no fence)
  for (int variant = 0; variant < 5; ++variant) {
    if (warp == 1) {
      const uint64_t a0 = umma_desc_k_sw128(smem_u32(tiles));
      const uint64_t b0 = umma_desc_k_sw128(smem_u32(tiles) + 16384);
      uint64_t* done = &bars[50 + variant];
      const long long t0 = clock64();
      for (int i = 0; i < kRep; ++i) {
        if (variant >= 3) mbar_wait(&bars[0], 0);
        if (variant == 2 || variant == 3) tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, a0 + 2 * k + (i & 1) * 2048, b0 + 2 * k + (i & 1) * 2048, idesc, (i | k) ? 1u : 0u);
          if (variant >= 1) tc_commit(&bars[30]);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) tc_commit(done);
      __syncwarp();
      mbar_wait(done, 0);
      const long long t2 = clock64();
      if (lane == 0) { o[24 + variant * 2] = t1 - t0; o[25 + variant * 2] = t2 - t0; }
    }
    __syncthreads();
  }
  // T5: commit issue cost (nothing outstanding)
  if (warp == 1) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < kRep; ++i) tc_commit(&bars[24]);
    }
    __syncwarp();
    if (lane == 0) o[20] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;
static CUtensorMap make_map2(void* base, int rows, int kcols, int box_rows) {
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
// the same matrix seen as [k chunk][row][64]: one box = 64 x rows x 2 chunks
static CUtensorMap make_map3(void* base, int rows, int kcols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(kcols / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)kcols * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode3 failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  const int rows = 1024, kcols = 1024;
  __nv_bfloat16* dA;
  long long* dOut;
  cudaMalloc(&dA, (size_t)rows * kcols * 2);
  cudaMemset(dA, 0, (size_t)rows * kcols * 2);
  cudaMalloc(&dOut, 148 * 64 * 8);
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  CUtensorMap tmS = make_map2(dA, rows, kcols, 16), tmL = make_map2(dA, rows, kcols, 128), tm3 = make_map3(dA, rows, kcols, 128);
  for (int grid : {1, 148}) {
    for (int N : {64, 160, 256}) {
      cudaMemset(dOut, 0, 148 * 64 * 8);
      for (int rep = 0; rep < 2; ++rep) issue_kernel<<<grid, 128, smem>>>(tmS, tmL, tm3, dOut, N);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(64);
      cudaMemcpy(h.data(), dOut, 64 * 8, cudaMemcpyDeviceToHost);
      const double r = kRep;
      printf("grid %3d N %3d | per instruction: TMA 2KB issue %.0f (all landed %.0f) | TMA 16KB issue %.0f (landed %.0f) | TMA 3-D 32KB issue %.0f "
             "(landed %.0f) | try_wait(done) %.0f | expect_tx %.0f | commit %.0f | single load latency 2KB %lld 16KB %lld\n",
             grid, N, h[0] / r, h[1] / r, h[2] / r, h[3] / r, h[4] / (r / 2), h[5] / (r / 2), h[6] / r, h[7] / r, h[20] / r, h[8], h[9]);
      printf("             consumer step (4 MMAs): issue/done cycles per step: MMAs only %.0f/%.0f | +commit %.0f/%.0f | +commit+fence %.0f/%.0f | "
             "+commit+fence+try_wait %.0f/%.0f | try_wait+commit (no fence) %.0f/%.0f\n", h[24] / r, h[25] / r, h[26] / r, h[27] / r, h[28] / r,
             h[29] / r, h[30] / r, h[31] / r, h[32] / r, h[33] / r);
      printf("             pipe-bound MMA, A start shifted by 0 / 1 / 27 / 54 rows: %.0f %.0f %.0f %.0f cycles per MMA\n", h[36] / (4 * r),
             h[37] / (4 * r), h[38] / (4 * r), h[39] / (4 * r));
      printf("             MMA x%d: one warp issue %.0f /MMA, done %.0f /MMA | two warps: issue %.0f, %.0f done %.0f, %.0f /MMA (per warp)\n",
             4 * kRep, h[10] / (4 * r), h[11] / (4 * r), h[14] / (4 * r), h[16] / (4 * r), h[15] / (4 * r), h[17] / (4 * r));
    }
  }
  return 0;
}
