// Experiment (not product code): does a tcgen05 K-major SWIZZLE_128B operand descriptor whose start address is 128-byte
// aligned but NOT 1024-byte aligned read the rows a TMA 128B-swizzled load wrote? This decides how the halo-tile 3x3
// convolution addresses its nine taps (one A tile per 64-channel chunk, tap (ky,kx) = row offset ky*pitch + kx).
// Variants: descriptor base_offset field (bits 49..51) = 0, or = (start_address >> 7) & 7 (the PTX ISA formula).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_rowshift tools/exp_rowshift.cu -lcuda && ./exp_rowshift
#include "../stablediffusioneo_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

using namespace sdeo;

constexpr int kRows = 384;   // A rows resident in shared memory (two TMA boxes of 192 rows)
constexpr int kN = 64;
constexpr int kShifts = 48;

__global__ void __launch_bounds__(128, 1)
rowshift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  uint8_t* a_tile = smem + 1024;
  uint8_t* b_tile = a_tile + kRows * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_load, (kRows + kN) * 128);
    tma_load_2d(a_tile, &tmA, bar_load, 0, 0);
    tma_load_2d(a_tile + 192 * 128, &tmA, bar_load, 0, 192);
    tma_load_2d(b_tile, &tmB, bar_load, 0, 0);
  }
  mbar_wait(bar_load, 0);
  tc_fence_after();
  const uint32_t idesc = umma_idesc_bf16(128, kN);
  uint32_t ph = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int sh = 0; sh < kShifts; ++sh) {
      if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)sh * 128u;
        uint64_t a_desc = umma_desc_k_sw128(a_addr);
        if (variant == 1) a_desc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(b_tile));
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k > 0);
        tc_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph);
      ph ^= 1u;
      tc_fence_after();
      uint32_t r[32];
      float* dst = out + ((size_t)(variant * kShifts + sh) * 128 + warp * 32 + lane) * kN;
      for (int c = 0; c < kN; c += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) dst[c + j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static CUtensorMap make_map(void* base, int rows, int box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  CUtensorMap m;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  std::vector<__nv_bfloat16> hA((size_t)kRows * 64), hB((size_t)kN * 64);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) * 0.125f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) * 0.25f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  const size_t out_n = (size_t)2 * kShifts * 128 * kN;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, out_n * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, out_n * 4);
  CUtensorMap tmA = make_map(dA, kRows, 192), tmB = make_map(dB, kN, kN);
  const size_t smem = 2048 + (size_t)(kRows + kN) * 128 + 1024;
  cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rowshift_kernel<<<1, 128, smem>>>(tmA, tmB, dOut);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hOut(out_n);
  cudaMemcpy(hOut.data(), dOut, out_n * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant) {
    int bad_shifts = 0;
    printf("variant %d (base_offset %s):", variant, variant ? "= (addr>>7)&7" : "= 0");
    for (int sh = 0; sh < kShifts; ++sh) {
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)fA[(size_t)(sh + m) * 64 + k] * fB[(size_t)n * 64 + k];
          const double d = fabs(ref - hOut[((size_t)(variant * kShifts + sh) * 128 + m) * kN + n]);
          if (d > maxerr) maxerr = d;
        }
      printf(" %d:%s", sh, maxerr < 1e-3 ? "ok" : "BAD");
      if (maxerr >= 1e-3) ++bad_shifts;
    }
    printf("\n  -> %d of %d shifts wrong\n", bad_shifts, kShifts);
  }
  return 0;
}
