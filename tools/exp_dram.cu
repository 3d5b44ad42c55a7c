// Experiment (not product code): does the weight matrix's memory layout limit the cold (HBM) weight stream of the
// small-M layers? 128 CTAs each pull 23 tiles of 160 rows x 64 bf16 (20 KB) through a TMA ring, L2 flushed before:
//   layout 0: [cout][K] row-major (row pitch 23 KB): a tile is 160 separate 128-byte segments (the round-1 packing)
//   layout 1: [cout/16][K/64][16][64]: a tile is 10 contiguous 2 KB blocks
//   layout 2: tile-contiguous (20 KB contiguous per tile)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_dram tools/exp_dram.cu -lcuda && ./exp_dram
#include "../stablediffusioneo_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <vector>

using namespace sdeo;

constexpr int kCout = 1280, kK = 11520, kBN = 160, kSplits = 8, kStages = 6;

__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int layout, int steps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 16;
  uint8_t* tiles = smem + 1024;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int n_tile = blockIdx.x % (kCout / kBN), split = blockIdx.x / (kCout / kBN);
  const int k0 = split * steps;
  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < steps; ++i) {
        if (i >= kStages) mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], kBN * 128);
        uint8_t* dst = tiles + s * kBN * 128;
        if (layout == 0) tma_load_2d(dst, &tm, &full_bar[s], (k0 + i) * 64, n_tile * kBN);
        else tma_load_4d(dst, &tm, &full_bar[s], 0, 0, k0 + i, n_tile * (kBN / 16));
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < steps; ++i) {
        mbar_wait(&full_bar[s], ph);
        mbar_arrive(&empty_bar[s]);
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  }
  __syncthreads();
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  const size_t bytes = (size_t)kCout * kK * 2;
  uint8_t *dW, *dFlush;
  cudaMalloc(&dW, bytes);
  cudaMemset(dW, 1, bytes);
  cudaMalloc(&dFlush, (size_t)256 << 20);
  const size_t smem = 2048 + (size_t)kStages * kBN * 128;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int steps = kK / 64 / kSplits;  // 22 (the last half chunk is dropped)
  for (int layout = 0; layout < 3; ++layout) {
    CUtensorMap tm;
    CUresult r;
    if (layout == 0) {
      cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)kCout};
      cuuint64_t strides[1] = {(cuuint64_t)kK * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)kBN};
      cuuint32_t es[2] = {1, 1};
      r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      // [row block][k chunk][16][64] (layout 1) or [n tile][k chunk][10 blocks][16][64] = tile-contiguous (layout 2)
      const cuuint64_t kc = kK / 64;
      cuuint64_t dims[4] = {64, 16, kc, (cuuint64_t)kCout / 16};
      cuuint64_t strides[3] = {128, 2048, kc * 2048};
      if (layout == 2) { strides[1] = 2048 * (kBN / 16); strides[2] = 2048; }   // chunk stride 20 KB, row-block stride 2 KB
      if (layout == 2) { dims[2] = kc; }
      cuuint32_t box[4] = {64, 16, 1, (cuuint32_t)kBN / 16};
      cuuint32_t es[4] = {1, 1, 1, 1};
      r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dW, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d (layout %d)\n", (int)r, layout); continue; }
    for (int grid : {64, 128}) {
      float best = 1e9f;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int splits = grid / (kCout / kBN);
      const int st = kK / 64 / splits;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(dFlush, rep, (size_t)256 << 20);
        cudaEventRecord(e0);
        stream_kernel<<<grid, 128, smem>>>(tm, layout, st);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double mb = (double)grid * st * kBN * 128 / 1e6;
      printf("layout %d grid %3d: %6.1f us for %5.1f MB cold = %5.2f TB/s\n", layout, grid, best * 1e3, mb, mb / (best * 1e3) / 1e3 * 1e0);
    }
  }
  return 0;
}
