// Experiment (not product code): how fast can ONE SM write a tile to global memory (L2)? The conv epilogue writes
// 40..120 KB per CTA; measured it runs at 11..20 B/clk. Variants, each writing the same 96 KB from 384 threads:
//   0: st.global.v4 (16 B per lane, lanes contiguous)
//   1: st.global.v8.f32 (32 B per lane)
//   2: cp.async.bulk.global.shared::cta (1-D bulk stores of 2 KB rows, issued by one thread)
//   3: the same issued by 12 threads (one per warp)
//   4: cp.async.bulk.tensor.2d store (one box of 128 rows x 640 B ... issued as several boxes by one thread)
// grid = 1 and grid = 148 (every SM writes its own region).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_store tools/exp_store.cu -lcuda && ./exp_store
#include "../stablediffusioneo_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

using namespace sdeo;

constexpr int kBytes = 96 * 1024;
constexpr int kRowBytes = 2048;
constexpr int kRows = kBytes / kRowBytes;  // 48

__global__ void __launch_bounds__(384, 1)
store_kernel(const __grid_constant__ CUtensorMap tmY, uint8_t* y, long long* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kBytes / 16; i += 384) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, tid, 3, 4);
  fence_proxy_async_smem();
  __syncthreads();
  uint8_t* dst = y + (size_t)blockIdx.x * kBytes;
  const long long t0 = clock64();
  if (variant == 0) {
    for (int i = tid; i < kBytes / 16; i += 384) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(smem)[i];
  } else if (variant == 1) {
    for (int i = tid; i < kBytes / 32; i += 384) {
      const float4 a = reinterpret_cast<const float4*>(smem)[2 * i], b = reinterpret_cast<const float4*>(smem)[2 * i + 1];
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + (size_t)i * 32), "f"(a.x), "f"(a.y), "f"(a.z),
                   "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
    }
  } else if (variant == 2) {
    if (warp == 0 && elect_one()) {
      for (int r = 0; r < kRows; ++r)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)r * kRowBytes),
                     "r"(smem_u32(smem) + r * kRowBytes), "r"(kRowBytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (variant == 3) {
    if (elect_one()) {
      for (int r = warp; r < kRows; r += 12)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)r * kRowBytes),
                     "r"(smem_u32(smem) + r * kRowBytes), "r"(kRowBytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (variant == 4) {
    if (warp == 0 && elect_one()) {
      // boxes of 16 rows x 2048 B (fp32 x 512 is over the 256-element box limit: the map is 256 fp32 wide, two boxes per row block)
      for (int r = 0; r < kRows; r += 16)
        for (int h = 0; h < 2; ++h)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmY)),
                       "r"(smem_u32(smem) + (r * 2 + h * 16) * 1024), "r"(h * 256), "r"((int)blockIdx.x * kRows + r) : "memory");
      asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  else if (variant >= 5) {
    // the conv epilogue's mapping: 128 rows x 20 items of 8 columns; fp32 row pitch 1280 B, bf16 twin row pitch 640 B
    // 5: fp32 (32 B) + twin (16 B) per item, 6: fp32 only, 7: twin only, 8: as 5 with the values read from shared memory
    uint8_t* y2 = dst + 128 * 1280 / 2;   // (inside this CTA's 96 KB region only for the timing's sake: 80 KB + 40 KB overlap is harmless)
    for (int it = tid; it < 128 * 20; it += 384) {
      const int row = it / 20, ci = it % 20;
      float4 a = make_float4(1.f, 2.f, 3.f, 4.f), b = a;
      if (variant == 8) {
        a = reinterpret_cast<const float4*>(smem + (row * 164 + ci * 8) * 4)[0];
        b = reinterpret_cast<const float4*>(smem + (row * 164 + ci * 8) * 4)[1];
      }
      if (variant != 7)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + ((size_t)row * 1280 + ci * 32) % (kBytes - 32)), "f"(a.x), "f"(a.y),
                     "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
      if (variant != 6)
        *reinterpret_cast<uint4*>(y2 + ((size_t)row * 640 + ci * 16) % (kBytes / 2 - 16)) = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(b.x), __float_as_uint(b.y));
    }
  }
  if (variant == 9 || variant == 10) {
    // thread = row (the TMEM drain's own layout): warp w holds rows 32 (w & 3) .. +31, 32-column chunks c = w >> 2, +3, ...;
    // every thread writes 128 contiguous bytes of its fp32 row (4 x 32 B) and 64 bytes of the bf16 twin (2 x 32 B)
    uint8_t* y2 = dst + 128 * 1280 / 2;
    const int row = (warp & 3) * 32 + lane;
    for (int c = (warp >> 2) * 32; c < 160; c += 96) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + ((size_t)row * 1280 + c * 4 + j * 32) % (kBytes - 32)), "f"(1.f), "f"(2.f),
                     "f"(3.f), "f"(4.f), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
      if (variant == 9) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(y2 + ((size_t)row * 640 + c * 2 + j * 32) % (kBytes / 2 - 32)), "f"(1.f), "f"(2.f),
                       "f"(3.f), "f"(4.f), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  uint8_t* dY;
  long long* dOut;
  cudaMalloc(&dY, (size_t)148 * kBytes);
  cudaMalloc(&dOut, 148 * 8);
  CUtensorMap tmY;
  {
    cuuint64_t dims[2] = {512, (cuuint64_t)148 * kRows};
    cuuint64_t strides[1] = {kRowBytes};
    cuuint32_t box[2] = {256, 16};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dY, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  const size_t smem = kBytes + 2048;
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int vbytes[11] = {kBytes, kBytes, kBytes, kBytes, kBytes, 128 * 160 * 6, 128 * 160 * 4, 128 * 160 * 2, 128 * 160 * 6, 128 * 160 * 6, 128 * 160 * 4};
  const char* names[11] = {"st.global.v4 (16 B/lane)", "st.global.v8.f32 (32 B/lane)", "cp.async.bulk 2 KB rows, 1 thread", "cp.async.bulk 2 KB rows, 12 threads",
                          "cp.async.bulk.tensor 16 KB boxes, 1 thread", "epilogue mapping fp32 + twin", "epilogue mapping fp32 only",
                          "epilogue mapping twin only", "epilogue mapping fp32 + twin from smem", "thread = row: fp32 + twin, 32 B per lane", "thread = row: fp32 only"};
  for (int grid : {1, 148}) {
    for (int v = 0; v < 11; ++v) {
      for (int rep = 0; rep < 3; ++rep) store_kernel<<<grid, 384, smem>>>(tmY, dY, dOut, v);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), dOut, grid * 8, cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("grid %3d  %-44s  %6lld cycles (max %6lld)  %5.1f B/clk/SM\n", grid, names[v], h[grid / 2], h[grid - 1], (double)vbytes[v] / h[grid / 2]);
    }
  }
  return 0;
}
