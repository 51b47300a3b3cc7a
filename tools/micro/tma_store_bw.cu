// Micro-benchmark: how fast can 148 CTAs push shared-memory tiles to HBM?  (context for the stash-writing training kernels)
//   mode 0: TMA 2D tensor stores, box 64 x 128 bf16 (128 B rows, global row pitch 512 B) - the stash layout
//   mode 1: TMA linear bulk stores of 16 KB (tile-contiguous layout)
//   mode 2: st.global.v4 from registers, fully coalesced
//   mode 3: TMA 2D tensor stores, box 64 x 128 bf16, global row pitch 128 B (contiguous rows)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_bw tma_store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, uint8_t* g, int iters, int mode, int depth) {
  extern __shared__ __align__(1024) uint8_t smem[];  // 8 x 16 KB
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  if (mode == 2) {
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (int it = 0; it < iters; ++it) {
      // 16 KB per iteration per CTA: 128 threads x 8 x 16 B
      uint8_t* dst = g + ((size_t)(it * gridDim.x + blockIdx.x)) * 16384;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(dst + (j * 128 + threadIdx.x) * 16) = v;
    }
    return;
  }
  if (threadIdx.x != 0) return;
  for (int it = 0; it < iters; ++it) {
    const uint32_t src = sb + (it & 7) * 16384;
    const int tile = it * gridDim.x + blockIdx.x;  // global tile index: 128 rows
    if (mode == 1) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 16384;" ::"l"(g + (size_t)tile * 16384), "r"(src) : "memory");
    } else {
      // mode 0: tensor [rows][256 cols]: tile -> (chunk = tile & 3, rows (tile >> 2) * 128);  mode 3: tensor [rows][64]
      const int c0 = mode == 0 ? (tile & 3) * 64 : 0, c1 = mode == 0 ? (tile >> 2) * 128 : tile * 128;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(c0), "r"(c1), "r"(src) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int grid = 148, iters = 2048;                 // 148 x 2048 x 16 KB = 4.97 GB
  const size_t bytes = (size_t)grid * iters * 16384;
  uint8_t* g;
  CK(cudaMalloc(&g, bytes));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 4; ++mode) {
    CUtensorMap map;
    const uint64_t rows = (mode == 0) ? bytes / 512 : bytes / 128;
    cuuint64_t dims[2] = {(cuuint64_t)(mode == 0 ? 256 : 64), rows};
    cuuint64_t strides[1] = {(cuuint64_t)(mode == 0 ? 512 : 128)};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int depth : {1, 2, 4, 7}) {
      if (mode == 2 && depth != 1) continue;
      k<<<grid, 128, 8 * 16384>>>(map, g, 64, mode, depth);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      k<<<grid, 128, 8 * 16384>>>(map, g, iters, mode, depth);
      cudaEventRecord(b);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      printf("mode %d depth %d: %.3f ms  %.2f TB/s  %.1f B/clk/SM @1.9GHz\n", mode, depth, ms, bytes / ms / 1e9, bytes / ms / 1e6 / 148 / 1.9e3 * 1e3 / 1e3);
    }
  }
  return 0;
}
