// which L2 cache-hint forms run on sm_100a?  nvcc -gencode arch=compute_100a,code=sm_100a -o hint_test hint_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
__global__ void k(const uint8_t* src, uint8_t* dst, int mode, const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint64_t pl, pf;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
  if (mode == 0) {  // cp.async with hint
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sb + threadIdx.x * 16), "l"(src + threadIdx.x * 16), "l"(pl) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (mode == 1 && threadIdx.x == 0) {  // bulk load with hint
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4096) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(sb), "l"(src), "r"(4096), "r"(b), "l"(pl) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  } else if (mode == 2 && threadIdx.x == 0) {  // bulk store with hint
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], 4096, %2;" ::"l"(dst), "r"(sb), "l"(pf) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (mode == 4 && threadIdx.x == 0) {  // 2D tensor store with hint
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(&map), "r"(0), "r"(0), "r"(sb), "l"(pf) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) dst[8192] = smem[0] + (uint8_t)(pl ^ pf);
}
int main(int argc, char** argv) {
  uint8_t *s, *d;
  cudaMalloc(&s, 1 << 20); cudaMalloc(&d, 1 << 20);
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  CUtensorMap map;
  cuuint64_t dims[2] = {256, 1024};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {64, 32}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  k<<<1, 128, 8192>>>(s, d, mode, map);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d: %s\n", mode, cudaGetErrorString(e));
  return 0;
}
