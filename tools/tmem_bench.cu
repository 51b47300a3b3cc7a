// Microbenchmark: TMEM -> register read bandwidth (tcgen05.ld.32x32b.x32) per SM, for nw warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int UNROLL>
__global__ void k(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[UNROLL][32];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) ld32(base + ((i * UNROLL + u) * 32 & 511), r[u]);
    ldwait();
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[u][j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* d_c; uint32_t* d_s;
  cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 148 * 1024 * 4);
  const int iters = 2000;
  for (int nw : {4, 8, 16}) {
    for (int un : {1, 2}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (un == 1) k<1><<<148, nw * 32>>>(iters, d_c, d_s); else k<2><<<148, nw * 32>>>(iters, d_c, d_s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      double bytes = (double)iters * un * nw * 4096.0;
      printf("warps=%2d unroll=%d cycles=%lld  bytes/cycle/SM=%.1f  cycles per x32 load per warp=%.1f\n", nw, un, c, bytes / c, (double)c / (iters * un));
    }
  }
  return 0;
}
