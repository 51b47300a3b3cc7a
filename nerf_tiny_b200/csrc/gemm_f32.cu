// fp32 SIMT GEMM with fused epilogues — the NT_PREC_FP32 ("fp32 mode", rgb within 1e-3) MLP path and the
// layer-major backward.  Replaces the ATen addmm / relu / abs / sigmoid / autograd launches behind
// Network.forward (nerf.py:101-124) and loss.backward() (nerf.py:473).
//
//   C[M,N] = epi( sum_seg  A_seg (.) B_seg )         up to two K-segments (skip / view concatenations,
//                                                     nerf.py:109, :118, are two segments — no cat is materialised)
//   A stored [M][K] (a_km = false) or [K][M] (true); B stored [N][K] (b_kn = false) or [K][N] (true).
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, k-major shared tiles.
#include "common.cuh"

#define BM 128
#define BN 128
#define BK 16
#define PADM 4

template <bool KM>  // KM: source stored [K][X] (X contiguous); else [X][K] (K contiguous)
__device__ __forceinline__ void load_tile(float (*dst)[BM + PADM], const float* __restrict__ src, int ld, int x0, int xmax,
                                          int k0, int kmax, int tid) {
  if (KM) {
    // 16 rows (k) x 128 cols (x): thread -> (k = tid/32 + 8*i, x = (tid%32)*4 .. +3)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int k = (tid >> 5) + 8 * i;
      int x = (tid & 31) * 4;
      int gk = k0 + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gx = x0 + x + j;
        dst[k][x + j] = (gk < kmax && gx < xmax) ? src[(int64_t)gk * ld + gx] : 0.f;
      }
    }
  } else {
    // 128 rows (x) x 16 cols (k): thread -> (x = tid/16 + 16*i, k = tid%16)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int x = (tid >> 4) + 16 * i;
      int k = tid & 15;
      int gx = x0 + x, gk = k0 + k;
      dst[k][x] = (gk < kmax && gx < xmax) ? src[(int64_t)gx * ld + gk] : 0.f;
    }
  }
}

template <bool A_KM, bool B_KN>
__global__ void __launch_bounds__(256)
    gemm_f32_kernel(int M, int N, GemmSeg s0, GemmSeg s1, float* __restrict__ C, int ldc, GemmEpi epi, int split_k) {
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> n, ty -> m
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int seg = 0; seg < 2; ++seg) {
    const GemmSeg s = seg == 0 ? s0 : s1;
    if (s.K <= 0 || s.A == nullptr) continue;
    // split-K: this z-slice covers [kb, ke)
    int kb = 0, ke = s.K;
    if (split_k > 1) {
      int per = ((s.K + split_k - 1) / split_k + BK - 1) / BK * BK;
      kb = blockIdx.z * per;
      ke = min(s.K, kb + per);
    }
    for (int k0 = kb; k0 < ke; k0 += BK) {
      load_tile<A_KM>(As, s.A, s.lda, m0, M, k0, ke, tid);
      load_tile<B_KN>(Bs, s.B, s.ldb, n0, N, k0, ke, tid);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[8], b[8];
        *(float4*)&a[0] = *(const float4*)&As[k][ty * 4];
        *(float4*)&a[4] = *(const float4*)&As[k][64 + ty * 4];
        *(float4*)&b[0] = *(const float4*)&Bs[k][tx * 4];
        *(float4*)&b[4] = *(const float4*)&Bs[k][64 + tx * 4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const bool first_slice = (split_k <= 1) || blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int nn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (nn >= N) continue;
      float v = acc[i][j];
      if (epi.bias && first_slice) v += epi.bias[nn];
      if (epi.pre_out) epi.pre_out[(int64_t)m * epi.ldpre + nn] = v;
      if (epi.act == ACT_RELU)
        v = fmaxf(v, 0.f);
      else if (epi.act == ACT_ABS)
        v = fabsf(v);
      else if (epi.act == ACT_SIGMOID)
        v = 1.f / (1.f + expf(-v));
      float* c = C + (int64_t)m * ldc + nn;
      if (epi.atomic) {
        atomicAdd(c, v);
        continue;
      }
      if (epi.accumulate) v += *c;
      if (epi.mask && !(epi.mask[(int64_t)m * epi.ldmask + nn] > 0.f)) v = 0.f;  // ReLU'(out) from the stored output
      *c = v;
    }
  }
}

int nt_launch_gemm(nt_ctx* ctx, int M, int N, GemmSeg s0, GemmSeg s1, bool a_km, bool b_kn, float* C, int ldc,
                   GemmEpi epi, int split_k, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NT_OK;
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !epi.atomic) {
    nt_set_error("split-K GEMM needs the atomic epilogue");
    return NT_ERR_INVALID;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split_k);
  if (grid.y > 65535) {
    nt_set_error("GEMM M too large for one launch (chunk the rays)");
    return NT_ERR_INVALID;
  }
  if (!a_km && !b_kn)
    gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(M, N, s0, s1, C, ldc, epi, split_k);
  else if (!a_km && b_kn)
    gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(M, N, s0, s1, C, ldc, epi, split_k);
  else if (a_km && !b_kn)
    gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(M, N, s0, s1, C, ldc, epi, split_k);
  else
    gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(M, N, s0, s1, C, ldc, epi, split_k);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

// column sums of G[rows, cols] (ld) accumulated into out[cols] (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ G, int64_t rows, int cols, int ld, float* __restrict__ out,
                              int rows_per_block) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += G[r * ld + c];
  atomicAdd(out + c, acc);
}

int nt_launch_colsum(nt_ctx* ctx, const float* G, int64_t rows, int cols, int ld, float* out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return NT_OK;
  int rpb = 256;
  int tx = cols >= 128 ? 128 : (cols >= 32 ? 32 : cols);
  dim3 grid((cols + tx - 1) / tx, (unsigned)((rows + rpb - 1) / rpb));
  if (grid.y > 65535) {
    rpb = (int)((rows + 65534) / 65535);
    grid.y = (unsigned)((rows + rpb - 1) / rpb);
  }
  colsum_kernel<<<grid, tx, 0, st>>>(G, rows, cols, ld, out, rpb);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
