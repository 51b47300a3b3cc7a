// Inverse-CDF hierarchical resampling, one warp per ray (warp-cooperative searchsorted + lerp).
// Reference: NeRFModel.resample nerf.py:225-261.  Exact arithmetic spec: SURVEY.md Appendix A.3;
// closed-form backward: Appendix B.5.
//
//  cdf_i   = (float) sum_{j<=i} (double) w_j          (CPU cumsum accumulates in fp64)
//  lo, hi  = min, max of cdf                          (nerf.py:232-233)
//  u_k     = fl(fl(k*step) + lo), step = fl((hi-lo)/(Nf+1)), k = 1..Nf      (np.linspace()[1:-1], nerf.py:243-246)
//  idx_k   = #{i : cdf_i < u_k} - 1                   (searchsorted left, nerf.py:248)
//  slope_j = fl(delta0 / fl(w_{j+1} + 1e-7)), slope_{Nc-1} = 0             (nerf.py:239, 257)
//  t_fine  = fl(t_c[idx] + fl(fl(u - cdf[idx]) * slope[idx]))               (nerf.py:259)
// idx < 0 or idx > Nf-1 makes the reference exit(0) (nerf.py:251-253); here it raises the ctx status flag.
#include "common.cuh"

#define FULL 0xffffffffu
#define PDF_WARPS 4
#define NC 64
#define NF 128

__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double u = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// builds cdf (smem), returns lo/hi/step; every lane gets the same values
__device__ __forceinline__ void build_cdf(const float* __restrict__ w_row, float* cdf, float* wsm, int lane, float& lo,
                                          float& hi, float& step) {
  float w0 = w_row[lane * 2], w1 = w_row[lane * 2 + 1];
  wsm[lane * 2] = w0;
  wsm[lane * 2 + 1] = w1;
  double s0 = (double)w0, s1 = s0 + (double)w1;
  double incl = warp_incl_scan_d(s1, lane);
  double base = incl - s1;
  float c0 = (float)(base + s0), c1 = (float)(base + s1);
  cdf[lane * 2] = c0;
  cdf[lane * 2 + 1] = c1;
  float mx = fmaxf(c0, c1), mn = fminf(c0, c1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
  }
  lo = mn;
  hi = mx;
  step = __fdiv_rn(__fsub_rn(hi, lo), (float)(NF + 1));
  __syncwarp();
}

__device__ __forceinline__ int count_less(const float* cdf, float u) {
  // number of cdf entries strictly below u; cdf is non-decreasing (w >= 0)
  int lo = 0, hi = NC;
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    if (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] < u)
        lo = mid + 1;
      else
        hi = mid;
    }
  }
  return lo;
}

__global__ void __launch_bounds__(PDF_WARPS * 32)
    sample_pdf_kernel(int64_t n, const float* __restrict__ t_coarse, const float* __restrict__ w,
                      const float* __restrict__ delta0_ptr, float* __restrict__ t_fine, int32_t* __restrict__ idx_out,
                      int* __restrict__ status) {
  __shared__ float s_cdf[PDF_WARPS][NC], s_w[PDF_WARPS][NC], s_t[PDF_WARPS][NC];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * PDF_WARPS + wib;
  if (ray >= n) return;
  float* cdf = s_cdf[wib];
  float* wsm = s_w[wib];
  float* tc = s_t[wib];
  tc[lane * 2] = t_coarse[ray * NC + lane * 2];
  tc[lane * 2 + 1] = t_coarse[ray * NC + lane * 2 + 1];
  float lo, hi, step;
  build_cdf(w + ray * NC, cdf, wsm, lane, lo, hi, step);
  const float delta0 = delta0_ptr ? *delta0_ptr : __fsub_rn(t_coarse[1], t_coarse[0]);  // ray 0 only (nerf.py:234)
  bool bad = false;
#pragma unroll
  for (int q = 0; q < NF / 32; ++q) {
    int k = q * 32 + lane;  // output slot, u index k+1
    float u = __fadd_rn(__fmul_rn((float)(k + 1), step), lo);
    int j = count_less(cdf, u) - 1;
    if (idx_out) idx_out[ray * NF + k] = j;
    if (j < 0 || j > NF - 1) bad = true;
    int jc = min(max(j, 0), NC - 1);
    float slope = (jc < NC - 1) ? __fdiv_rn(delta0, __fadd_rn(wsm[jc + 1], 1e-7f)) : 0.f;
    float tf = __fadd_rn(tc[jc], __fmul_rn(__fsub_rn(u, cdf[jc]), slope));
    t_fine[ray * NF + k] = tf;
  }
  if (__any_sync(FULL, bad) && lane == 0) atomicOr(status, 1);
}

// backward (B.5): g_tf -> g_w.   u, idx, delta0, t_coarse carry no gradient.
__global__ void __launch_bounds__(PDF_WARPS * 32)
    sample_pdf_bwd_kernel(int64_t n, const float* __restrict__ t_coarse, const float* __restrict__ w,
                          const float* __restrict__ delta0_ptr, const float* __restrict__ g_tf,
                          float* __restrict__ g_w) {
  __shared__ float s_cdf[PDF_WARPS][NC], s_w[PDF_WARPS][NC], s_gcdf[PDF_WARPS][NC], s_gslope[PDF_WARPS][NC];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * PDF_WARPS + wib;
  if (ray >= n) return;
  float* cdf = s_cdf[wib];
  float* wsm = s_w[wib];
  float* gcdf = s_gcdf[wib];
  float* gsl = s_gslope[wib];
  gcdf[lane * 2] = gcdf[lane * 2 + 1] = 0.f;
  gsl[lane * 2] = gsl[lane * 2 + 1] = 0.f;
  float lo, hi, step;
  build_cdf(w + ray * NC, cdf, wsm, lane, lo, hi, step);
  const float delta0 = delta0_ptr ? *delta0_ptr : __fsub_rn(t_coarse[1], t_coarse[0]);
#pragma unroll
  for (int q = 0; q < NF / 32; ++q) {
    int k = q * 32 + lane;
    float u = __fadd_rn(__fmul_rn((float)(k + 1), step), lo);
    int j = min(max(count_less(cdf, u) - 1, 0), NC - 1);
    float g = g_tf[ray * NF + k];
    float slope = (j < NC - 1) ? __fdiv_rn(delta0, __fadd_rn(wsm[j + 1], 1e-7f)) : 0.f;
    atomicAdd(&gcdf[j], -g * slope);
    if (j < NC - 1) atomicAdd(&gsl[j], g * (u - cdf[j]));
  }
  __syncwarp();
  // g_w_j = sum_{i>=j} g_cdf_i  (cumsum backward)  +  g_slope_{j-1} * (-delta0/(w_j+eps)^2)
  float a0 = gcdf[lane * 2], a1 = gcdf[lane * 2 + 1];
  float lsum = a0 + a1;
  float v = lsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float x = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += x;
  }
  float after = v - lsum;
  float gw1 = after + a1;
  float gw0 = gw1 + a0;
  int j0 = lane * 2, j1 = lane * 2 + 1;
  if (j0 > 0) {
    float d = wsm[j0] + 1e-7f;
    gw0 += gsl[j0 - 1] * (-delta0 / (d * d));
  }
  {
    float d = wsm[j1] + 1e-7f;
    gw1 += gsl[j1 - 1] * (-delta0 / (d * d));
  }
  g_w[ray * NC + j0] = gw0;
  g_w[ray * NC + j1] = gw1;
}

extern "C" int nt_sample_pdf(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                             float* t_fine, int32_t* idx, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_coarse && w && t_fine, "null pointer");
  NT_REQUIRE(ctx->n_coarse == NC && ctx->n_fine == NF, "sample_pdf is built for Nc=64, Nf=128");
  if (n <= 0) return NT_OK;
  sample_pdf_kernel<<<(unsigned)((n + PDF_WARPS - 1) / PDF_WARPS), PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      n, t_coarse, w, delta0, t_fine, idx, ctx->d_flags + 1);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_sample_pdf_backward(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w,
                                      const float* delta0, const float* g_t_fine, float* g_w, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_coarse && w && g_t_fine && g_w, "null pointer");
  NT_REQUIRE(ctx->n_coarse == NC && ctx->n_fine == NF, "sample_pdf is built for Nc=64, Nf=128");
  if (n <= 0) return NT_OK;
  sample_pdf_bwd_kernel<<<(unsigned)((n + PDF_WARPS - 1) / PDF_WARPS), PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      n, t_coarse, w, delta0, g_t_fine, g_w);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_check_status(nt_ctx* ctx, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx, "null ctx");
  int h = 0;
  NT_CUDA(cudaMemcpyAsync(&h, ctx->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (h) {
    NT_CUDA(cudaMemsetAsync(ctx->d_flags + 1, 0, sizeof(int), (cudaStream_t)stream));
    nt_set_error("resample index out of range (the reference exits here, nerf.py:251-253)");
    return NT_ERR_RANGE;
  }
  return NT_OK;
}
