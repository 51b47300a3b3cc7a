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

__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double u = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// builds cdf (smem), returns lo/hi/step; every lane gets the same values.  NC = 32 * EC coarse samples, lane owns
// elements lane*EC .. lane*EC + EC-1.
template <int EC>
__device__ __forceinline__ void build_cdf(const float* __restrict__ w_row, float* cdf, float* wsm, int lane, int nf, float& lo,
                                          float& hi, float& step) {
  float wv[EC];
  double pre[EC], run = 0.0;
#pragma unroll
  for (int k = 0; k < EC; ++k) {
    wv[k] = w_row[lane * EC + k];
    wsm[lane * EC + k] = wv[k];
    run += (double)wv[k];
    pre[k] = run;
  }
  const double base = warp_incl_scan_d(run, lane) - run;
  float mx = -INFINITY, mn = INFINITY;
#pragma unroll
  for (int k = 0; k < EC; ++k) {
    const float c = (float)(base + pre[k]);  // every prefix rounded to fp32, like CPU cumsum
    cdf[lane * EC + k] = c;
    mx = fmaxf(mx, c);
    mn = fminf(mn, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
  }
  lo = mn;
  hi = mx;
  step = __fdiv_rn(__fsub_rn(hi, lo), (float)(nf + 1));
  __syncwarp();
}

template <int NC>
__device__ __forceinline__ int count_less(const float* cdf, float u) {
  // number of cdf entries strictly below u; cdf is non-decreasing (w >= 0)
  int lo = 0, hi = NC;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if ((1 << it) <= NC && lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] < u)
        lo = mid + 1;
      else
        hi = mid;
    }
  }
  return lo;
}

// Forward.  Each lane owns nf/32 CONSECUTIVE outputs k: u is non-decreasing in k, so the lane needs ONE 7-step binary search
// (for its last u); the counts of its other u's lie between the previous lane's last count (one shuffle) and its own, an
// interval of ~nc/32 entries on average that a short forward scan resolves (the strided assignment of the first version
// did nf/32 independent binary searches per lane: 28 bank-conflicting shared-memory reads instead of ~10).  The slope
// delta0 / (w_{j+1} + 1e-7) is formed once per coarse bin (nc divisions per ray instead of nf) and the lane's outputs
// leave as one 16-byte store when nf = 128.  Same operations on the same operands as before: idx / t_fine stay bit-exact.
template <int EC>
__global__ void __launch_bounds__(PDF_WARPS * 32)
    sample_pdf_kernel(int64_t n, int nf, const float* __restrict__ t_coarse, const float* __restrict__ w,
                      const float* __restrict__ delta0_ptr, float* __restrict__ t_fine, int32_t* __restrict__ idx_out,
                      int* __restrict__ status) {
  constexpr int NC = 32 * EC;
  __shared__ float s_cdf[PDF_WARPS][NC], s_w[PDF_WARPS][NC], s_t[PDF_WARPS][NC], s_slope[PDF_WARPS][NC];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * PDF_WARPS + wib;
  if (ray >= n) return;
  float* cdf = s_cdf[wib];
  float* wsm = s_w[wib];
  float* tc = s_t[wib];
  float* slp = s_slope[wib];
#pragma unroll
  for (int k = 0; k < EC; ++k) tc[lane * EC + k] = t_coarse[ray * NC + lane * EC + k];
  float lo, hi, step;
  build_cdf<EC>(w + ray * NC, cdf, wsm, lane, nf, lo, hi, step);
  const float delta0 = delta0_ptr ? *delta0_ptr : __fsub_rn(t_coarse[1], t_coarse[0]);  // ray 0 only (nerf.py:234)
#pragma unroll
  for (int k = 0; k < EC; ++k) {
    const int j = lane * EC + k;
    slp[j] = (j < NC - 1) ? __fdiv_rn(delta0, __fadd_rn(wsm[j + 1], 1e-7f)) : 0.f;
  }
  __syncwarp();
  const int kf = nf >> 5;        // outputs per lane (nf is a multiple of 32, at most 224)
  const int k0 = lane * kf;      // first output slot of this lane; u index = slot + 1
  const float u_last = __fadd_rn(__fmul_rn((float)(k0 + kf), step), lo);
  const int c_last = count_less<NC>(cdf, u_last);
  int c = __shfl_up_sync(FULL, c_last, 1);
  if (lane == 0) c = 0;
  bool bad = false;
  float tf[8];
  int jj[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < kf) {
      const float u = __fadd_rn(__fmul_rn((float)(k0 + i + 1), step), lo);
      if (i == kf - 1) {
        c = c_last;
      } else {
        while (c < c_last && cdf[c] < u) ++c;  // #{cdf < u}: every entry below c is < an earlier (smaller) u already
      }
      const int j = c - 1;
      jj[i] = j;
      if (j < 0 || j > nf - 1) bad = true;  // nerf.py:251 compares against num_fine - 1
      const int jc = min(max(j, 0), NC - 1);
      tf[i] = __fadd_rn(tc[jc], __fmul_rn(__fsub_rn(u, cdf[jc]), slp[jc]));
    }
  }
  float* out = t_fine + ray * nf + k0;
  if (kf == 4) {
    *reinterpret_cast<float4*>(out) = make_float4(tf[0], tf[1], tf[2], tf[3]);
    if (idx_out) *reinterpret_cast<int4*>(idx_out + ray * nf + k0) = make_int4(jj[0], jj[1], jj[2], jj[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < kf) {
        out[i] = tf[i];
        if (idx_out) idx_out[ray * nf + k0 + i] = jj[i];
      }
  }
  if (__any_sync(FULL, bad) && lane == 0) atomicOr(status, 1);
}

// backward (B.5): g_tf (+ g_tf2, the second gradient path into t_fine, summed on the fly) -> g_w.
// u, idx, delta0, t_coarse carry no gradient.
template <int EC>
__global__ void __launch_bounds__(PDF_WARPS * 32)
    sample_pdf_bwd_kernel(int64_t n, int nf, const float* __restrict__ t_coarse, const float* __restrict__ w,
                          const float* __restrict__ delta0_ptr, const float* __restrict__ g_tf,
                          const float* __restrict__ g_tf2, float* __restrict__ g_w) {
  constexpr int NC = 32 * EC;
  __shared__ float s_cdf[PDF_WARPS][NC], s_w[PDF_WARPS][NC], s_gcdf[PDF_WARPS][NC], s_gslope[PDF_WARPS][NC];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * PDF_WARPS + wib;
  if (ray >= n) return;
  float* cdf = s_cdf[wib];
  float* wsm = s_w[wib];
  float* gcdf = s_gcdf[wib];
  float* gsl = s_gslope[wib];
#pragma unroll
  for (int k = 0; k < EC; ++k) gcdf[lane * EC + k] = gsl[lane * EC + k] = 0.f;
  float lo, hi, step;
  build_cdf<EC>(w + ray * NC, cdf, wsm, lane, nf, lo, hi, step);
  const float delta0 = delta0_ptr ? *delta0_ptr : __fsub_rn(t_coarse[1], t_coarse[0]);
  for (int q = 0; q < nf / 32; ++q) {
    int k = q * 32 + lane;
    float u = __fadd_rn(__fmul_rn((float)(k + 1), step), lo);
    int j = min(max(count_less<NC>(cdf, u) - 1, 0), NC - 1);
    float g = g_tf[ray * nf + k];
    if (g_tf2) g += g_tf2[ray * nf + k];
    float slope = (j < NC - 1) ? __fdiv_rn(delta0, __fadd_rn(wsm[j + 1], 1e-7f)) : 0.f;
    atomicAdd(&gcdf[j], -g * slope);
    if (j < NC - 1) atomicAdd(&gsl[j], g * (u - cdf[j]));
  }
  __syncwarp();
  // g_w_j = sum_{i>=j} g_cdf_i  (cumsum backward)  +  g_slope_{j-1} * (-delta0/(w_j+eps)^2)
  float a[EC], lsum = 0.f;
#pragma unroll
  for (int k = 0; k < EC; ++k) {
    a[k] = gcdf[lane * EC + k];
    lsum += a[k];
  }
  float v = lsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float x = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += x;
  }
  float run = v - lsum;  // lanes after this one
#pragma unroll
  for (int k = EC - 1; k >= 0; --k) {
    run += a[k];
    const int j = lane * EC + k;
    float gw = run;
    if (j > 0) {
      float d = wsm[j] + 1e-7f;
      gw += gsl[j - 1] * (-delta0 / (d * d));
    }
    g_w[ray * NC + j] = gw;
  }
}

#define NT_DISPATCH_EC(NCV, CALL)                        \
  switch ((NCV) / 32) {                                  \
    case 1: { constexpr int E = 1; CALL; } break;        \
    case 2: { constexpr int E = 2; CALL; } break;        \
    case 3: { constexpr int E = 3; CALL; } break;        \
    default: { constexpr int E = 4; CALL; } break;       \
  }

int nt_launch_sample_pdf_backward(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                                  const float* g_t_fine, const float* g_t_fine2, float* g_w, cudaStream_t st) {
  if (n <= 0) return NT_OK;
  const unsigned blocks = (unsigned)((n + PDF_WARPS - 1) / PDF_WARPS);
  NT_DISPATCH_EC(ctx->n_coarse, (sample_pdf_bwd_kernel<E><<<blocks, PDF_WARPS * 32, 0, st>>>(n, ctx->n_fine, t_coarse, w, delta0,
                                                                                            g_t_fine, g_t_fine2, g_w)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_sample_pdf(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                             float* t_fine, int32_t* idx, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_coarse && w && t_fine, "null pointer");
  if (n <= 0) return NT_OK;
  const unsigned blocks = (unsigned)((n + PDF_WARPS - 1) / PDF_WARPS);
  NT_DISPATCH_EC(ctx->n_coarse, (sample_pdf_kernel<E><<<blocks, PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(
                                    n, ctx->n_fine, t_coarse, w, delta0, t_fine, idx, ctx->d_flags + 1)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_sample_pdf_backward(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w,
                                      const float* delta0, const float* g_t_fine, float* g_w, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_coarse && w && g_t_fine && g_w, "null pointer");
  return nt_launch_sample_pdf_backward(ctx, n, t_coarse, w, delta0, g_t_fine, nullptr, g_w, (cudaStream_t)stream);
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
