// Alpha compositing, one warp per ray.
// Reference: get_density nerf.py:263-272, color_cum nerf.py:274-281, the merge + per-channel sort of
// render_rays nerf.py:302-321, ray_loss nerf.py:325-331.  Closed-form backward: SURVEY.md Appendix B.2/B.3.
//
// Transmittance uses the INCLUSIVE prefix sum (nerf.py:267).  CPU torch.cumsum accumulates in fp64 and
// rounds every prefix to fp32, so the warp-shuffle scans below run in fp64.
#include "common.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double u = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += u;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// reverse (suffix) inclusive scan across lanes
__device__ __forceinline__ float warp_suffix_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += u;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// Generic per-ray compositing over EPL elements per lane (elements lane*EPL .. lane*EPL+EPL-1).
// a[] = sigma*delta.  Produces w[], and T*E (needed by backward).
// ---------------------------------------------------------------------------------------------
template <int EPL>
__device__ __forceinline__ void density_weights(const float (&a)[EPL], int lane, float (&w)[EPL], float (&te)[EPL]) {
  double run = 0.0;
  double pre[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    run += (double)a[k];
    pre[k] = run;
  }
  double incl = warp_incl_scan(run, lane);
  double base = incl - run;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    float S = (float)(base + pre[k]);  // each prefix rounded to fp32, like CPU cumsum
    float T = expf(-S);
    float E = expf(-a[k]);
    w[k] = T * (1.f - E);
    te[k] = T * E;
  }
}

// ---------------------------------------------------------------------------------------------
// coarse: Nc = 64 -> 2 per lane; delta = (far-near)/Nc (nerf.py:293)
// ---------------------------------------------------------------------------------------------
template <int EPL>
__global__ void composite_coarse_kernel(int64_t n, const float* __restrict__ near_, const float* __restrict__ far_,
                                        const float* __restrict__ rgb, const float* __restrict__ sigma,
                                        float* __restrict__ weights, float* __restrict__ c_out) {
  constexpr int P = EPL * 32;
  int lane = threadIdx.x & 31;
  int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  const float delta = __fdiv_rn(__fsub_rn(far_[ray], near_[ray]), (float)P);
  float a[EPL], w[EPL], te[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) a[k] = __fmul_rn(delta, sigma[ray * P + lane * EPL + k]);
  density_weights<EPL>(a, lane, w, te);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const float* c = rgb + (ray * P + lane * EPL + k) * 3;
    c0 += w[k] * c[0];
    c1 += w[k] * c[1];
    c2 += w[k] * c[2];
    weights[ray * P + lane * EPL + k] = w[k];
  }
  c0 = warp_sum(c0);
  c1 = warp_sum(c1);
  c2 = warp_sum(c2);
  if (lane == 0) {
    c_out[ray * 3 + 0] = c0;
    c_out[ray * 3 + 1] = c1;
    c_out[ray * 3 + 2] = c2;
  }
}

// ---------------------------------------------------------------------------------------------
// piecewise reference methods on caller-supplied tensors: get_density (nerf.py:263-272) and color_cum (nerf.py:274-281),
// P = 32*EPL samples per ray, one warp per ray, same arithmetic as the fused kernels above
// ---------------------------------------------------------------------------------------------
template <int EPL>
__global__ void get_density_kernel(int64_t n, const float* __restrict__ delta, const float* __restrict__ sigma,
                                   float* __restrict__ weights) {
  constexpr int P = EPL * 32;
  int lane = threadIdx.x & 31;
  int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  float a[EPL], w[EPL], te[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) a[k] = __fmul_rn(delta[ray * P + lane * EPL + k], sigma[ray * P + lane * EPL + k]);
  density_weights<EPL>(a, lane, w, te);
#pragma unroll
  for (int k = 0; k < EPL; ++k) weights[ray * P + lane * EPL + k] = w[k];
}

template <int EPL>
__global__ void color_cum_kernel(int64_t n, const float* __restrict__ weights, const float* __restrict__ rgb,
                                 float* __restrict__ c_out) {
  constexpr int P = EPL * 32;
  int lane = threadIdx.x & 31;
  int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int64_t e = ray * P + lane * EPL + k;
    const float w = weights[e];
    c0 += w * rgb[e * 3 + 0];
    c1 += w * rgb[e * 3 + 1];
    c2 += w * rgb[e * 3 + 2];
  }
  c0 = warp_sum(c0);
  c1 = warp_sum(c1);
  c2 = warp_sum(c2);
  if (lane == 0) {
    c_out[ray * 3 + 0] = c0;
    c_out[ray * 3 + 1] = c1;
    c_out[ray * 3 + 2] = c2;
  }
}

template <int EPL>
__global__ void composite_coarse_bwd_kernel(int64_t n, const float* __restrict__ near_, const float* __restrict__ far_,
                                            const float* __restrict__ rgb, const float* __restrict__ sigma,
                                            const float* __restrict__ g_c, const float* __restrict__ g_w_ext,
                                            float* __restrict__ g_rgb, float* __restrict__ g_sigma, int accumulate) {
  constexpr int P = EPL * 32;
  int lane = threadIdx.x & 31;
  int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  const float delta = __fdiv_rn(__fsub_rn(far_[ray], near_[ray]), (float)P);
  float a[EPL], w[EPL], te[EPL], gw[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) a[k] = __fmul_rn(delta, sigma[ray * P + lane * EPL + k]);
  density_weights<EPL>(a, lane, w, te);
  const float g0 = g_c[ray * 3], g1 = g_c[ray * 3 + 1], g2 = g_c[ray * 3 + 2];
  float lsum = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    int64_t e = ray * P + lane * EPL + k;
    const float* c = rgb + e * 3;
    gw[k] = c[0] * g0 + c[1] * g1 + c[2] * g2 + (g_w_ext ? g_w_ext[e] : 0.f);
    if (accumulate) {  // the coarse samples also received a gradient from the fine compositing: add to it
      g_rgb[e * 3 + 0] += w[k] * g0;
      g_rgb[e * 3 + 1] += w[k] * g1;
      g_rgb[e * 3 + 2] += w[k] * g2;
    } else {
      g_rgb[e * 3 + 0] = w[k] * g0;
      g_rgb[e * 3 + 1] = w[k] * g1;
      g_rgb[e * 3 + 2] = w[k] * g2;
    }
    lsum += gw[k] * w[k];
  }
  // suffix sums of g_w*w: lanes after this one, then within the lane from the back
  float after = warp_suffix_scan(lsum, lane) - lsum;
  float run = after;
#pragma unroll
  for (int k = EPL - 1; k >= 0; --k) {
    run += gw[k] * w[k];
    float ga = te[k] * gw[k] - run;  // B.2
    if (accumulate)
      g_sigma[ray * P + lane * EPL + k] += ga * delta;
    else
      g_sigma[ray * P + lane * EPL + k] = ga * delta;
  }
}

// ---------------------------------------------------------------------------------------------
// fine: merge coarse|fine, sort each of the 5 channels independently (nerf.py:308), composite.
// Bitonic sort of 256 (key,index) pairs per channel in shared memory, one warp per ray.
// Ties are broken by the original index = the order a stable sort (ATen CPU) produces.
// ---------------------------------------------------------------------------------------------
#define FINE_PAD 256
#define FINE_WARPS 4

__device__ __forceinline__ bool pair_less(float ka, int ia, float kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

__device__ __forceinline__ void warp_bitonic_256(float* key, uint8_t* idx, int lane) {
  for (int k = 2; k <= FINE_PAD; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int q = 0; q < FINE_PAD / 64; ++q) {
        int t = lane + 32 * q;  // 128 compare-exchange pairs
        int i = 2 * t - (t & (j - 1));
        int l = i + j;
        bool up = ((i & k) == 0);
        float ki = key[i], kl = key[l];
        int ii = idx[i], il = idx[l];
        bool sw = up ? pair_less(kl, il, ki, ii) : pair_less(ki, ii, kl, il);
        if (sw) {
          key[i] = kl;
          key[l] = ki;
          idx[i] = (uint8_t)il;
          idx[l] = (uint8_t)ii;
        }
      }
      __syncwarp();
    }
  }
}

// loads channel ch of the merged (coarse|fine) array into key[0..255] (+inf padded), idx = identity
__device__ __forceinline__ void load_channel(int ch, int nc, int nf, int64_t ray, const float* t_c, const float* rgb_c,
                                             const float* sigma_c, const float* t_f, const float* rgb_f,
                                             const float* sigma_f, float* key, uint8_t* idx, int lane) {
  const int tot = nc + nf;
  for (int e = lane; e < FINE_PAD; e += 32) {
    float v = __int_as_float(0x7f800000);
    if (e < tot) {
      bool co = e < nc;
      int64_t s = co ? ray * nc + e : ray * nf + (e - nc);
      if (ch == 0)
        v = co ? t_c[s] : t_f[s];
      else if (ch == 4)
        v = co ? sigma_c[s] : sigma_f[s];
      else
        v = co ? rgb_c[s * 3 + ch - 1] : rgb_f[s * 3 + ch - 1];
    }
    key[e] = v;
    idx[e] = (uint8_t)e;
  }
  __syncwarp();
}

template <bool BWD, int EPL>   // EPL = (nc + nf) / 32
__global__ void __launch_bounds__(FINE_WARPS * 32)
    composite_fine_kernel(int64_t n, int nc, int nf, const float* __restrict__ t_c, const float* __restrict__ rgb_c,
                          const float* __restrict__ sigma_c, const float* __restrict__ t_f,
                          const float* __restrict__ rgb_f, const float* __restrict__ sigma_f, float last,
                          float* __restrict__ c_out, float* __restrict__ weights, uint8_t* __restrict__ perm_out,
                          // backward only
                          const uint8_t* __restrict__ perm_in, const float* __restrict__ g_c, float* __restrict__ g_rgb_c,
                          float* __restrict__ g_sigma_c, float* __restrict__ g_rgb_f, float* __restrict__ g_sigma_f,
                          float* __restrict__ g_t_f) {
  __shared__ float s_key[FINE_WARPS][5][FINE_PAD];
  __shared__ uint8_t s_idx[FINE_WARPS][5][FINE_PAD];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * FINE_WARPS + wib;
  if (ray >= n) return;
  const int tot = nc + nf;
  float(*key)[FINE_PAD] = s_key[wib];
  uint8_t(*idx)[FINE_PAD] = s_idx[wib];

  if (!BWD) {
    for (int ch = 0; ch < 5; ++ch) {
      load_channel(ch, nc, nf, ray, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, key[ch], idx[ch], lane);
      warp_bitonic_256(key[ch], idx[ch], lane);
      if (perm_out)
        for (int e = lane; e < tot; e += 32) perm_out[(ray * 5 + ch) * tot + e] = idx[ch][e];
    }
  } else {
    // re-gather the sorted channels through the stored permutations
    for (int ch = 0; ch < 5; ++ch) {
      for (int e = lane; e < tot; e += 32) {
        int src = perm_in[(ray * 5 + ch) * tot + e];
        bool co = src < nc;
        int64_t s = co ? ray * nc + src : ray * nf + (src - nc);
        float v;
        if (ch == 0)
          v = co ? t_c[s] : t_f[s];
        else if (ch == 4)
          v = co ? sigma_c[s] : sigma_f[s];
        else
          v = co ? rgb_c[s * 3 + ch - 1] : rgb_f[s * 3 + ch - 1];
        key[ch][e] = v;
        idx[ch][e] = (uint8_t)src;
      }
    }
    __syncwarp();
  }

  // compositing over the sorted channels; lane owns elements lane*6 .. lane*6+5
  float a[EPL], w[EPL], te[EPL], dl[EPL], sg[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    int e = lane * EPL + k;
    dl[k] = (e < tot - 1) ? __fsub_rn(key[0][e + 1], key[0][e]) : last;  // nerf.py:315
    sg[k] = key[4][e];
    a[k] = __fmul_rn(dl[k], sg[k]);
  }
  density_weights<EPL>(a, lane, w, te);

  if (!BWD) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      int e = lane * EPL + k;
      c0 += w[k] * key[1][e];
      c1 += w[k] * key[2][e];
      c2 += w[k] * key[3][e];
      if (weights) weights[ray * tot + e] = w[k];
    }
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    c2 = warp_sum(c2);
    if (lane == 0) {
      c_out[ray * 3 + 0] = c0;
      c_out[ray * 3 + 1] = c1;
      c_out[ray * 3 + 2] = c2;
    }
    return;
  }

  // ---- backward (B.2, B.3) ----
  const float g0 = g_c[ray * 3], g1 = g_c[ray * 3 + 1], g2 = g_c[ray * 3 + 2];
  float gw[EPL];
  float lsum = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    int e = lane * EPL + k;
    gw[k] = key[1][e] * g0 + key[2][e] * g1 + key[3][e] * g2;
    lsum += gw[k] * w[k];
  }
  float after = warp_suffix_scan(lsum, lane) - lsum;
  float run = after;
  float gdl[EPL], gsg[EPL];
#pragma unroll
  for (int k = EPL - 1; k >= 0; --k) {
    run += gw[k] * w[k];
    float ga = te[k] * gw[k] - run;
    gsg[k] = ga * dl[k];
    gdl[k] = ga * sg[k];
  }
  // g_t_i = g_delta_{i-1} - g_delta_i, delta_{tot-1} is the constant `last`
  __syncwarp();
  // reuse key[0] as scratch for g_delta (sorted t no longer needed)
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    int e = lane * EPL + k;
    key[0][e] = (e < tot - 1) ? gdl[k] : 0.f;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    int e = lane * EPL + k;
    float gt = (e > 0 ? key[0][e - 1] : 0.f) - key[0][e];
    // scatter through each channel's permutation (B.3): g_in[perm_c(e)] = g_sorted[e]
    int st = idx[0][e];
    if (st >= nc) g_t_f[ray * nf + (st - nc)] = gt;  // t_coarse is a constant
    int ss = idx[4][e];
    if (ss < nc)
      g_sigma_c[ray * nc + ss] = gsg[k];
    else
      g_sigma_f[ray * nf + (ss - nc)] = gsg[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int sc = idx[1 + c][e];
      float g = w[k] * (c == 0 ? g0 : (c == 1 ? g1 : g2));
      if (sc < nc)
        g_rgb_c[(ray * nc + sc) * 3 + c] = g;
      else
        g_rgb_f[(ray * nf + (sc - nc)) * 3 + c] = g;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ray_loss (nerf.py:325-331) + its seed gradients (B.1)
// ---------------------------------------------------------------------------------------------
__global__ void ray_loss_kernel(int64_t n3, const float* __restrict__ cc, const float* __restrict__ cf,
                                const float* __restrict__ ct, float* __restrict__ loss, float* __restrict__ g_cc,
                                float* __restrict__ g_cf) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
    float a = cc[i] - ct[i], b = cf[i] - ct[i];
    acc += a * a + b * b;
    if (g_cc) g_cc[i] = 2.f * a;
    if (g_cf) g_cf[i] = 2.f * b;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && loss) atomicAdd(loss, acc);
}

// ---------------------------------------------------------------------------------------------
#define NT_DISPATCH_EPL(P, CALL)                                                     \
  switch ((P) / 32) {                                                               \
    case 1: { constexpr int E = 1; CALL; } break;                                   \
    case 2: { constexpr int E = 2; CALL; } break;                                   \
    case 3: { constexpr int E = 3; CALL; } break;                                   \
    case 4: { constexpr int E = 4; CALL; } break;                                   \
    case 5: { constexpr int E = 5; CALL; } break;                                   \
    case 6: { constexpr int E = 6; CALL; } break;                                   \
    case 7: { constexpr int E = 7; CALL; } break;                                   \
    default: { constexpr int E = 8; CALL; } break;                                  \
  }

extern "C" int nt_composite_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, const float* rgb,
                                   const float* sigma, float* weights, float* c_out, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && near_ && far_ && rgb && sigma && weights && c_out, "null pointer");
  if (n <= 0) return NT_OK;
  const unsigned blocks = (unsigned)((n + 3) / 4);
  NT_DISPATCH_EPL(ctx->n_coarse, (composite_coarse_kernel<E><<<blocks, 128, 0, (cudaStream_t)stream>>>(n, near_, far_, rgb, sigma,
                                                                                                     weights, c_out)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_get_density(nt_ctx* ctx, int64_t n, int p, const float* delta, const float* sigma, float* weights,
                              void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx, "null ctx");
  NT_REQUIRE(p > 0 && p <= 256 && p % 32 == 0, "get_density: samples per ray must be a multiple of 32, at most 256");
  if (n <= 0) return NT_OK;
  NT_REQUIRE(delta && sigma && weights, "null pointer");
  const unsigned blocks = (unsigned)((n + 3) / 4);
  NT_DISPATCH_EPL(p, (get_density_kernel<E><<<blocks, 128, 0, (cudaStream_t)stream>>>(n, delta, sigma, weights)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_color_cum(nt_ctx* ctx, int64_t n, int p, const float* weights, const float* rgb, float* c_out,
                            void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx, "null ctx");
  NT_REQUIRE(p > 0 && p <= 256 && p % 32 == 0, "color_cum: samples per ray must be a multiple of 32, at most 256");
  if (n <= 0) return NT_OK;
  NT_REQUIRE(weights && rgb && c_out, "null pointer");
  const unsigned blocks = (unsigned)((n + 3) / 4);
  NT_DISPATCH_EPL(p, (color_cum_kernel<E><<<blocks, 128, 0, (cudaStream_t)stream>>>(n, weights, rgb, c_out)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

int nt_launch_composite_coarse_backward(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, const float* rgb,
                                        const float* sigma, const float* g_c, const float* g_w_ext, float* g_rgb,
                                        float* g_sigma, int accumulate, cudaStream_t st) {
  if (n <= 0) return NT_OK;
  const unsigned blocks = (unsigned)((n + 3) / 4);
  NT_DISPATCH_EPL(ctx->n_coarse, (composite_coarse_bwd_kernel<E><<<blocks, 128, 0, st>>>(n, near_, far_, rgb, sigma, g_c, g_w_ext,
                                                                                       g_rgb, g_sigma, accumulate)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_composite_coarse_backward(nt_ctx* ctx, int64_t n, const float* near_, const float* far_,
                                            const float* rgb, const float* sigma, const float* g_c,
                                            const float* g_w_ext, float* g_rgb, float* g_sigma, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && near_ && far_ && rgb && sigma && g_c && g_rgb && g_sigma, "null pointer");
  return nt_launch_composite_coarse_backward(ctx, n, near_, far_, rgb, sigma, g_c, g_w_ext, g_rgb, g_sigma, 0,
                                             (cudaStream_t)stream);
}

extern "C" int nt_composite_fine(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c, const float* sigma_c,
                                 const float* t_f, const float* rgb_f, const float* sigma_f, float last, float* c_out,
                                 float* weights, uint8_t* perm, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_c && rgb_c && sigma_c && t_f && rgb_f && sigma_f && c_out, "null pointer");
  if (n <= 0) return NT_OK;
  // forward: the five channel sorts run in registers (composite_fine_fwd.cu); the shared-memory kernel above is
  // kept for the backward pass, which re-gathers through the stored permutations
  return nt_launch_composite_fine_fwd(ctx, n, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, last, c_out, weights, perm,
                                      (cudaStream_t)stream);
}

extern "C" int nt_composite_fine_backward(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c,
                                          const float* sigma_c, const float* t_f, const float* rgb_f,
                                          const float* sigma_f, float last, const uint8_t* perm, const float* g_c,
                                          float* g_rgb_c, float* g_sigma_c, float* g_rgb_f, float* g_sigma_f,
                                          float* g_t_f, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t_c && rgb_c && sigma_c && t_f && rgb_f && sigma_f && perm && g_c, "null pointer");
  NT_REQUIRE(g_rgb_c && g_sigma_c && g_rgb_f && g_sigma_f && g_t_f, "null output pointer");
  if (n <= 0) return NT_OK;
  const unsigned blocks = (unsigned)((n + FINE_WARPS - 1) / FINE_WARPS);
  NT_DISPATCH_EPL(ctx->n_coarse + ctx->n_fine,
                  (composite_fine_kernel<true, E><<<blocks, FINE_WARPS * 32, 0, (cudaStream_t)stream>>>(
                      n, ctx->n_coarse, ctx->n_fine, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, last, nullptr, nullptr, nullptr,
                      perm, g_c, g_rgb_c, g_sigma_c, g_rgb_f, g_sigma_f, g_t_f)));
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_ray_loss(nt_ctx* ctx, int64_t n, const float* c_coarse, const float* c_fine, const float* c_true,
                           float* loss, float* g_cc, float* g_cf, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && c_coarse && c_fine && c_true, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (loss) NT_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  if (n <= 0) return NT_OK;
  int64_t n3 = n * 3;
  int blocks = (int)((n3 + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  ray_loss_kernel<<<blocks, 256, 0, st>>>(n3, c_coarse, c_fine, c_true, loss, g_cc, g_cf);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
