// Forward of the fine compositing (render_rays nerf.py:302-321) with the five independent channel sorts done in
// REGISTERS: one warp per ray, the nc + nf (192 by default, at most 256) merged samples padded to 256 = 32 lanes x 8 elements, a bitonic network
// whose strides >= 8 are warp shuffles and whose strides < 8 stay inside the lane.  Replaces the shared-memory
// version for the forward pass (the backward re-gathers through the stored permutations and needs no sort).
// Sorting (key, original index) pairs reproduces a stable sort (ATen CPU) and yields the permutations kept for
// backward; keys-only sorting is used when no permutation is requested (render).
#include "common.cuh"

#define FULL 0xffffffffu

namespace {

// ---- order-preserving float <-> uint32 (no negative zeros occur: t > 0, rgb in (0,1), sigma = |.|) ----
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

template <typename T>
__device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T>
__device__ __forceinline__ T tmax(T a, T b) { return a < b ? b : a; }
__device__ __forceinline__ float tmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }

__device__ __forceinline__ float shfl_x(float v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ unsigned long long shfl_x(unsigned long long v, int m) { return __shfl_xor_sync(FULL, v, m); }

// ascending bitonic sort of 256 elements, element e = lane*8 + r, in the direction-free formulation: the first round of
// merge stage k pairs e with its mirror e ^ (k-1) inside the k-block (two ascending halves -> bitonic split), the
// remaining rounds pair e with e ^ j; every compare-exchange is ascending (the lower index keeps the minimum).  Pairs
// inside a lane are compile-time (two FMNMX, no select); pairs across lanes cost one shuffle and one min-or-max chosen
// by a lane predicate.
template <typename T>
__device__ __forceinline__ void warp_sort_256(T (&v)[8], int lane) {
#pragma unroll
  for (int k = 2; k <= 256; k <<= 1) {
    // ---- mirror round: partner e ^ (k-1)
    if (k <= 8) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int q = r ^ (k - 1);
        if (r < q) {
          const T a = v[r], b = v[q];
          v[r] = tmin(a, b);
          v[q] = tmax(a, b);
        }
      }
    } else {
      const int m = (k >> 3) - 1;                       // partner lane = lane ^ m, partner register = 7 - r
      const bool lower = (lane & (k >> 4)) == 0;
      T o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] = shfl_x(v[7 - r], m);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = lower ? tmin(v[r], o[r]) : tmax(v[r], o[r]);
    }
    // ---- half-cleaner rounds: partner e ^ j, j = k/4 .. 1
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      if (j >= 8) {
        const int m = j >> 3;
        const bool lower = (lane & m) == 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const T o = shfl_x(v[r], m);
          v[r] = lower ? tmin(v[r], o) : tmax(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if ((r & j) == 0) {
            const T a = v[r], b = v[r | j];
            v[r] = tmin(a, b);
            v[r | j] = tmax(a, b);
          }
        }
      }
    }
  }
}

// The float instantiation is inlined five times (keys stay in registers: as an out-of-line call the array lived in local
// memory and the kernel was bound by the LSU pipe - 445 local stores per ray next to its 600 shuffles; inlining took the
// 160000-ray render launch from 1.17 to ~0.6 ms).  The 64-bit (key, index) instantiation of the training path is
// register-hungry (95 registers inlined) and is faster out of line.
__device__ __noinline__ void warp_sort_256_call(unsigned long long (&v)[8], int lane) { warp_sort_256<unsigned long long>(v, lane); }

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

// loads this lane's 8 merged elements (coarse samples 0..nc-1 | fine nc..nc+nf-1 | +inf padding) of every channel;
// nc and nf are multiples of 32, so a lane's 8 elements never straddle the coarse / fine boundary
__device__ __forceinline__ void load_lane(int64_t ray, int lane, int nc, int nf, const float* __restrict__ t_c, const float* __restrict__ rgb_c,
                                          const float* __restrict__ sigma_c, const float* __restrict__ t_f,
                                          const float* __restrict__ rgb_f, const float* __restrict__ sigma_f,
                                          float (&ch)[5][8]) {
  const float inf = __int_as_float(0x7f800000);
  if (lane * 8 >= nc + nf) {
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
      for (int r = 0; r < 8; ++r) ch[c][r] = inf;
    return;
  }
  const bool co = lane * 8 < nc;
  const int64_t s0 = co ? ray * nc + lane * 8 : ray * nf + (lane * 8 - nc);  // first sample of this lane
  const float4* pt = reinterpret_cast<const float4*>((co ? t_c : t_f) + s0);
  const float4* ps = reinterpret_cast<const float4*>((co ? sigma_c : sigma_f) + s0);
  const float4* pc = reinterpret_cast<const float4*>((co ? rgb_c : rgb_f) + s0 * 3);
  const float4 t0 = __ldg(pt), t1 = __ldg(pt + 1), g0 = __ldg(ps), g1 = __ldg(ps + 1);
  ch[0][0] = t0.x; ch[0][1] = t0.y; ch[0][2] = t0.z; ch[0][3] = t0.w;
  ch[0][4] = t1.x; ch[0][5] = t1.y; ch[0][6] = t1.z; ch[0][7] = t1.w;
  ch[4][0] = g0.x; ch[4][1] = g0.y; ch[4][2] = g0.z; ch[4][3] = g0.w;
  ch[4][4] = g1.x; ch[4][5] = g1.y; ch[4][6] = g1.z; ch[4][7] = g1.w;
  float c[24];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const float4 x = __ldg(pc + q);
    c[4 * q] = x.x; c[4 * q + 1] = x.y; c[4 * q + 2] = x.z; c[4 * q + 3] = x.w;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    ch[1][r] = c[3 * r];
    ch[2][r] = c[3 * r + 1];
    ch[3][r] = c[3 * r + 2];
  }
}

// compositing over the sorted channels of one ray (one warp): delta = diff(t) ++ [last] (nerf.py:315), inclusive fp64 prefix
__device__ __forceinline__ void composite_sorted(const float (&ch)[5][8], int64_t ray, int lane, int tot, float last,
                                                 float* __restrict__ c_out, float* __restrict__ weights) {
  const bool live = lane * 8 < tot;
  const float t_next_lane = __shfl_down_sync(FULL, ch[0][0], 1);
  float a[8], dl[8];
  double run = 0.0, pre[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int e = lane * 8 + r;
    const float tn = r < 7 ? ch[0][r + 1] : t_next_lane;
    dl[r] = e < tot - 1 ? __fsub_rn(tn, ch[0][r]) : last;
    a[r] = e < tot ? __fmul_rn(dl[r], ch[4][r]) : 0.f;
    run += (double)a[r];
    pre[r] = run;
  }
  const double base = warp_incl_scan(run, lane) - run;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, w[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float S = (float)(base + pre[r]);
    w[r] = live ? expf(-S) * (1.f - expf(-a[r])) : 0.f;
    if (live) {
      c0 += w[r] * ch[1][r];
      c1 += w[r] * ch[2][r];
      c2 += w[r] * ch[3][r];
    }
  }
  if (weights && live) {
    float4* pw = reinterpret_cast<float4*>(weights + ray * tot + lane * 8);
    pw[0] = make_float4(w[0], w[1], w[2], w[3]);
    pw[1] = make_float4(w[4], w[5], w[6], w[7]);
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

template <bool PERM>
__global__ void __launch_bounds__(128) composite_fine_fwd_kernel(int64_t n, int nc, int nf, const float* __restrict__ t_c,
                                                                 const float* __restrict__ rgb_c,
                                                                 const float* __restrict__ sigma_c,
                                                                 const float* __restrict__ t_f,
                                                                 const float* __restrict__ rgb_f,
                                                                 const float* __restrict__ sigma_f, float last,
                                                                 float* __restrict__ c_out, float* __restrict__ weights,
                                                                 uint8_t* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  const int tot = nc + nf;
  const bool live = lane * 8 < tot;
  float ch[5][8];
  load_lane(ray, lane, nc, nf, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, ch);
  (void)live;
  if (PERM) {
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      unsigned long long kv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) kv[r] = ((unsigned long long)f2ord(ch[c][r]) << 32) | (unsigned)(lane * 8 + r);
      warp_sort_256_call(kv, lane);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        ch[c][r] = ord2f((uint32_t)(kv[r] >> 32));
        const uint32_t id = (uint32_t)kv[r] & 0xffu;
        if (r < 4) lo |= id << (8 * r); else hi |= id << (8 * (r - 4));
      }
      if (live) *reinterpret_cast<uint2*>(perm + (ray * 5 + c) * tot + lane * 8) = make_uint2(lo, hi);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 5; ++c) warp_sort_256<float>(ch[c], lane);
  }
  composite_sorted(ch, ray, lane, tot, last, c_out, weights);
}

// Training variant: FIVE warps per ray, one per channel.  The (key, index) sorts that produce the permutations are five
// independent serial shuffle chains; on one warp they made the launch latency-bound (45 us for ANY batch up to ~2 000 rays,
// 7 warps per SM at the 1024-ray training batch).  Each warp sorts its own channel, parks the sorted values in shared
// memory, and warp 0 composites.
__global__ void __launch_bounds__(160) composite_fine_fwd5_kernel(int64_t n, int nc, int nf, const float* __restrict__ t_c,
                                                                  const float* __restrict__ rgb_c,
                                                                  const float* __restrict__ sigma_c,
                                                                  const float* __restrict__ t_f,
                                                                  const float* __restrict__ rgb_f,
                                                                  const float* __restrict__ sigma_f, float last,
                                                                  float* __restrict__ c_out, float* __restrict__ weights,
                                                                  uint8_t* __restrict__ perm) {
  __shared__ float s_ch[5][256];
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int64_t ray = blockIdx.x;
  const int tot = nc + nf;
  const bool live = lane * 8 < tot;
  float v[8];
  if (!live) {
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = __int_as_float(0x7f800000);
  } else {
    const bool co = lane * 8 < nc;
    const int64_t s0 = co ? ray * nc + lane * 8 : ray * nf + (lane * 8 - nc);
    if (c == 0 || c == 4) {
      const float4* p = reinterpret_cast<const float4*>((c == 0 ? (co ? t_c : t_f) : (co ? sigma_c : sigma_f)) + s0);
      const float4 x0 = __ldg(p), x1 = __ldg(p + 1);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
      v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    } else {
      const float* p = (co ? rgb_c : rgb_f) + s0 * 3 + (c - 1);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = __ldg(p + 3 * r);
    }
  }
  unsigned long long kv[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) kv[r] = ((unsigned long long)f2ord(v[r]) << 32) | (unsigned)(lane * 8 + r);
  warp_sort_256<unsigned long long>(kv, lane);
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    s_ch[c][lane * 8 + r] = ord2f((uint32_t)(kv[r] >> 32));
    const uint32_t id = (uint32_t)kv[r] & 0xffu;
    if (r < 4) lo |= id << (8 * r); else hi |= id << (8 * (r - 4));
  }
  if (live) *reinterpret_cast<uint2*>(perm + (ray * 5 + c) * tot + lane * 8) = make_uint2(lo, hi);
  __syncthreads();
  if (c != 0) return;
  float ch[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float4 x0 = *reinterpret_cast<const float4*>(&s_ch[k][lane * 8]), x1 = *reinterpret_cast<const float4*>(&s_ch[k][lane * 8 + 4]);
    ch[k][0] = x0.x; ch[k][1] = x0.y; ch[k][2] = x0.z; ch[k][3] = x0.w;
    ch[k][4] = x1.x; ch[k][5] = x1.y; ch[k][6] = x1.z; ch[k][7] = x1.w;
  }
  composite_sorted(ch, ray, lane, tot, last, c_out, weights);
}

}  // namespace

int nt_launch_composite_fine_fwd(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c, const float* sigma_c,
                                 const float* t_f, const float* rgb_f, const float* sigma_f, float last, float* c_out,
                                 float* weights, uint8_t* perm, cudaStream_t st) {
  const unsigned blocks = (unsigned)((n + 3) / 4);
  if (perm)
    composite_fine_fwd5_kernel<<<(unsigned)n, 160, 0, st>>>(n, ctx->n_coarse, ctx->n_fine, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, last, c_out,
                                                            weights, perm);
  else
    composite_fine_fwd_kernel<false><<<blocks, 128, 0, st>>>(n, ctx->n_coarse, ctx->n_fine, t_c, rgb_c, sigma_c, t_f, rgb_f, sigma_f, last, c_out,
                                                             weights, perm);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
