// Fused Adam over the flat parameter buffer — one launch instead of torch.optim.Adam's per-tensor
// sequence (nerf.py:425, :474).  Math follows torch/optim/adam.py single-tensor path with
// weight_decay=0, amsgrad=False:  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM-bound: 28 B per parameter.
#include "common.cuh"
#include <math.h>

__global__ void adam_kernel(int64_t count, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float step_size, float inv_sqrt_bc2, float b1, float b2, float eps,
                            float grad_scale) {
  int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < count) {
    float4 pp = *(float4*)(p + i4), gg = *(const float4*)(g + i4), mm = *(float4*)(m + i4), vv = *(float4*)(v + i4);
    float* pa = (float*)&pp;
    float* ga = (float*)&gg;
    float* ma = (float*)&mm;
    float* va = (float*)&vv;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = ga[k] * grad_scale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gr;
      va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
      float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    *(float4*)(p + i4) = pp;
    *(float4*)(m + i4) = mm;
    *(float4*)(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < count; ++i) {
      float gr = g[i] * grad_scale;
      float mi = b1 * m[i] + (1.f - b1) * gr;
      float vi = b2 * v[i] + (1.f - b2) * gr * gr;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
  }
}

extern "C" int nt_adam_step(nt_ctx* ctx, int64_t count, float* params, const float* grads, float* m, float* v, float lr,
                            float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && params && grads && m && v, "null pointer");
  NT_REQUIRE(step >= 1, "step is 1-based");
  NT_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "buffers must be 16B aligned");
  if (count <= 0) return NT_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int64_t threads = (count + 3) / 4;
  adam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, params, grads, m, v, step_size,
                                                                                  inv_sqrt_bc2, beta1, beta2, eps, grad_scale);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused gradient all-reduce + Adam over NVLink peer memory: every rank keeps its flat gradient in a symmetric
// (peer-mapped) buffer; this ONE kernel reads all ranks' gradients with plain loads over NVLink 5 / NVSwitch, sums them
// in rank order (so every rank computes bit-identical parameters) and applies the Adam update — the 2.4 MB SUM
// all-reduce of SURVEY.md §8(e) never exists as a separate collective.  The caller brackets it with cross-rank barriers
// (all gradients written before; all reads done after).
// ---------------------------------------------------------------------------------------------------------------------
struct PeerGrads {
  const float* g[8];
  int n;
};

__global__ void adam_allreduce_kernel(int64_t count, float* __restrict__ p, PeerGrads pg, float* __restrict__ m,
                                      float* __restrict__ v, float step_size, float inv_sqrt_bc2, float b1, float b2,
                                      float eps, float grad_scale, float* __restrict__ g_sum_out) {
  int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < count) {
    float4 gg = *reinterpret_cast<const float4*>(pg.g[0] + i4);
    for (int r = 1; r < pg.n; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(pg.g[r] + i4);
      gg.x += x.x;
      gg.y += x.y;
      gg.z += x.z;
      gg.w += x.w;
    }
    float4 pp = *(float4*)(p + i4), mm = *(float4*)(m + i4), vv = *(float4*)(v + i4);
    float* pa = (float*)&pp;
    float* ga = (float*)&gg;
    float* ma = (float*)&mm;
    float* va = (float*)&vv;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * grad_scale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gr;
      va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
      pa[k] -= step_size * (ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps));
    }
    *(float4*)(p + i4) = pp;
    *(float4*)(m + i4) = mm;
    *(float4*)(v + i4) = vv;
    if (g_sum_out) *(float4*)(g_sum_out + i4) = gg;
  } else {
    for (int64_t i = i4; i < count; ++i) {
      float gr = 0.f;
      for (int r = 0; r < pg.n; ++r) gr += pg.g[r][i];
      if (g_sum_out) g_sum_out[i] = gr;
      gr *= grad_scale;
      const float mi = b1 * m[i] + (1.f - b1) * gr;
      const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
  }
}

extern "C" int nt_adam_step_allreduce(nt_ctx* ctx, int64_t count, float* params, const float* const* rank_grads,
                                      int n_ranks, float* m, float* v, float lr, float beta1, float beta2, float eps,
                                      int64_t step, float grad_scale, float* grad_sum_out, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && params && rank_grads && m && v, "null pointer");
  NT_REQUIRE(n_ranks >= 1 && n_ranks <= 8, "1..8 ranks");
  NT_REQUIRE(step >= 1, "step is 1-based");
  PeerGrads pg;
  pg.n = n_ranks;
  for (int r = 0; r < 8; ++r) pg.g[r] = r < n_ranks ? rank_grads[r] : nullptr;
  for (int r = 0; r < n_ranks; ++r) NT_REQUIRE(pg.g[r] && (((uintptr_t)pg.g[r]) & 15) == 0, "gradient buffers must be 16B aligned");
  NT_REQUIRE((((uintptr_t)params | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad_sum_out) & 15) == 0, "buffers must be 16B aligned");
  if (count <= 0) return NT_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int64_t threads = (count + 3) / 4;
  adam_allreduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      count, params, pg, m, v, step_size, inv_sqrt_bc2, beta1, beta2, eps, grad_scale, grad_sum_out);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
