// Ray generation, coarse sample placement and the positional encoder.
// Reference: nerf.py:186-216 (net_out geometry), nerf.py:288 (np.linspace), nerf.py:135-167 (Encoder).
// The bit-exact targets of SURVEY.md Appendix A.1/A.2 are computed with explicit round-to-nearest
// intrinsics (__fmul_rn/__fadd_rn never contract into FMA), IEEE sqrt and division.
#include "common.cuh"

__constant__ uint32_t c_freq_point[10] = NT_FREQ_POINT_INIT;
__constant__ uint32_t c_freq_dir[4] = NT_FREQ_DIR_INIT;

// ---------------------------------------------------------------------------------------------
// raygen: one thread per ray (nerf.py:186-197, :211) + view-direction features (nerf.py:135-167)
// ---------------------------------------------------------------------------------------------
__global__ void raygen_kernel(int64_t n, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                              const float* __restrict__ c2w, int c2w_stride, const float* __restrict__ kinv,
                              float* __restrict__ rays, float* __restrict__ dir_wrd, float* __restrict__ dir_enc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = (float)row[i];  // nerf.py:343: the loader's row index is the horizontal coordinate
  const float y = (float)col[i];
  float v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)  // (N,1,3)x(3,3): ((x*k0j + y*k1j) + 1*k2j), each op rounded
    v[j] = __fadd_rn(__fadd_rn(__fmul_rn(x, kinv[j]), __fmul_rn(y, kinv[3 + j])), kinv[6 + j]);
  float n2 = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
  float nrm = fmaxf(__fsqrt_rn(n2), 1e-12f);  // functional.normalize eps (nerf.py:197)
  float d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) d[j] = __fdiv_rn(v[j], nrm);
  const float* m = c2w + i * c2w_stride;
  const int rs = c2w_stride == 17 ? 5 : 4;  // 17 = the loader's pose row [3x5 | near far] (loader.py:33)
  float* r = rays + i * 16;
  r[0] = d[0];
  r[1] = d[1];
  r[2] = d[2];
  float dw[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float r0 = m[a * rs + 0], r1 = m[a * rs + 1], r2 = m[a * rs + 2];
    r[3 + a * 3 + 0] = r0;
    r[3 + a * 3 + 1] = r1;
    r[3 + a * 3 + 2] = r2;
    r[12 + a] = m[a * rs + 3];
    dw[a] = __fadd_rn(__fadd_rn(__fmul_rn(r0, d[0]), __fmul_rn(r1, d[1])), __fmul_rn(r2, d[2]));  // nerf.py:211
  }
  r[15] = 0.f;
  if (dir_wrd) {
    dir_wrd[i * 3 + 0] = dw[0];
    dir_wrd[i * 3 + 1] = dw[1];
    dir_wrd[i * 3 + 2] = dw[2];
  }
  if (dir_enc) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        float s, co;
        sincosf(__fmul_rn(__uint_as_float(c_freq_dir[l]), dw[c]), &s, &co);
        dir_enc[i * 24 + c * 8 + 2 * l + 0] = s;
        dir_enc[i * 24 + c * 8 + 2 * l + 1] = co;
      }
  }
}

// ---------------------------------------------------------------------------------------------
// t_coarse: np.linspace(near, far, Nc) in fp32 (nerf.py:288)
// ---------------------------------------------------------------------------------------------
__global__ void step_zero_flag_kernel(int64_t n, const float* __restrict__ near_, const float* __restrict__ far_,
                                      int nc, int* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int z = 0;
  if (i < n) z = (__fdiv_rn(__fsub_rn(far_[i], near_[i]), (float)(nc - 1)) == 0.f);
  z = __any_sync(0xffffffffu, z);
  if (z && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Batch-global quantities of a ray-sharded launch: out[0] = 1 if any LOCAL ray has a zero linspace step
// (max-accumulated: the caller zeroes out[0] first), out[1] / out[2] = t_coarse[0,1] - t_coarse[0,0] of this shard's ray 0
// as np.linspace produces it when the global flag is 0 / 1 (nerf.py:234, :288), or -inf when this is not the first shard.
// An element-wise MAX over the ranks' vectors is then the global triple.
__global__ void shard_globals_local_kernel(int64_t n, const float* __restrict__ near_, const float* __restrict__ far_,
                                           int nc, int first_shard, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int z = 0;
  if (i < n) z = (__fdiv_rn(__fsub_rn(far_[i], near_[i]), (float)(nc - 1)) == 0.f);
  z = __any_sync(0xffffffffu, z);
  if (z && (threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(1.0f));  // 0.0f / 1.0f order as ints
  if (i == 0) {
    float d0 = -INFINITY, d1 = -INFINITY;
    if (first_shard) {
      const float a = near_[0], b = far_[0], div = (float)(nc - 1), delta = __fsub_rn(b, a);
      const float t1a = nc == 2 ? b : __fadd_rn(__fmul_rn(1.f, __fdiv_rn(delta, div)), a);
      const float t1b = nc == 2 ? b : __fadd_rn(__fmul_rn(__fdiv_rn(1.f, div), delta), a);
      d0 = __fsub_rn(t1a, a);  // t_coarse[0,0] == near in both branches
      d1 = __fsub_rn(t1b, a);
    }
    out[1] = d0;
    out[2] = d1;
    out[3] = 0.f;
  }
}
__global__ void shard_globals_resolve_kernel(float* __restrict__ g) {
  const float flag = g[0] != 0.f ? 1.f : 0.f;
  const float d = flag != 0.f ? g[2] : g[1];
  g[0] = d;     // delta0, where nt_sample_pdf / nt_render_* read it
  g[1] = flag;  // any_step_zero, read by the coarse sampler when any_step_zero == NT_ANY_STEP_ZERO_DEVICE
}

__global__ void sample_coarse_kernel(int64_t n, const float* __restrict__ near_, const float* __restrict__ far_,
                                     int nc, int forced, const int* __restrict__ flag, const float* __restrict__ flag_f,
                                     float* __restrict__ t) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * nc) return;
  int64_t r = gid / nc;
  int i = (int)(gid - r * nc);
  const float a = near_[r], b = far_[r];
  const float div = (float)(nc - 1);
  const float delta = __fsub_rn(b, a);
  const int zero = forced >= 0 ? forced : (flag_f ? (*flag_f != 0.f) : *flag);
  float y;
  if (zero)
    y = __fadd_rn(__fmul_rn(__fdiv_rn((float)i, div), delta), a);
  else
    y = __fadd_rn(__fmul_rn((float)i, __fdiv_rn(delta, div)), a);
  if (i == nc - 1 && nc > 1) y = b;
  t[gid] = y;
}

// ---------------------------------------------------------------------------------------------
// positional encoding of sample positions into a dense [S, ld] fp32 matrix (layer-major fp32 path)
// p_cam = fl(d_cam*t); p_wrd = ((R0*p0 + R1*p1) + R2*p2) + T  (nerf.py:200-216, no FMA)
// feature index c*20 + 2l + s (nerf.py:103-104, 159)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sample_position(const float* __restrict__ ray, float t, float p[3]) {
  float pc0 = __fmul_rn(ray[0], t), pc1 = __fmul_rn(ray[1], t), pc2 = __fmul_rn(ray[2], t);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    p[a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ray[3 + a * 3], pc0), __fmul_rn(ray[4 + a * 3], pc1)),
                               __fmul_rn(ray[5 + a * 3], pc2)),
                     ray[12 + a]);
}

__global__ void encode_points_kernel(int64_t total, int p, const float* __restrict__ t, const float* __restrict__ rays,
                                     float* __restrict__ enc, int ld) {
  // thread = (sample, channel c): 10 sincos each
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total * 3) return;
  int64_t s = gid / 3;
  int c = (int)(gid - s * 3);
  const float* ray = rays + (s / p) * 16;
  float pos[3];
  sample_position(ray, t[s], pos);
  const float x = c == 0 ? pos[0] : (c == 1 ? pos[1] : pos[2]);
  float* o = enc + s * ld + c * 20;
#pragma unroll
  for (int l = 0; l < 10; ++l) {
    float sn, cs;
    sincosf(__fmul_rn(__uint_as_float(c_freq_point[l]), x), &sn, &cs);
    o[2 * l] = sn;
    o[2 * l + 1] = cs;
  }
}

// Encoder.forward (nerf.py:135-167) on caller-supplied coordinates: x [S,3] -> out [S,3,2L], feature (c, 2l+s) =
// sin / cos of fl(w_l * x_c); thread = (sample, channel)
template <int L>
__global__ void encode_generic_kernel(int64_t total, const float* __restrict__ x, const uint32_t* __restrict__ freq_unused,
                                      float* __restrict__ out) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total * 3) return;
  const float v = x[gid];
  float* o = out + gid * (2 * L);
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float w = __uint_as_float(L == 10 ? c_freq_point[l] : c_freq_dir[l]);
    float sn, cs;
    sincosf(__fmul_rn(w, v), &sn, &cs);
    o[2 * l] = sn;
    o[2 * l + 1] = cs;
  }
}

__global__ void expand_dir_enc_kernel(int64_t total, int p, const float* __restrict__ dir_enc, float* __restrict__ out,
                                      int ld) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total * 24) return;
  int64_t s = gid / 24;
  int f = (int)(gid - s * 24);
  out[s * ld + f] = dir_enc[(s / p) * 24 + f];
}

// g_enc [S,ld] (60 used) -> g_t [S]  (SURVEY.md B.4): g_p_c = sum_l w_l (g_sin*cos - g_cos*sin); g_t = g_p . d_wrd
// FAST (bf16 training path, ld = 64): the row is read as 15 float4 and sin/cos come from the SFU after the same
// two-constant range reduction the fused forward kernel uses (abs error 5e-7, far inside the bf16 operand rounding);
// the fp32 path keeps sincosf.
__device__ __forceinline__ void fast_sincos_rr(float x, float& s, float& c) {
  const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.f), -12582912.f);
  float r = fmaf(k, -6.2831854820251465f, x);
  r = fmaf(k, 1.7484555314695172e-07f, r);
  s = __sinf(r);
  c = __cosf(r);
}
template <bool FAST>
__global__ void encode_backward_kernel(int64_t total, int p, const float* __restrict__ t, const float* __restrict__ rays,
                                       const float* __restrict__ g_enc, int ld, float* __restrict__ g_t) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const float* ray = rays + (s / p) * 16;
  float pos[3];
  sample_position(ray, t[s], pos);
  float gl[60];
  if (FAST) {
    const float4* g4 = reinterpret_cast<const float4*>(g_enc + s * ld);
#pragma unroll
    for (int q = 0; q < 15; ++q) {
      const float4 v = __ldg(g4 + q);
      gl[4 * q] = v.x; gl[4 * q + 1] = v.y; gl[4 * q + 2] = v.z; gl[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 60; ++q) gl[q] = g_enc[s * ld + q];
  }
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // d_wrd_c = (R_c0*d0 + R_c1*d1) + R_c2*d2 (nerf.py:211)
    float dw = __fadd_rn(__fadd_rn(__fmul_rn(ray[3 + c * 3], ray[0]), __fmul_rn(ray[4 + c * 3], ray[1])),
                         __fmul_rn(ray[5 + c * 3], ray[2]));
    float gp = 0.f;
#pragma unroll
    for (int l = 0; l < 10; ++l) {
      float w = __uint_as_float(c_freq_point[l]);
      float sn, cs;
      if (FAST)
        fast_sincos_rr(__fmul_rn(w, pos[c]), sn, cs);
      else
        sincosf(__fmul_rn(w, pos[c]), &sn, &cs);
      gp += w * (gl[c * 20 + 2 * l] * cs - gl[c * 20 + 2 * l + 1] * sn);
    }
    acc += gp * dw;
  }
  g_t[s] = acc;
}

// ---------------------------------------------------------------------------------------------
// C-ABI + internal launchers
// ---------------------------------------------------------------------------------------------
extern "C" int nt_raygen(nt_ctx* ctx, int64_t n, const int64_t* row, const int64_t* col, const float* c2w,
                         int c2w_stride, const float* kinv, float* rays, float* dir_wrd, float* dir_enc, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && row && col && c2w && kinv && rays, "null pointer");
  NT_REQUIRE(c2w_stride == 16 || c2w_stride == 12 || c2w_stride == 17, "c2w_stride must be 12, 16 or 17");
  if (n <= 0) return NT_OK;
  raygen_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, row, col, c2w, c2w_stride, kinv, rays,
                                                                                dir_wrd, dir_enc);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

int nt_launch_sample_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int any_step_zero,
                            const float* flag_dev, float* t_coarse, cudaStream_t st) {
  const int nc = ctx->n_coarse;
  if (any_step_zero == NT_ANY_STEP_ZERO_DEVICE) {
    NT_REQUIRE(flag_dev, "NT_ANY_STEP_ZERO_DEVICE needs the device-resident globals (nt_shard_globals_resolve)");
  } else {
    flag_dev = nullptr;
    if (any_step_zero < 0) {
      NT_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), st));
      step_zero_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, near_, far_, nc, ctx->d_flags);
      NT_LAUNCH_CHECK(ctx);
    }
  }
  int64_t total = n * nc;
  sample_coarse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n, near_, far_, nc, flag_dev ? -1 : any_step_zero,
                                                                       ctx->d_flags, flag_dev, t_coarse);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_sample_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int any_step_zero,
                                float* t_coarse, void* stream) {
  NT_ENTER(ctx);
  if (n <= 0) return NT_OK;
  NT_REQUIRE(ctx && near_ && far_ && t_coarse, "null pointer");
  NT_REQUIRE(any_step_zero >= -1 && any_step_zero <= 1, "any_step_zero must be -1, 0 or 1 here");
  return nt_launch_sample_coarse(ctx, n, near_, far_, any_step_zero, nullptr, t_coarse, (cudaStream_t)stream);
}

extern "C" int nt_shard_globals_local(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int first_shard,
                                      float* out4, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && out4 && (n == 0 || (near_ && far_)), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  NT_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(float), st));
  // an empty shard (more ranks than rays) still contributes the neutral element
  const int64_t threads = n > 0 ? n : 1;
  shard_globals_local_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, near_, far_, ctx->n_coarse,
                                                                              (first_shard && n > 0) ? 1 : 0, out4);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_shard_globals_resolve(nt_ctx* ctx, float* g4, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && g4, "null pointer");
  shard_globals_resolve_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g4);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

int nt_launch_encode_points(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, float* enc, int ld_enc,
                            cudaStream_t st) {
  int64_t total = n * p;
  if (total <= 0) return NT_OK;
  encode_points_kernel<<<(unsigned)((total * 3 + 255) / 256), 256, 0, st>>>(total, p, t, rays, enc, ld_enc);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

extern "C" int nt_encode(nt_ctx* ctx, int64_t total, const float* points, const float* dirs, float* gamma_point,
                         float* gamma_dir, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx, "null ctx");
  if (total <= 0) return NT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((total * 3 + 255) / 256);
  if (points) {
    NT_REQUIRE(gamma_point, "null output pointer");
    encode_generic_kernel<10><<<blocks, 256, 0, st>>>(total, points, nullptr, gamma_point);
    NT_LAUNCH_CHECK(ctx);
  }
  if (dirs) {
    NT_REQUIRE(gamma_dir, "null output pointer");
    encode_generic_kernel<4><<<blocks, 256, 0, st>>>(total, dirs, nullptr, gamma_dir);
    NT_LAUNCH_CHECK(ctx);
  }
  return NT_OK;
}

int nt_launch_expand_dir_enc(nt_ctx* ctx, int64_t n, int p, const float* dir_enc, float* out, int ld, cudaStream_t st) {
  int64_t total = n * p;
  if (total <= 0) return NT_OK;
  expand_dir_enc_kernel<<<(unsigned)((total * 24 + 255) / 256), 256, 0, st>>>(total, p, dir_enc, out, ld);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

int nt_launch_encode_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* g_enc,
                              int ld, float* g_t, cudaStream_t st) {
  int64_t total = n * p;
  if (total <= 0) return NT_OK;
  if (ld == 64 && (reinterpret_cast<uintptr_t>(g_enc) & 15) == 0)   // bf16 training path: padded, 16-byte aligned rows
    encode_backward_kernel<true><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(total, p, t, rays, g_enc, ld, g_t);
  else
    encode_backward_kernel<false><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(total, p, t, rays, g_enc, ld, g_t);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
