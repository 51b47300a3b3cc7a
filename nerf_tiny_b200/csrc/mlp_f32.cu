// NT_PREC_FP32 encode+MLP: layer-major fp32 forward and backward built from gemm_f32.cu.
// Reference: net_out nerf.py:200-219 -> Encoder.forward nerf.py:135-167 -> Network.forward nerf.py:101-124,
// and the autograd of those (nerf.py:473).
// This is the accuracy path (rgb within 1e-3 of the reference, in practice ~1e-6); the throughput path is
// mlp_tc.cu (tcgen05).  Activations live in HBM as [S, features] fp32, S = N*P samples.
#include "common.cuh"

struct F32Ws {
  // forward stash (train) / scratch (render)
  float *enc, *denc, *h[8], *zsig, *info, *u, *rgb;
  // backward scratch
  float *gA, *gB, *g_enc, *g_u, *g_zsig, *g_zrgb;
  size_t bytes;
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static F32Ws carve_f32(void* base, int64_t S, bool train) {
  F32Ws w;
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t floats) {
    float* p = (float*)(b + off);
    off += align256(floats * sizeof(float));
    return p;
  };
  w.enc = take(S * 60);
  w.denc = take(S * 24);
  if (train) {
    for (int i = 0; i < 8; ++i) w.h[i] = take(S * 256);
    w.zsig = take(S);
    w.info = take(S * 256);
    w.u = take(S * 128);
    w.rgb = take(S * 3);
    w.gA = take(S * 256);
    w.gB = take(S * 256);
    w.g_enc = take(S * 60);
    w.g_u = take(S * 128);
    w.g_zsig = take(S);
    w.g_zrgb = take(S * 3);
  } else {
    float* a = take(S * 256);
    float* c = take(S * 256);
    for (int i = 0; i < 8; ++i) w.h[i] = (i & 1) ? c : a;
    w.zsig = nullptr;
    w.info = a;  // h7 lives in c (i=7 odd), info overwrites a
    w.u = take(S * 128);
    w.rgb = nullptr;
    w.gA = w.gB = w.g_enc = w.g_u = w.g_zsig = w.g_zrgb = nullptr;
  }
  w.bytes = off;
  return w;
}

#define F32_RENDER_CHUNK_SAMPLES (1 << 19)

size_t nt_mlp_f32_workspace_bytes(int64_t n, int p, int train) {
  int64_t S = n * p;
  if (!train && S > F32_RENDER_CHUNK_SAMPLES) S = (F32_RENDER_CHUNK_SAMPLES / p) * (int64_t)p;
  return carve_f32(nullptr, S, train != 0).bytes;
}

static GemmEpi epi_of(const float* bias, int act) {
  GemmEpi e;
  e.bias = bias;
  e.act = act;
  e.mask = nullptr;
  e.ldmask = 0;
  e.accumulate = 0;
  e.atomic = 0;
  e.pre_out = nullptr;
  e.ldpre = 0;
  return e;
}
static GemmSeg seg_of(const float* A, int lda, const float* B, int ldb, int K) {
  GemmSeg s;
  s.A = A;
  s.lda = lda;
  s.B = B;
  s.ldb = ldb;
  s.K = K;
  return s;
}
static const GemmSeg kNoSeg = {nullptr, 0, nullptr, 0, 0};

#define NT_TRY(x)            \
  do {                       \
    int rc__ = (x);          \
    if (rc__ != NT_OK) return rc__; \
  } while (0)

static int forward_from_enc(nt_ctx* ctx, int S, const float* P, float* rgb, float* sigma, F32Ws& w, bool train,
                            cudaStream_t st);

static int forward_chunk(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                         const float* P, float* rgb, float* sigma, F32Ws& w, bool train, cudaStream_t st) {
  NT_TRY(nt_launch_encode_points(ctx, n, p, t, rays, w.enc, 60, st));
  NT_TRY(nt_launch_expand_dir_enc(ctx, n, p, dir_enc, w.denc, 24, st));
  return forward_from_enc(ctx, (int)(n * p), P, rgb, sigma, w, train, st);
}

// Network.forward (nerf.py:101-124) on the dense feature matrices w.enc [S,60] / w.denc [S,24]
static int forward_from_enc(nt_ctx* ctx, int S, const float* P, float* rgb, float* sigma, F32Ws& w, bool train,
                            cudaStream_t st) {
  const LayerTable L = nt_layers();
  // L0 (nerf.py:85, :111)
  NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.enc, 60, P + L.w[L_P0], 60, 60), kNoSeg, false, false, w.h[0], 256,
                        epi_of(P + L.b[L_P0], ACT_RELU), 1, st));
  for (int i = 1; i < 8; ++i) {
    if (i == 4)  // skip: cat((hidden, enc)) (nerf.py:109) as two K-segments
      NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.h[3], 256, P + L.w[L_P4], 316, 256),
                            seg_of(w.enc, 60, P + L.w[L_P4] + 256, 316, 60), false, false, w.h[4], 256,
                            epi_of(P + L.b[L_P4], ACT_RELU), 1, st));
    else
      NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.h[i - 1], 256, P + L.w[i], 256, 256), kNoSeg, false, false, w.h[i],
                            256, epi_of(P + L.b[i], ACT_RELU), 1, st));
  }
  // sigma = |w.h7 + b| (nerf.py:74, :114)
  GemmEpi es = epi_of(P + L.b[L_SIGMA], ACT_ABS);
  if (train) {
    es.pre_out = w.zsig;
    es.ldpre = 1;
  }
  NT_TRY(nt_launch_gemm(ctx, S, 1, seg_of(w.h[7], 256, P + L.w[L_SIGMA], 256, 256), kNoSeg, false, false, sigma, 1, es,
                        1, st));
  // point_info: linear, no activation (nerf.py:117)
  NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.h[7], 256, P + L.w[L_INFO], 256, 256), kNoSeg, false, false, w.info, 256,
                        epi_of(P + L.b[L_INFO], ACT_NONE), 1, st));
  // dir_info on cat((dir_enc, point_info)) (nerf.py:118)
  NT_TRY(nt_launch_gemm(ctx, S, 128, seg_of(w.denc, 24, P + L.w[L_DIR], 280, 24),
                        seg_of(w.info, 256, P + L.w[L_DIR] + 24, 280, 256), false, false, w.u, 128,
                        epi_of(P + L.b[L_DIR], ACT_RELU), 1, st));
  // colour = sigmoid (nerf.py:99, :119)
  NT_TRY(nt_launch_gemm(ctx, S, 3, seg_of(w.u, 128, P + L.w[L_COLOR], 128, 128), kNoSeg, false, false, rgb, 3,
                        epi_of(P + L.b[L_COLOR], ACT_SIGMOID), 1, st));
  if (train) NT_CUDA(cudaMemcpyAsync(w.rgb, rgb, (size_t)S * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return NT_OK;
}

int nt_mlp_f32_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                       const float* params, float* rgb, float* sigma, void* ws, size_t ws_bytes, int train,
                       cudaStream_t st) {
  if (ws_bytes < nt_mlp_f32_workspace_bytes(n, p, train)) {
    nt_set_error("fp32 MLP workspace too small: have %zu need %zu", ws_bytes, nt_mlp_f32_workspace_bytes(n, p, train));
    return NT_ERR_WORKSPACE;
  }
  if (train) {
    if (n * p > (int64_t)65535 * 128) {
      nt_set_error("fp32 train batch too large");
      return NT_ERR_INVALID;
    }
    F32Ws w = carve_f32(ws, n * p, true);
    return forward_chunk(ctx, n, p, t, rays, dir_enc, params, rgb, sigma, w, true, st);
  }
  int64_t chunk_rays = F32_RENDER_CHUNK_SAMPLES / p;
  if (chunk_rays < 1) chunk_rays = 1;
  for (int64_t r0 = 0; r0 < n; r0 += chunk_rays) {
    int64_t nr = n - r0 < chunk_rays ? n - r0 : chunk_rays;
    F32Ws w = carve_f32(ws, nr * p, false);
    NT_TRY(forward_chunk(ctx, nr, p, t + r0 * p, rays + r0 * 16, dir_enc + r0 * 24, params, rgb + r0 * p * 3,
                         sigma + r0 * p, w, false, st));
  }
  return NT_OK;
}

// Network.forward on caller-supplied encodings (the reference's stand-alone module call): fp32, render-only, chunked
int nt_network_f32_forward(nt_ctx* ctx, int64_t S, const float* enc_point, const float* enc_dir, const float* params,
                           float* rgb, float* sigma, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t chunk = F32_RENDER_CHUNK_SAMPLES;
  if (ws_bytes < carve_f32(nullptr, S < chunk ? S : chunk, false).bytes) {
    nt_set_error("network forward workspace too small");
    return NT_ERR_WORKSPACE;
  }
  for (int64_t s0 = 0; s0 < S; s0 += chunk) {
    const int64_t ns = S - s0 < chunk ? S - s0 : chunk;
    F32Ws w = carve_f32(ws, ns, false);
    w.enc = const_cast<float*>(enc_point) + s0 * 60;   // read-only in the forward
    w.denc = const_cast<float*>(enc_dir) + s0 * 24;
    NT_TRY(forward_from_enc(ctx, (int)ns, params, rgb + s0 * 3, sigma + s0, w, false, st));
  }
  return NT_OK;
}

// ---- head activations backward (B.6) -------------------------------------------------------------
__global__ void heads_backward_kernel(int64_t S, const float* __restrict__ rgb, const float* __restrict__ zsig,
                                      const float* __restrict__ g_rgb, const float* __restrict__ g_sigma,
                                      float* __restrict__ g_zrgb, float* __restrict__ g_zsig) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float y = rgb[i * 3 + c];
    g_zrgb[i * 3 + c] = g_rgb[i * 3 + c] * y * (1.f - y);
  }
  float z = zsig[i];
  g_zsig[i] = z > 0.f ? g_sigma[i] : (z < 0.f ? -g_sigma[i] : 0.f);
}

static int split_for(int64_t S) {
  int64_t s = S / 512;
  if (s < 1) s = 1;
  if (s > 96) s = 96;
  return (int)s;
}

int nt_mlp_f32_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* params,
                        const float* g_rgb, const float* g_sigma, float* G, float* g_t, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  if (ws_bytes < nt_mlp_f32_workspace_bytes(n, p, 1)) {
    nt_set_error("fp32 MLP workspace too small for backward");
    return NT_ERR_WORKSPACE;
  }
  const LayerTable L = nt_layers();
  const float* P = params;
  const int S = (int)(n * p);
  if (S == 0) return NT_OK;
  F32Ws w = carve_f32(ws, S, true);
  const int sk = split_for(S);
  GemmEpi at = epi_of(nullptr, ACT_NONE);
  at.atomic = 1;

  heads_backward_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(S, w.rgb, w.zsig, g_rgb, g_sigma, w.g_zrgb, w.g_zsig);
  NT_LAUNCH_CHECK(ctx);

  // colour layer: dW = g_z^T u, db, g_u = (g_z W) * relu'(u)
  NT_TRY(nt_launch_gemm(ctx, 3, 128, seg_of(w.g_zrgb, 3, w.u, 128, S), kNoSeg, true, true, G + L.w[L_COLOR], 128, at, sk, st));
  NT_TRY(nt_launch_colsum(ctx, w.g_zrgb, S, 3, 3, G + L.b[L_COLOR], st));
  {
    GemmEpi e = epi_of(nullptr, ACT_NONE);
    e.mask = w.u;
    e.ldmask = 128;
    NT_TRY(nt_launch_gemm(ctx, S, 128, seg_of(w.g_zrgb, 3, P + L.w[L_COLOR], 128, 3), kNoSeg, false, true, w.g_u, 128, e, 1, st));
  }
  // dir_info: input [denc | info]
  NT_TRY(nt_launch_gemm(ctx, 128, 24, seg_of(w.g_u, 128, w.denc, 24, S), kNoSeg, true, true, G + L.w[L_DIR], 280, at, sk, st));
  NT_TRY(nt_launch_gemm(ctx, 128, 256, seg_of(w.g_u, 128, w.info, 256, S), kNoSeg, true, true, G + L.w[L_DIR] + 24, 280, at, sk, st));
  NT_TRY(nt_launch_colsum(ctx, w.g_u, S, 128, 128, G + L.b[L_DIR], st));
  // g_info = g_u W_d[:, 24:]  (no activation on point_info)
  float* g_info = w.gA;
  NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.g_u, 128, P + L.w[L_DIR] + 24, 280, 128), kNoSeg, false, true, g_info, 256,
                        epi_of(nullptr, ACT_NONE), 1, st));
  // point_info layer
  NT_TRY(nt_launch_gemm(ctx, 256, 256, seg_of(g_info, 256, w.h[7], 256, S), kNoSeg, true, true, G + L.w[L_INFO], 256, at, sk, st));
  NT_TRY(nt_launch_colsum(ctx, g_info, S, 256, 256, G + L.b[L_INFO], st));
  // sigma head
  NT_TRY(nt_launch_gemm(ctx, 1, 256, seg_of(w.g_zsig, 1, w.h[7], 256, S), kNoSeg, true, true, G + L.w[L_SIGMA], 256, at, sk, st));
  NT_TRY(nt_launch_colsum(ctx, w.g_zsig, S, 1, 1, G + L.b[L_SIGMA], st));
  // g_pre7 = (g_zsig w_sigma + g_info W_p) * relu'(h7)
  float* gcur = w.gB;
  NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(w.g_zsig, 1, P + L.w[L_SIGMA], 256, 1), kNoSeg, false, true, gcur, 256,
                        epi_of(nullptr, ACT_NONE), 1, st));
  {
    GemmEpi e = epi_of(nullptr, ACT_NONE);
    e.accumulate = 1;
    e.mask = w.h[7];
    e.ldmask = 256;
    NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(g_info, 256, P + L.w[L_INFO], 256, 256), kNoSeg, false, true, gcur, 256, e, 1, st));
  }
  float* gnext = w.gA;
  for (int i = 7; i >= 1; --i) {
    // dW_i, db_i
    if (i == 4) {
      NT_TRY(nt_launch_gemm(ctx, 256, 256, seg_of(gcur, 256, w.h[3], 256, S), kNoSeg, true, true, G + L.w[L_P4], 316, at, sk, st));
      NT_TRY(nt_launch_gemm(ctx, 256, 60, seg_of(gcur, 256, w.enc, 60, S), kNoSeg, true, true, G + L.w[L_P4] + 256, 316, at, sk, st));
      if (g_t)  // g_enc = g_pre4 W4[:, 256:]
        NT_TRY(nt_launch_gemm(ctx, S, 60, seg_of(gcur, 256, P + L.w[L_P4] + 256, 316, 256), kNoSeg, false, true, w.g_enc, 60,
                              epi_of(nullptr, ACT_NONE), 1, st));
    } else {
      NT_TRY(nt_launch_gemm(ctx, 256, 256, seg_of(gcur, 256, w.h[i - 1], 256, S), kNoSeg, true, true, G + L.w[i], 256, at, sk, st));
    }
    NT_TRY(nt_launch_colsum(ctx, gcur, S, 256, 256, G + L.b[i], st));
    // g_pre_{i-1} = (g_pre_i W_i[:, :256]) * relu'(h_{i-1})
    GemmEpi e = epi_of(nullptr, ACT_NONE);
    e.mask = w.h[i - 1];
    e.ldmask = 256;
    NT_TRY(nt_launch_gemm(ctx, S, 256, seg_of(gcur, 256, P + L.w[i], kLayerIn[i], 256), kNoSeg, false, true, gnext, 256, e, 1, st));
    float* tmp = gcur;
    gcur = gnext;
    gnext = tmp;
  }
  // layer 0
  NT_TRY(nt_launch_gemm(ctx, 256, 60, seg_of(gcur, 256, w.enc, 60, S), kNoSeg, true, true, G + L.w[L_P0], 60, at, sk, st));
  NT_TRY(nt_launch_colsum(ctx, gcur, S, 256, 256, G + L.b[L_P0], st));
  if (g_t) {
    GemmEpi e = epi_of(nullptr, ACT_NONE);
    e.accumulate = 1;
    NT_TRY(nt_launch_gemm(ctx, S, 60, seg_of(gcur, 256, P + L.w[L_P0], 60, 256), kNoSeg, false, true, w.g_enc, 60, e, 1, st));
    NT_TRY(nt_launch_encode_backward(ctx, n, p, t, rays, w.g_enc, 60, g_t, st));
  }
  return NT_OK;
}
