// C-ABI glue: context, error reporting, precision dispatch and the fused render drivers
// (render_rays nerf.py:286-323 forward; SURVEY.md §3.5 order for backward).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

void nt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* nt_last_error(void) { return g_err; }
extern "C" int nt_version(void) { return 100; }
extern "C" int64_t nt_param_count(void) { return NT_N_PARAMS; }

extern "C" int nt_layer_table(nt_layer_desc out[NT_N_LAYERS]) {
  if (!out) return NT_ERR_INVALID;
  LayerTable L = nt_layers();
  for (int i = 0; i < NT_N_LAYERS; ++i) {
    out[i].out_features = kLayerOut[i];
    out[i].in_features = kLayerIn[i];
    out[i].weight_offset = L.w[i];
    out[i].bias_offset = L.b[i];
  }
  return NT_OK;
}

extern "C" int nt_create(nt_ctx** out, int device, int n_coarse, int n_fine) {
  NT_REQUIRE(out, "null out");
  // nerf.py:170 takes any counts (main.py:28-29 reads N_COARSE / N_FINE from the .ini); the one-warp-per-ray kernels hold
  // 32-sample slices per lane and the merged ray must fit their 256 sort slots
  NT_REQUIRE(n_coarse >= 32 && n_coarse <= 128 && n_coarse % 32 == 0 && n_fine >= 32 && n_fine % 32 == 0 &&
                 n_coarse + n_fine <= 256,
             "n_coarse / n_fine must be multiples of 32 with n_coarse <= 128 and n_coarse + n_fine <= 256");
  int count = 0;
  NT_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    nt_set_error("no CUDA device %d (found %d); libnerftiny has no CPU fallback", device, count);
    return NT_ERR_CUDA;
  }
  nt_ctx probe;
  probe.device = device;
  NT_ENTER(&probe);  // allocate the status flags on `device`; the caller's current device is restored on return
  cudaDeviceProp prop;
  NT_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    nt_set_error("device %d is sm_%d%d; libnerftiny is built for sm_100a only", device, prop.major, prop.minor);
    return NT_ERR_UNSUPPORTED;
  }
  nt_ctx* c = new (std::nothrow) nt_ctx();
  NT_REQUIRE(c, "out of host memory");
  c->device = device;
  c->n_coarse = n_coarse;
  c->n_fine = n_fine;
  c->sm_count = prop.multiProcessorCount;
  c->launches = 0;
  c->opt_detach_t_fine = 0;
  c->opt_tc_version = 0;
  c->attr_done = 0;
  c->side = nullptr;
  c->ev_fork = c->ev_join = nullptr;
  c->defer_dw = c->dw_pending = 0;
  c->shared_wb = nullptr;
  c->share_wb = 0;
  {
    // measured on B200 (1024 / 4096 rays per step, gpurun_out/r2j_sweep*.log): 0 -> 1.508 / 5.133 ms, 72 -> 1.480 / 5.208,
    // 88 -> 1.466 / 5.003 (best), 96 -> 1.472 / 5.123, 112 -> 1.668 / 5.669: the HBM-bound launch needs ~60 % of the SMs
    const char* e = getenv("NT_DW_OVERLAP_CTAS");
    c->opt_dw_overlap_ctas = e ? atoi(e) : (c->sm_count * 80) / 148;  // re-swept after the chain kernel got faster: 80 -> 1.31 / 4.66 ms, 88 -> 1.34 / 4.70
    if (c->opt_dw_overlap_ctas < 0 || c->opt_dw_overlap_ctas >= c->sm_count) c->opt_dw_overlap_ctas = 0;
  }
  c->last_delta = 1e-4f;
  c->d_flags = nullptr;
  if (cudaMalloc(&c->d_flags, 4 * sizeof(int)) != cudaSuccess || cudaMemset(c->d_flags, 0, 4 * sizeof(int)) != cudaSuccess) {
    nt_set_error("cudaMalloc of the status flags failed");
    delete c;
    return NT_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    nt_set_error("creating the side stream / events failed");
    nt_destroy(c);
    return NT_ERR_CUDA;
  }
  *out = c;
  return NT_OK;
}

extern "C" void nt_destroy(nt_ctx* ctx) {
  if (!ctx) return;
  NT_ENTER(ctx);
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  delete ctx;
}

extern "C" int64_t nt_launch_count(const nt_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int nt_set_option(nt_ctx* ctx, int key, int value) {
  NT_REQUIRE(ctx, "null ctx");
  if (key == NT_OPT_DETACH_T_FINE) {
    ctx->opt_detach_t_fine = value != 0;
    return NT_OK;
  }
  if (key == NT_OPT_MLP_TC_VERSION) {
    NT_REQUIRE(value == 0 || value == 5 || value == 7, "mlp_tc version must be 0 (default), 5 or 7");
    ctx->opt_tc_version = value;
    return NT_OK;
  }
  if (key == NT_OPT_DW_OVERLAP_CTAS) {
    NT_REQUIRE(value >= 0 && value < ctx->sm_count, "dW overlap CTAs must be in [0, SM count)");
    ctx->opt_dw_overlap_ctas = value;
    return NT_OK;
  }
  if (key == NT_OPT_LAST_DELTA) {
    float f;
    memcpy(&f, &value, sizeof(f));
    NT_REQUIRE(f == f && f >= 0.f, "last delta must be a non-negative float (passed as its bit pattern)");
    ctx->last_delta = f;
    return NT_OK;
  }
  nt_set_error("unknown option %d", key);
  return NT_ERR_INVALID;
}

// ---- precision dispatch ------------------------------------------------------------------------
static bool is_tc16(int precision) { return precision == NT_PREC_BF16 || precision == NT_PREC_FP16; }
static bool known_precision(int precision) {
  return precision == NT_PREC_FP32 || precision == NT_PREC_TC32 || precision == NT_PREC_MIXED || is_tc16(precision);
}
// NT_PREC_MIXED packs both operand images: [split-fp16 image (coarse pass) | fp16 image (fine pass)]
static size_t mixed_fp16_offset() { return (nt_mlp_tc32_packed_bytes() + 1023) & ~(size_t)1023; }

extern "C" size_t nt_mlp_workspace_bytes(nt_ctx* ctx, int precision, int64_t n, int p, int train) {
  (void)ctx;
  if (precision == NT_PREC_FP32) return nt_mlp_f32_workspace_bytes(n, p, train);
  if (precision == NT_PREC_BF16) return train ? nt_mlp_bf16_train_workspace_bytes(n, p) : 256;
  if (precision == NT_PREC_FP16 || precision == NT_PREC_TC32 || precision == NT_PREC_MIXED) return train ? 0 : 256;  // rendering modes
  return 0;
}
extern "C" size_t nt_packed_weight_bytes(nt_ctx* ctx, int precision) {
  (void)ctx;
  if (is_tc16(precision)) return nt_mlp_tc_packed_bytes();
  if (precision == NT_PREC_MIXED) return mixed_fp16_offset() + nt_mlp_tc_packed_bytes();
  return precision == NT_PREC_TC32 ? nt_mlp_tc32_packed_bytes() : 0;
}
extern "C" int nt_pack_weights(nt_ctx* ctx, int precision, const float* params, void* packed, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && params, "null pointer");
  if (precision == NT_PREC_FP32) return NT_OK;
  NT_REQUIRE(known_precision(precision) && packed, "bad precision / null packed buffer");
  if (precision == NT_PREC_MIXED) {
    int rc = nt_mlp_tc_pack(ctx, params, packed, 2, (cudaStream_t)stream);
    if (rc != NT_OK) return rc;
    return nt_mlp_tc_pack(ctx, params, (char*)packed + mixed_fp16_offset(), 1, (cudaStream_t)stream);
  }
  const int mode = precision == NT_PREC_BF16 ? 0 : (precision == NT_PREC_FP16 ? 1 : 2);
  return nt_mlp_tc_pack(ctx, params, packed, mode, (cudaStream_t)stream);
}

extern "C" int nt_mlp_forward(nt_ctx* ctx, int precision, int64_t n, int p, const float* t, const float* rays,
                              const float* dir_enc, const float* params, const void* packed, float* rgb, float* sigma,
                              void* ws, size_t ws_bytes, int train, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t && rays && dir_enc && params && rgb && sigma, "null pointer");
  NT_REQUIRE(p > 0 && n >= 0, "bad shape");
  if (n == 0) return NT_OK;
  if (precision == NT_PREC_FP32) {
    NT_REQUIRE(ws, "fp32 MLP needs a workspace");
    return nt_mlp_f32_forward(ctx, n, p, t, rays, dir_enc, params, rgb, sigma, ws, ws_bytes, train, (cudaStream_t)stream);
  }
  if (!known_precision(precision)) {
    nt_set_error("unknown precision %d", precision);
    return NT_ERR_UNSUPPORTED;
  }
  NT_REQUIRE(packed, "the tensor-core precisions need packed weights (nt_pack_weights with the same precision)");
  if (train) {  // fused tcgen05 forward + bf16 activation stash for the tensor-core backward
    if (precision != NT_PREC_BF16) {
      nt_set_error("NT_PREC_FP16 / NT_PREC_TC32 / NT_PREC_MIXED are rendering modes; train with NT_PREC_BF16 or NT_PREC_FP32");
      return NT_ERR_UNSUPPORTED;
    }
    NT_REQUIRE(ws, "bf16 training needs a workspace");
    return nt_mlp_bf16_train_forward(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, ws, ws_bytes,
                                     (cudaStream_t)stream);
  }
  if (precision == NT_PREC_MIXED) {
    // the pass over the COARSE samples decides where the fine samples go (resample amplifies its errors):
    // coarse-sized passes (p == n_coarse) run split-fp16, everything else single-pass fp16
    if (p == ctx->n_coarse) return nt_mlp_tc32_forward(ctx, n, p, t, rays, dir_enc, packed, rgb, sigma, (cudaStream_t)stream);
    return nt_mlp_tc_forward(ctx, n, p, t, rays, dir_enc, params, (const char*)packed + mixed_fp16_offset(), rgb, sigma, 1,
                             (cudaStream_t)stream);
  }
  if (precision == NT_PREC_TC32) return nt_mlp_tc32_forward(ctx, n, p, t, rays, dir_enc, packed, rgb, sigma, (cudaStream_t)stream);
  return nt_mlp_tc_forward(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, precision == NT_PREC_FP16 ? 1 : 0,
                           (cudaStream_t)stream);
}

extern "C" int nt_mlp_backward(nt_ctx* ctx, int precision, int64_t n, int p, const float* t, const float* rays,
                               const float* dir_enc, const float* params, const void* packed, const float* rgb,
                               const float* g_rgb, const float* g_sigma, float* grads, float* g_t, void* ws,
                               size_t ws_bytes, void* stream) {
  NT_ENTER(ctx);
  (void)dir_enc;
  (void)packed;
  NT_REQUIRE(ctx && t && rays && params && g_rgb && g_sigma && grads && ws, "null pointer");
  NT_REQUIRE(precision == NT_PREC_FP32 || precision == NT_PREC_BF16, "unknown precision");
  if (n == 0) return NT_OK;
  if (precision == NT_PREC_BF16)
    return nt_mlp_bf16_train_backward(ctx, n, p, t, rays, params, rgb, g_rgb, g_sigma, grads, g_t, ws, ws_bytes,
                                      (cudaStream_t)stream);
  return nt_mlp_f32_backward(ctx, n, p, t, rays, params, g_rgb, g_sigma, grads, g_t, ws, ws_bytes, (cudaStream_t)stream);
}

// ---- fused render driver ---------------------------------------------------------------------------
struct RenderWs {
  float *rays, *dir_enc, *t_c, *rgb_c, *sig_c, *w_c, *t_f, *rgb_f, *sig_f;
  uint8_t* perm;
  float *g_rgb_c, *g_sig_c, *g_rgb_f, *g_sig_f, *g_t_f, *g_w_c, *g_t_mlp;
  void *mlp_c, *mlp_f;
  size_t mlp_c_bytes, mlp_f_bytes;
  size_t bytes;
};

static RenderWs carve_render(nt_ctx* ctx, void* base, int precision, int64_t n, int train) {
  RenderWs w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) {
    void* p = (void*)(b + off);
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  const int nc = ctx->n_coarse, nf = ctx->n_fine;
  w.rays = (float*)take(n * 16 * 4);
  w.dir_enc = (float*)take(n * 24 * 4);
  w.t_c = (float*)take(n * nc * 4);
  w.rgb_c = (float*)take(n * nc * 12);
  w.sig_c = (float*)take(n * nc * 4);
  w.w_c = (float*)take(n * nc * 4);
  w.t_f = (float*)take(n * nf * 4);
  w.rgb_f = (float*)take(n * nf * 12);
  w.sig_f = (float*)take(n * nf * 4);
  w.mlp_c_bytes = nt_mlp_workspace_bytes(ctx, precision, n, nc, train);
  w.mlp_f_bytes = nt_mlp_workspace_bytes(ctx, precision, n, nf, train);
  if (train) {
    w.perm = (uint8_t*)take(n * 5 * (nc + nf));
    w.g_rgb_c = (float*)take(n * nc * 12);
    w.g_sig_c = (float*)take(n * nc * 4);
    w.g_rgb_f = (float*)take(n * nf * 12);
    w.g_sig_f = (float*)take(n * nf * 4);
    w.g_t_f = (float*)take(n * nf * 4);
    w.g_w_c = (float*)take(n * nc * 4);
    w.g_t_mlp = (float*)take(n * nf * 4);
    w.mlp_c = take(w.mlp_c_bytes);
    w.mlp_f = take(w.mlp_f_bytes);
  } else {
    size_t m = w.mlp_c_bytes > w.mlp_f_bytes ? w.mlp_c_bytes : w.mlp_f_bytes;
    w.mlp_c = w.mlp_f = take(m);
  }
  w.bytes = off;
  return w;
}

extern "C" size_t nt_render_workspace_bytes(nt_ctx* ctx, int precision, int64_t n, int train) {
  if (!ctx || n < 0) return 0;
  return carve_render(ctx, nullptr, precision, n, train).bytes;
}

#define NT_TRY(x)                   \
  do {                              \
    int rc__ = (x);                 \
    if (rc__ != NT_OK) return rc__; \
  } while (0)

extern "C" int nt_render_forward(nt_ctx* ctx, int precision, int64_t n, const int64_t* row, const int64_t* col,
                                 const float* c2w, int c2w_stride, const float* kinv, const float* near_,
                                 const float* far_, const float* params, const void* packed, int any_step_zero,
                                 const float* delta0, float* c_coarse, float* c_fine, void* ws, size_t ws_bytes,
                                 int train, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && row && col && c2w && kinv && near_ && far_ && params && c_coarse && c_fine && ws, "null pointer");
  if (n <= 0) return NT_OK;
  RenderWs w = carve_render(ctx, ws, precision, n, train);
  if (ws_bytes < w.bytes) {
    nt_set_error("render workspace too small: have %zu need %zu", ws_bytes, w.bytes);
    return NT_ERR_WORKSPACE;
  }
  const int nc = ctx->n_coarse, nf = ctx->n_fine;
  NT_TRY(nt_raygen(ctx, n, row, col, c2w, c2w_stride, kinv, w.rays, nullptr, w.dir_enc, stream));
  NT_TRY(nt_launch_sample_coarse(ctx, n, near_, far_, any_step_zero, delta0 ? delta0 + 1 : nullptr, w.t_c,
                                 (cudaStream_t)stream));                                                // nerf.py:288
  NT_TRY(nt_mlp_forward(ctx, precision, n, nc, w.t_c, w.rays, w.dir_enc, params, packed, w.rgb_c, w.sig_c, w.mlp_c,
                        w.mlp_c_bytes, train, stream));                                                // nerf.py:289
  NT_TRY(nt_composite_coarse(ctx, n, near_, far_, w.rgb_c, w.sig_c, w.w_c, c_coarse, stream));         // nerf.py:293-295, 320
  NT_TRY(nt_sample_pdf(ctx, n, w.t_c, w.w_c, delta0, w.t_f, nullptr, stream));                         // nerf.py:298
  NT_TRY(nt_mlp_forward(ctx, precision, n, nf, w.t_f, w.rays, w.dir_enc, params, packed, w.rgb_f, w.sig_f, w.mlp_f,
                        w.mlp_f_bytes, train, stream));                                                // nerf.py:299
  NT_TRY(nt_composite_fine(ctx, n, w.t_c, w.rgb_c, w.sig_c, w.t_f, w.rgb_f, w.sig_f, ctx->last_delta, c_fine, nullptr,
                           train ? w.perm : nullptr, stream));                                         // nerf.py:302-321
  return NT_OK;
}

extern "C" int nt_render_backward(nt_ctx* ctx, int precision, int64_t n, const float* near_, const float* far_,
                                  const float* params, const void* packed, const float* delta0,
                                  const float* g_c_coarse, const float* g_c_fine, float* grads, void* ws,
                                  size_t ws_bytes, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && near_ && far_ && params && g_c_coarse && g_c_fine && grads && ws, "null pointer");
  if (n <= 0) return NT_OK;
  RenderWs w = carve_render(ctx, ws, precision, n, 1);
  if (ws_bytes < w.bytes) {
    nt_set_error("render workspace too small: have %zu need %zu", ws_bytes, w.bytes);
    return NT_ERR_WORKSPACE;
  }
  const int nc = ctx->n_coarse, nf = ctx->n_fine;
  // C_fine <- sort/composite (B.2, B.3)
  NT_TRY(nt_composite_fine_backward(ctx, n, w.t_c, w.rgb_c, w.sig_c, w.t_f, w.rgb_f, w.sig_f, ctx->last_delta, w.perm, g_c_fine,
                                    w.g_rgb_c, w.g_sig_c, w.g_rgb_f, w.g_sig_f, w.g_t_f, stream));
  // fine MLP: dW + input gradient down to t_fine (B.4, B.6, B.7)
  const bool detach = ctx->opt_detach_t_fine != 0;
  ctx->defer_dw = (precision == NT_PREC_BF16 && ctx->opt_dw_overlap_ctas > 0) ? 1 : 0;
  ctx->share_wb = 1;  // both passes use the same parameters: pack the transposed weights once
  ctx->shared_wb = nullptr;
  int rc_fine = nt_mlp_backward(ctx, precision, n, nf, w.t_f, w.rays, w.dir_enc, params, packed, w.rgb_f, w.g_rgb_f, w.g_sig_f, grads,
                                detach ? nullptr : w.g_t_mlp, w.mlp_f, w.mlp_f_bytes, stream);
  ctx->defer_dw = 0;
  if (rc_fine != NT_OK) ctx->share_wb = 0;
  NT_TRY(rc_fine);
  // g_t_fine = compositing path + MLP-input path (summed inside the kernel); then resample backward (B.5)
  if (!detach)
    NT_TRY(nt_launch_sample_pdf_backward(ctx, n, w.t_c, w.w_c, delta0, w.g_t_f, w.g_t_mlp, w.g_w_c, (cudaStream_t)stream));
  // C_coarse <- composite; rgb/sigma of the coarse samples also feed the fine composite: accumulate onto those gradients
  NT_TRY(nt_launch_composite_coarse_backward(ctx, n, near_, far_, w.rgb_c, w.sig_c, g_c_coarse, detach ? nullptr : w.g_w_c,
                                             w.g_rgb_c, w.g_sig_c, 1, (cudaStream_t)stream));
  // coarse MLP: dW only (t_coarse is a constant)
  int rc_coarse = nt_mlp_backward(ctx, precision, n, nc, w.t_c, w.rays, w.dir_enc, params, packed, w.rgb_c, w.g_rgb_c, w.g_sig_c,
                                  grads, nullptr, w.mlp_c, w.mlp_c_bytes, stream);
  ctx->share_wb = 0;
  ctx->shared_wb = nullptr;
  NT_TRY(rc_coarse);
  NT_TRY(nt_join_deferred_dw(ctx, (cudaStream_t)stream));
  return NT_OK;
}

// Network.forward (nerf.py:101-124) as the reference exposes it: caller-supplied encodings, fp32 accuracy path.
// The workspace is that of nt_mlp_workspace_bytes(ctx, NT_PREC_FP32, total, 1, 0).
extern "C" int nt_network_forward(nt_ctx* ctx, int64_t total, const float* enc_point, const float* enc_dir,
                                  const float* params, float* rgb, float* sigma, void* ws, size_t ws_bytes, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx, "null ctx");
  if (total <= 0) return NT_OK;
  NT_REQUIRE(enc_point && enc_dir && params && rgb && sigma && ws, "null pointer");
  return nt_network_f32_forward(ctx, total, enc_point, enc_dir, params, rgb, sigma, ws, ws_bytes, (cudaStream_t)stream);
}

// diagnostic: NT_PREC_BF16 forward that also dumps one MMA layer's post-activation fp32 output (dbg dev [N*P,256];
// layer 0..7 = trunk, 8 = point_info, 9 = dir_info) — used by the layer-by-layer parity test
extern "C" int nt_mlp_forward_debug(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays,
                                    const float* dir_enc, const float* params, const void* packed, float* rgb,
                                    float* sigma, float* dbg, int layer, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && t && rays && dir_enc && params && packed && rgb && sigma && dbg, "null pointer");
  if (n <= 0) return NT_OK;
  return nt_mlp_tc_forward_dbg(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, dbg, layer, (cudaStream_t)stream);
}

// diagnostic: raw bf16 tensor-core GEMM (gemm_tc.cu) for the unit tests.  mn_major = 0: A [M][K], B [N][K];
// 1: A [K][M], B [K][N].  out_f32 = 0: C bf16 [M][ldc]; 1: fp32 atomically accumulated into C.
extern "C" int nt_gemm_bf16_debug(nt_ctx* ctx, int mn_major, int m, int n, int k, const void* a, int lda, const void* b,
                                  int ldb, void* c, int ldc, int out_f32, const void* mask, int ldmask, void* stream) {
  NT_ENTER(ctx);
  NT_REQUIRE(ctx && a && b && c, "null pointer");
  GemmTcEpi e;
  memset(&e, 0, sizeof(e));
  e.C = c;
  e.ldc = ldc;
  e.atomic_f32 = out_f32;
  e.mask = mask;
  e.ldmask = ldmask;
  return nt_launch_gemm_tc(ctx, mn_major, m, n, k, a, lda, b, ldb, e, (cudaStream_t)stream);
}
