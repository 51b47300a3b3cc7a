// Shared declarations for libnerftiny (sm_100a).  Host+device helpers, the layer table and
// the internal launch functions each .cu file exports to api.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nerftiny.h"

#define NT_WARP 32
#define NT_MAX_COARSE 128
#define NT_MAX_MERGED 256
#define NT_MAX_DEVICES 64

struct nt_ctx {
  int device;
  int n_coarse, n_fine;
  int sm_count;
  int* d_flags;  // [0] = any_step_zero scratch, [1] = resample range status
  int64_t launches;
  int opt_detach_t_fine;
  int opt_tc_version;
  float last_delta;    // delta of the last merged sample (render_rays `last`, nerf.py:286, :311); default 1e-4
  // training: the HBM-bound weight-gradient launch of the FINE pass can run on `side`, capped to opt_dw_overlap_ctas CTAs,
  // next to the MMA-bound backward of the coarse pass (fork / join by events, capturable into a CUDA graph)
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int opt_dw_overlap_ctas;  // 0 = off
  int defer_dw;             // set by nt_render_backward around the fine pass
  int dw_pending;           // a deferred launch has not been joined yet
  // nt_render_backward packs the transposed weights once: the fine pass packs into its workspace and leaves the pointer here
  // for the coarse pass (same parameters); -1 = not inside nt_render_backward (every nt_mlp_backward call packs for itself)
  const void* shared_wb;
  int share_wb;
  unsigned attr_done;  // NT_ATTR_*: cudaFuncSetAttribute is per DEVICE, so the "already opted in" bits live in the ctx
};
enum {
  NT_ATTR_MLP_TC5 = 1u << 0, NT_ATTR_MLP_TC7 = 1u << 1, NT_ATTR_MLP_TC32 = 1u << 2, NT_ATTR_GEMM_TC = 1u << 3,
  NT_ATTR_DW_GROUPED = 1u << 4, NT_ATTR_DW_GEMM = 1u << 5, NT_ATTR_BWD_TC = 1u << 6
};

void nt_set_error(const char* fmt, ...);

// Every launching entry point runs on the context's device, whatever device the caller has current, and leaves the
// caller's current device untouched (nt_create does not switch it either).
struct NtDeviceGuard {
  int prev;
  bool changed;
  explicit NtDeviceGuard(const nt_ctx* ctx) : prev(-1), changed(false) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) changed = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~NtDeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};
#define NT_ENTER(ctx) NtDeviceGuard nt_guard__(ctx)

#define NT_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      nt_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
      return NT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define NT_LAUNCH_CHECK(ctx)                                                               \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      nt_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return NT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define NT_REQUIRE(cond, msg)                                                              \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      nt_set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg);                 \
      return NT_ERR_INVALID;                                                               \
    }                                                                                      \
  } while (0)

// ---- layer table (nerf.py:85-99) --------------------------------------------------------
enum { L_P0 = 0, L_P1, L_P2, L_P3, L_P4, L_P5, L_P6, L_P7, L_SIGMA, L_INFO, L_DIR, L_COLOR };
static const int kLayerOut[NT_N_LAYERS] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
static const int kLayerIn[NT_N_LAYERS] = {60, 256, 256, 256, 316, 256, 256, 256, 256, 256, 280, 128};

struct LayerTable {
  int64_t w[NT_N_LAYERS], b[NT_N_LAYERS];
};
static inline LayerTable nt_layers() {
  LayerTable t;
  int64_t off = 0;
  for (int i = 0; i < NT_N_LAYERS; ++i) {
    t.w[i] = off;
    off += (int64_t)kLayerOut[i] * kLayerIn[i];
    t.b[i] = off;
    off += kLayerOut[i];
  }
  return t;
}

// ---- encoder frequencies (nerf.py:141-145; SURVEY.md Appendix A.4, fp32 bit patterns) -------
#define NT_FREQ_POINT_INIT                                                                       \
  {0x40490fdbu, 0x40d928aeu, 0x416a8b6cu, 0x41fd527bu, 0x4288cd33u, 0x4313c0fau, 0x439f953cu,   \
   0x442c5befu, 0x44ba2881u, 0x45490fdbu}
#define NT_FREQ_DIR_INIT {0x40490fdbu, 0x40fd527au, 0x419f953cu, 0x42490fdbu}

// ---- internal launchers (one per .cu) ---------------------------------------------------------
// geom.cu
int nt_launch_sample_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int any_step_zero,
                            const float* flag_dev, float* t_coarse, cudaStream_t st);
int nt_launch_encode_points(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, float* enc, int ld_enc,
                            cudaStream_t st);
int nt_launch_expand_dir_enc(nt_ctx* ctx, int64_t n, int p, const float* dir_enc, float* out, int ld, cudaStream_t st);
int nt_launch_encode_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* g_enc,
                              int ld, float* g_t, cudaStream_t st);

// gemm_f32.cu
struct GemmSeg {
  const float* A;
  int lda;
  const float* B;
  int ldb;
  int K;
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_ABS = 2, ACT_SIGMOID = 3 };
struct GemmEpi {
  const float* bias;   // [N] or null
  int act;             // ACT_*
  const float* mask;   // [M, ldmask]: out = mask>0 ? acc : 0  (null = none)
  int ldmask;
  int accumulate;      // C += result
  int atomic;          // atomicAdd into C (split-K)
  float* pre_out;      // optional: pre-activation copy (ACT_ABS backward needs the sign)
  int ldpre;
};
// C[M,N] = epi(sum_seg A_seg op B_seg).  a_km: A stored [K][M] (else [M][K]); b_kn: B stored [K][N] (else [N][K]).
int nt_launch_gemm(nt_ctx* ctx, int M, int N, GemmSeg s0, GemmSeg s1, bool a_km, bool b_kn, float* C, int ldc,
                   GemmEpi epi, int split_k, cudaStream_t st);
int nt_launch_colsum(nt_ctx* ctx, const float* G, int64_t rows, int cols, int ld, float* out, cudaStream_t st);

// mlp_f32.cu
size_t nt_mlp_f32_workspace_bytes(int64_t n, int p, int train);
int nt_network_f32_forward(nt_ctx* ctx, int64_t S, const float* enc_point, const float* enc_dir, const float* params,
                           float* rgb, float* sigma, void* ws, size_t ws_bytes, cudaStream_t st);
int nt_mlp_f32_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                       const float* params, float* rgb, float* sigma, void* ws, size_t ws_bytes, int train,
                       cudaStream_t st);
int nt_mlp_f32_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* params,
                        const float* g_rgb, const float* g_sigma, float* grads, float* g_t, void* ws, size_t ws_bytes,
                        cudaStream_t st);

// mlp_tc.cu (tcgen05 / TMEM fused encode+MLP, bf16 operands)
size_t nt_mlp_tc_packed_bytes();
int nt_mlp_tc_pack(nt_ctx* ctx, const float* params, void* packed, int mode /*0 bf16, 1 fp16, 2 split fp16 (tc32)*/,
                   cudaStream_t st);
int nt_mlp_tc_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                      const float* params, const void* packed, float* rgb, float* sigma, int fp16, cudaStream_t st);
int nt_mlp_tc_forward_dbg(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                          const float* params, const void* packed, float* rgb, float* sigma, float* dbg, int dbg_layer,
                          cudaStream_t st);
// mlp_tc32.cu (NT_PREC_TC32: 3-pass split-fp16 tcgen05 forward with two TMEM accumulators)
size_t nt_mlp_tc32_packed_bytes();
int nt_mlp_tc32_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                        const void* packed, float* rgb, float* sigma, cudaStream_t st);
// composite.cu / sample_pdf.cu: backward launchers with the fusions nt_render_backward uses (accumulate onto an existing
// gradient; second gradient path into t_fine summed on the fly)
int nt_launch_composite_coarse_backward(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, const float* rgb,
                                        const float* sigma, const float* g_c, const float* g_w_ext, float* g_rgb,
                                        float* g_sigma, int accumulate, cudaStream_t st);
int nt_launch_sample_pdf_backward(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                                  const float* g_t_fine, const float* g_t_fine2, float* g_w, cudaStream_t st);
// composite_fine_fwd.cu
int nt_launch_composite_fine_fwd(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c, const float* sigma_c,
                                 const float* t_f, const float* rgb_f, const float* sigma_f, float last, float* c_out,
                                 float* weights, uint8_t* perm, cudaStream_t st);

// gemm_tc.cu — bf16 tcgen05 GEMM (training path)
struct GemmTcEpi {
  void* C;            // bf16 [M][ldc] (store) or fp32 [M][ldc] (atomic_f32 / store_f32)
  int ldc;
  int atomic_f32;     // fp32 vector atomics (split-K weight gradients, fp32 accumulation into an existing buffer)
  int store_f32;      // plain fp32 store
  const float* bias;  // [N] or null
  int relu;
  const void* mask;   // bf16 [M][ldmask] or null: out = mask > 0 ? v : 0
  int ldmask;
  const float* r1_row;  // optional rank-1 term added before the mask: v += r1_row[m] * r1_col[n]
  const float* r1_col;
  float* colsum;        // optional fp32 [N]: += column sums of the stored matrix (fused bias gradient)
  // optional fused encoder backward (N = 64 only): the fp32 row is the gradient w.r.t. the 60 xyz features of sample m; instead
  // of storing it, g_t[m] += sum_c d_wrd_c * sum_l w_l (g_sin cos(w_l p_c) - g_cos sin(w_l p_c))  (SURVEY.md B.4).  enc_gt must
  // be zero on entry (the two column halves of a row add their parts with one atomic each).
  const float* enc_t;     // [M] sample positions t
  const float* enc_rays;  // ray records [M / enc_p][16]
  int enc_p;              // samples per ray
  float* enc_gt;          // [M] or null
};
int nt_launch_gemm_tc(nt_ctx* ctx, int mn_major, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                      const GemmTcEpi& epi, cudaStream_t st);

// bf16 activation stash written by the fused forward (mlp_tc.cu) and consumed by train_bf16.cu
struct TcStash {
  void* layer[10];  // bf16 [S][256] for layers 0..8 (h0..h7, point_info), [S][128] for dir_info (u)
  void* enc;        // bf16 [S][64]
  void* denc;       // bf16 [S][32]
  float* zsig;      // fp32 [S]
  uint32_t* bits;   // ReLU' bit masks of the trunk layers [8][S][8]
};
int nt_mlp_tc_forward_stash(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                            const float* params, const void* packed, float* rgb, float* sigma, const TcStash* stash,
                            cudaStream_t st);
// train_bf16.cu
int nt_join_deferred_dw(nt_ctx* ctx, cudaStream_t st);
size_t nt_mlp_bf16_train_workspace_bytes(int64_t n, int p);
int nt_mlp_bf16_train_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                              const float* params, const void* packed, float* rgb, float* sigma, void* ws, size_t ws_bytes,
                              cudaStream_t st);
int nt_mlp_bf16_train_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* params,
                               const float* rgb, const float* g_rgb, const float* g_sigma, float* grads, float* g_t,
                               void* ws, size_t ws_bytes, cudaStream_t st);
int nt_launch_dw_gemm(nt_ctx* ctx, int S, const void* G, int ldg, int m_valid, const void* H, int ldh, int n_valid,
                      float* C, int ldc, cudaStream_t st);
// grouped weight-gradient GEMM: queue problems of one backward pass (all with the same sample count S), flush once
int nt_dw_group_begin(int S);
int nt_dw_group_add(const void* G, int ldg, int m_valid, const void* H, int ldh, int n_valid, float* C, int ldc);
int nt_dw_group_flush(nt_ctx* ctx, cudaStream_t st, int max_ctas /*0 = all SMs*/);
int nt_make_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows);
// bwd_tc.cu — fused backward-data chain
size_t nt_bwd_tc_packed_bytes();
int nt_bwd_tc_pack(nt_ctx* ctx, const float* params, void* packed, cudaStream_t st);
int nt_bwd_tc_chain(nt_ctx* ctx, int64_t S, const void* g_u, void* const outs[9], const int out_ld[9], const void* packed,
                    int max_ctas, const uint32_t* bits,
                    const float* gzsig, const float* wsig, float* const db[9], cudaStream_t st);
