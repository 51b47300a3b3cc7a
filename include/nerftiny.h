/*
 * nerftiny.h — C-ABI of libnerftiny.so, the B200-native (sm_100a) implementation of the
 * NeRF-tiny per-ray hot path.
 *
 * The reference (D-Hank/NeRF-tiny) is pure Python/PyTorch and has no operator/FFI boundary
 * (SURVEY.md §1, §8(b)); this library is inserted UNDERNEATH the Python methods of
 * `nerf.NeRFModel`.  Each entry point below names the reference code it replaces
 * (file:line into the reference's nerf.py).
 *
 * Conventions
 *  - plain C: pointers + sizes only, no torch / C++ types.
 *  - every pointer marked "dev" is DEVICE memory owned by the caller (e.g. a torch tensor's
 *    data_ptr()); the library never frees or retains it past the call.  Pointers marked
 *    "host" are host memory.
 *  - `stream` is a cudaStream_t passed as void*; calls enqueue work on it and never
 *    synchronise it (exceptions are documented).  All functions are re-entrant per nt_ctx.
 *  - return value: 0 = NT_OK, <0 = error; nt_last_error() gives a thread-local message.
 *  - tensors are dense row-major fp32 unless stated.  N = rays, P = samples per ray,
 *    Nc / Nf = coarse / fine sample counts fixed at nt_create() (64 / 128 by default).
 *  - parameters live in ONE flat fp32 buffer of nt_param_count() floats in the reference's
 *    state_dict order (nerf.py:85-99): for each of the 12 Linear layers, weight (out,in)
 *    row-major followed by bias.  Gradients and Adam moments use the same layout.
 */
#ifndef NERFTINY_H
#define NERFTINY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NT_OK 0
#define NT_ERR_INVALID (-1)   /* bad argument                                        */
#define NT_ERR_CUDA (-2)      /* a CUDA runtime call failed                           */
#define NT_ERR_RANGE (-3)     /* resample index out of range (reference: exit(0))     */
#define NT_ERR_WORKSPACE (-4) /* workspace too small                                  */
#define NT_ERR_UNSUPPORTED (-5)

/* arithmetic of the MLP contraction */
#define NT_PREC_FP32 0 /* fp32 FFMA on CUDA cores, layer-major: the reference arithmetic, forward + backward        */
#define NT_PREC_TC32 1 /* fp32-tolerance mode on the tensor cores: every operand split into two fp16 (hi + 2^-12 lo), */
                       /* 3 tcgen05 MMAs per product into two TMEM accumulators; rendering (forward) only            */
#define NT_PREC_BF16 2 /* bf16 operands, fp32 accumulate, tcgen05/TMEM fused kernels: forward + backward (training)   */
#define NT_PREC_FP16 3 /* fp16 operands (saturating), fp32 accumulate: same kernels and speed as BF16 with 8x smaller */
                       /* operand rounding (the network's activations are bounded); rendering (forward) only          */
#define NT_PREC_MIXED 4 /* rendering: the coarse pass (n_coarse samples per ray, whose weights steer the resampling) in  */
                        /* NT_PREC_TC32 arithmetic, the fine pass in NT_PREC_FP16                                      */

#define NT_N_LAYERS 12
#define NT_N_PARAMS 593924

typedef struct nt_ctx nt_ctx;

typedef struct nt_layer_desc {
  int out_features, in_features;
  int64_t weight_offset, bias_offset; /* float offsets into the flat parameter buffer */
} nt_layer_desc;

/* ---- lifetime / introspection ------------------------------------------------------- */
int nt_version(void);
const char* nt_last_error(void);
/* nerf.py:170 NeRFModel(num_coarse, num_fine, ...) */
int nt_create(nt_ctx** out, int device, int n_coarse, int n_fine);
void nt_destroy(nt_ctx* ctx);
int64_t nt_param_count(void);
/* nerf.py:85-99: the 12 nn.Linear layers in creation order */
int nt_layer_table(nt_layer_desc out[NT_N_LAYERS]);
/* Options.  NT_OPT_DETACH_T_FINE (diagnostic, default 0): nt_render_backward treats t_fine as a constant, i.e. drops
 * the gradient path t_fine -> cdf/weights -> coarse sigma that the reference keeps (nerf.py:255-259).  The fp32
 * gradient along that path is ill-conditioned; the switch lets tests compare the well-conditioned
 * part of the gradient tightly.
 * NT_OPT_MLP_TC_VERSION (default 0 = 7 for rendering; training always 5): schedule of the fused 16-bit encode+MLP kernel —
 * 5 tile pair in lock-step, 7 staggered tiles + tcgen05 cta_group::2.
 * NT_OPT_LAST_DELTA: `last` of render_rays (nerf.py:286, :311; default 1e-4), passed as the BIT PATTERN of the float. */
#define NT_OPT_DETACH_T_FINE 1
#define NT_OPT_MLP_TC_VERSION 2
#define NT_OPT_LAST_DELTA 3
/* NT_OPT_DW_OVERLAP_CTAS (training, NT_PREC_BF16; 0 = off): nt_render_backward launches the HBM-bound weight-gradient
 * contraction of the fine pass on an internal side stream, capped to this many CTAs, so that it overlaps the MMA-bound
 * backward of the coarse pass (which runs on the remaining SMs); joined again before the call's last launch. */
#define NT_OPT_DW_OVERLAP_CTAS 4
int nt_set_option(nt_ctx* ctx, int key, int value);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t nt_launch_count(const nt_ctx* ctx);

/* ---- ray generation: nerf.py:186-197 (pixel -> unit camera dir), :211 (world dir),
 *      and the view-direction half of Encoder.forward nerf.py:135-167 -------------------
 * row/col dev int64[N] (x := row, y := column, nerf.py:343-344); c2w dev [N,4,4] or [N,3,4]
 * (c2w_stride = 16 or 12 floats) or the loader's fp32 pose rows [N,17] = [3x5 | near far]
 * (c2w_stride = 17, loader.py:33 / poses_extract nerf.py:52-67); kinv dev [3,3] (transposed inverse intrinsics, nerf.py:433).
 * rays  dev [N,16]: d_cam(3) | R row-major(9) | T(3) | 0
 * dir_wrd dev [N,3] (may be NULL); dir_enc dev [N,24] = sin/cos features of dir_wrd, index c*8+2l+s. */
int nt_raygen(nt_ctx* ctx, int64_t n, const int64_t* row, const int64_t* col, const float* c2w, int c2w_stride,
              const float* kinv, float* rays, float* dir_wrd, float* dir_enc, void* stream);

/* ---- the two sub-modules as the reference exposes them, on caller-supplied tensors ------------
 * nt_encode: Encoder.forward nerf.py:135-167.  points / dirs dev [total,3] (either may be NULL);
 *   gamma_point dev [total,3,20], gamma_dir dev [total,3,8] with feature (c, 2l+s) = sin / cos(fl(w_l * x_c)).
 * nt_network_forward: Network.forward nerf.py:101-124 on those encodings (flattened channel-major to [total,60] and
 *   [total,24], nerf.py:103-104), fp32 accuracy path.  rgb dev [total,3], sigma dev [total].  Workspace:
 *   nt_mlp_workspace_bytes(ctx, NT_PREC_FP32, total, 1, 0).  The fused tensor-core paths never materialise the
 *   encodings: use nt_mlp_forward for speed. */
int nt_encode(nt_ctx* ctx, int64_t total, const float* points, const float* dirs, float* gamma_point, float* gamma_dir,
              void* stream);
int nt_network_forward(nt_ctx* ctx, int64_t total, const float* enc_point, const float* enc_dir, const float* params,
                       float* rgb, float* sigma, void* ws, size_t ws_bytes, void* stream);

/* ---- coarse samples: nerf.py:288  np.linspace(near, far, Nc) in fp32 -------------------
 * any_step_zero: -1 = decide on this launch's rays (numpy's batch-global branch), 0/1 = forced
 * (multi-GPU shards pass the globally reduced flag).  t_coarse dev [N,Nc]. */
int nt_sample_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int any_step_zero,
                     float* t_coarse, void* stream);

/* ---- batch-global quantities of a ray-SHARDED launch -------------------------------
 * The reference derives two numbers from the whole batch: numpy's linspace switches formula for EVERY ray when any ray
 * has a zero step (nerf.py:288) and resample uses t_coarse[0,1] - t_coarse[0,0] of the batch's FIRST ray for all rays
 * (nerf.py:234).  A shard computes them without a host round trip:
 *   nt_shard_globals_local   out4 dev float[4] = { flag, delta0 if flag == 0, delta0 if flag == 1, 0 } of THIS shard
 *                            (flag = 1.0 if any local ray has a zero step; the delta0 candidates come from the shard's ray 0
 *                            when first_shard != 0 and are -inf otherwise);
 *   [caller]                 element-wise MAX all-reduce of the 4 floats over the ranks (NCCL, 16 bytes);
 *   nt_shard_globals_resolve g4 -> g4[0] = delta0 (the candidate the global flag selects), g4[1] = flag.
 * nt_render_forward then takes any_step_zero = NT_ANY_STEP_ZERO_DEVICE with delta0 = g4 (it reads the flag at
 * delta0[1]), and nt_render_backward / nt_sample_pdf* take delta0 = g4 as before. */
#define NT_ANY_STEP_ZERO_DEVICE (-2)
int nt_shard_globals_local(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, int first_shard,
                           float* out4, void* stream);
int nt_shard_globals_resolve(nt_ctx* ctx, float* g4, void* stream);

/* ---- encode + MLP: net_out nerf.py:200-219 = sample positions, Encoder.forward
 *      nerf.py:135-167, Network.forward nerf.py:101-124 -----------------------------------
 * t dev [N,P]; rays/dir_enc from nt_raygen; params dev flat fp32; packed dev = output of
 * nt_pack_weights WITH THE SAME PRECISION (every precision but NT_PREC_FP32, which takes NULL).  rgb dev [N,P,3],
 * sigma dev [N,P].  NT_PREC_FP16 / TC32 / MIXED are rendering modes: train != 0 returns NT_ERR_UNSUPPORTED.
 * ws/ws_bytes: scratch of at least nt_mlp_workspace_bytes(); train != 0 keeps the activations
 * in `ws` for nt_mlp_backward. */
size_t nt_mlp_workspace_bytes(nt_ctx* ctx, int precision, int64_t n, int p, int train);
size_t nt_packed_weight_bytes(nt_ctx* ctx, int precision);
int nt_pack_weights(nt_ctx* ctx, int precision, const float* params, void* packed, void* stream);
int nt_mlp_forward(nt_ctx* ctx, int precision, int64_t n, int p, const float* t, const float* rays,
                   const float* dir_enc, const float* params, const void* packed, float* rgb, float* sigma,
                   void* ws, size_t ws_bytes, int train, void* stream);
/* autograd of the above (nerf.py:473): rgb dev [N,P,3] = the colour output of the matching nt_mlp_forward call
 * (sigmoid' = y(1-y); the caller owns that buffer anyway); g_rgb dev [N,P,3], g_sigma dev [N,P]; grads dev flat
 * (ACCUMULATED into); g_t dev [N,P] or NULL (written, not accumulated; needed for the fine pass
 * only, because t_fine is not detached, nerf.py:255-259). */
int nt_mlp_backward(nt_ctx* ctx, int precision, int64_t n, int p, const float* t, const float* rays,
                    const float* dir_enc, const float* params, const void* packed, const float* rgb,
                    const float* g_rgb, const float* g_sigma, float* grads, float* g_t, void* ws, size_t ws_bytes,
                    void* stream);

/* Diagnostic (NT_PREC_BF16): as nt_mlp_forward, and dumps the fp32 post-activation output of one tensor-core
 * layer (0..7 trunk nerf.py:85-91, 8 point_info :96, 9 dir_info :98) to dbg dev [N*P,256]. */
int nt_mlp_forward_debug(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                         const float* params, const void* packed, float* rgb, float* sigma, float* dbg, int layer,
                         void* stream);

/* Diagnostic: the bf16 tensor-core GEMM used by the NT_PREC_BF16 training path.  mn_major = 0: a [M][K], b [N][K];
 * 1: a [K][M], b [K][N] (bf16, K multiple of 64, N <= 256).  out_f32 = 0: c bf16 [M][ldc] (optionally masked by
 * mask > 0, bf16 [M][ldmask]); 1: fp32 atomically accumulated into c [M][ldc]. */
int nt_gemm_bf16_debug(nt_ctx* ctx, int mn_major, int m, int n, int k, const void* a, int lda, const void* b, int ldb,
                       void* c, int ldc, int out_f32, const void* mask, int ldmask, void* stream);

/* ---- compositing: get_density nerf.py:263-272 + color_cum nerf.py:274-281 ----------------
 * coarse: delta = (far-near)/Nc for every sample (nerf.py:293).  weights dev [N,Nc], c_out dev [N,3]. */
int nt_composite_coarse(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, const float* rgb,
                        const float* sigma, float* weights, float* c_out, void* stream);
/* The two reference methods on their own (NeRFModel.get_density nerf.py:263-272, NeRFModel.color_cum nerf.py:274-281)
 * for caller-supplied tensors with p samples per ray (multiple of 32, <= 256): delta, sigma, weights dev [N,p];
 * rgb dev [N,p,3]; c_out dev [N,3].  Inclusive prefix sum accumulated in fp64 and rounded per prefix like the CPU. */
int nt_get_density(nt_ctx* ctx, int64_t n, int p, const float* delta, const float* sigma, float* weights, void* stream);
int nt_color_cum(nt_ctx* ctx, int64_t n, int p, const float* weights, const float* rgb, float* c_out, void* stream);
/* fine: nerf.py:302-321 — concatenate coarse|fine, sort EACH of the 5 channels (t,r,g,b,sigma)
 * independently, delta = diff(t) ++ [last], composite.  perm dev uint8 [N,5,Nc+Nf] or NULL
 * (sorted[i] = in[perm[i]], kept for backward); weights dev [N,Nc+Nf] or NULL. */
int nt_composite_fine(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c, const float* sigma_c,
                      const float* t_f, const float* rgb_f, const float* sigma_f, float last, float* c_out,
                      float* weights, uint8_t* perm, void* stream);
/* backward of nt_composite_coarse. g_c dev [N,3]; g_w_ext dev [N,Nc] or NULL (from nt_sample_pdf_backward);
 * outputs g_rgb dev [N,Nc,3], g_sigma dev [N,Nc] (written). */
int nt_composite_coarse_backward(nt_ctx* ctx, int64_t n, const float* near_, const float* far_, const float* rgb,
                                 const float* sigma, const float* g_c, const float* g_w_ext, float* g_rgb,
                                 float* g_sigma, void* stream);
/* backward of nt_composite_fine (same inputs + perm).  Outputs (written): g_rgb_c [N,Nc,3], g_sigma_c [N,Nc],
 * g_rgb_f [N,Nf,3], g_sigma_f [N,Nf], g_t_f [N,Nf]. */
int nt_composite_fine_backward(nt_ctx* ctx, int64_t n, const float* t_c, const float* rgb_c, const float* sigma_c,
                               const float* t_f, const float* rgb_f, const float* sigma_f, float last,
                               const uint8_t* perm, const float* g_c, float* g_rgb_c, float* g_sigma_c,
                               float* g_rgb_f, float* g_sigma_f, float* g_t_f, void* stream);

/* ---- inverse-CDF resampling: resample nerf.py:225-261 -------------------------------------
 * w dev [N,Nc] coarse weights; delta0 dev float[1] or NULL (NULL: t_coarse[0,1]-t_coarse[0,0] of THIS
 * launch's ray 0, nerf.py:234; shards pass the global value).  t_fine dev [N,Nf]; idx dev int32 [N,Nf]
 * or NULL.  An out-of-range index (reference: exit(0), nerf.py:251-253) sets the ctx status flag;
 * read it with nt_check_status. */
int nt_sample_pdf(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                  float* t_fine, int32_t* idx, void* stream);
/* g_t_fine dev [N,Nf] -> g_w dev [N,Nc] (written).  nerf.py:255-259 is differentiable w.r.t. w. */
int nt_sample_pdf_backward(nt_ctx* ctx, int64_t n, const float* t_coarse, const float* w, const float* delta0,
                           const float* g_t_fine, float* g_w, void* stream);
/* Synchronises `stream`, returns NT_ERR_RANGE if any nt_sample_pdf since the last check saw an
 * out-of-range index, and clears the flag. */
int nt_check_status(nt_ctx* ctx, void* stream);

/* ---- loss seed: ray_loss nerf.py:325-331 ------------------------------------------------------
 * loss dev float[1] (written) = sum (Cc-Ct)^2 + sum (Cf-Ct)^2; g_cc/g_cf dev [N,3] = 2(C-Ct). */
int nt_ray_loss(nt_ctx* ctx, int64_t n, const float* c_coarse, const float* c_fine, const float* c_true,
                float* loss, float* g_cc, float* g_cf, void* stream);

/* ---- fused drivers: render_rays nerf.py:286-323 and its autograd --------------------------------
 * One call = raygen -> t_coarse -> MLP -> composite -> sample_pdf -> MLP -> sort+composite.
 * near_/far_ dev [N].  ws from nt_render_workspace_bytes(n, train).  With train != 0 the workspace
 * holds everything nt_render_backward needs and must be passed to it unchanged. */
size_t nt_render_workspace_bytes(nt_ctx* ctx, int precision, int64_t n, int train);
int nt_render_forward(nt_ctx* ctx, int precision, int64_t n, const int64_t* row, const int64_t* col,
                      const float* c2w, int c2w_stride, const float* kinv, const float* near_, const float* far_,
                      const float* params, const void* packed, int any_step_zero, const float* delta0,
                      float* c_coarse, float* c_fine, void* ws, size_t ws_bytes, int train, void* stream);
int nt_render_backward(nt_ctx* ctx, int precision, int64_t n, const float* near_, const float* far_,
                       const float* params, const void* packed, const float* delta0, const float* g_c_coarse,
                       const float* g_c_fine, float* grads, void* ws, size_t ws_bytes, void* stream);

/* ---- optimiser: torch.optim.Adam as configured at nerf.py:425 (betas .9/.999, eps 1e-7, no decay) --
 * One fused kernel over the flat buffers; step is 1-based.  grad_scale multiplies g first (1.0). */
int nt_adam_step(nt_ctx* ctx, int64_t count, float* params, const float* grads, float* m, float* v, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* Fused gradient all-reduce + Adam over NVLink peer memory (multi-GPU replacement of "NCCL all-reduce, then
 * nt_adam_step"): rank_grads is a HOST array of n_ranks device pointers — this rank's flat gradient and the peer-mapped
 * (symmetric-memory) gradients of the other ranks, in RANK ORDER on every rank.  One kernel reads them over
 * NVLink, sums in rank order (bit-identical parameters on all ranks) and applies the update of nt_adam_step.
 * grad_sum_out dev [count] or NULL receives the summed gradient.  The caller provides the cross-rank barriers: all
 * ranks' gradients complete before the call, all ranks' calls complete before any gradient buffer is rewritten. */
int nt_adam_step_allreduce(nt_ctx* ctx, int64_t count, float* params, const float* const* rank_grads, int n_ranks,
                           float* m, float* v, float lr, float beta1, float beta2, float eps, int64_t step,
                           float grad_scale, float* grad_sum_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFTINY_H */
