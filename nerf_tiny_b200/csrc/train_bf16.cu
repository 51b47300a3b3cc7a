// NT_PREC_BF16 training path of the encode+MLP: fused tcgen05 forward that stashes bf16 activations, and a
// layer-major backward made of bf16 tensor-core GEMMs (gemm_tc.cu).  Autograd of Network.forward (nerf.py:101-124)
// + Encoder.forward (nerf.py:135-167) as loss.backward() does it (nerf.py:473); closed forms SURVEY.md B.4, B.6, B.7.
//   dX:  G_{i-1} = relu'(h_{i-1}) * (G_i . W_i)          K-major GEMM, B = pre-transposed weight, mask epilogue
//   dW:  dW_i   += G_i^T . h_{i-1}                        MN-major GEMM over the sample axis, split-K fp32 atomics
//   db:  db_i   += colsum(G_i)
// Gradients w.r.t. the xyz features stay fp32 (they are contracted with w_l up to 3217 in encode_backward).
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

struct Ws {
  TcStash st;
  bf16 *GI, *GS[8], *Gu, *Gz, *Gzs, *WT;  // GI = d point_info, GS[i] = d pre-activation of trunk layer i
  int ldgs[8];                             // row pitch of GS[i]: g_4 and g_0 share one [S][512] buffer (columns 0-255 | 256-511)
  float *gzsig, *genc;
  void* WB;  // transposed weights in the fused backward-data kernel's chunk format
  size_t bytes;
};

// transposed weights of the two layers that read the xyz features, for the ONE GEMM that produces the fp32 gradient
// w.r.t. those features:  g_enc [S][64] = [g_4 | g_0] [S][512] . WT^T,  WT [64 (feature n, rows 60..63 zero)][512]:
// k < 256 -> W_4[k][256 + n] (skip connection, nerf.py:109), k >= 256 -> W_0[k - 256][n]
constexpr int WT_ELEMS = 64 * 512;

Ws carve(void* base, int64_t S) {
  Ws w;
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) {
    void* p = (void*)(b + off);
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  for (int i = 0; i < 9; ++i) w.st.layer[i] = take(S * 256 * 2);
  w.st.layer[9] = take(S * 128 * 2);
  w.st.enc = take(S * 64 * 2);
  w.st.denc = take(S * 64 * 2);
  w.st.zsig = (float*)take(S * 4);
  w.GI = (bf16*)take(S * 256 * 2);
  bf16* g40 = (bf16*)take(S * 512 * 2);
  for (int i = 0; i < 8; ++i) {
    w.ldgs[i] = (i == 4 || i == 0) ? 512 : 256;
    w.GS[i] = i == 4 ? g40 : (i == 0 ? g40 + 256 : (bf16*)take(S * 256 * 2));
  }
  w.Gu = (bf16*)take(S * 128 * 2);
  w.Gz = (bf16*)take(S * 8 * 2);
  w.Gzs = (bf16*)take(S * 8 * 2);
  w.gzsig = (float*)take(S * 4);
  w.genc = (float*)take(S * 64 * 4);
  w.WT = (bf16*)take((size_t)WT_ELEMS * 2);
  w.WB = take(nt_bwd_tc_packed_bytes());
  w.st.bits = (uint32_t*)take((size_t)S * 8 * 8 * 4);
  w.bytes = off;
  return w;
}

struct TransposeArgs {
  int w_off[NT_N_LAYERS];
  int in_f[NT_N_LAYERS];
};

__global__ void pack_transposed_kernel(const float* __restrict__ params, bf16* __restrict__ wt, TransposeArgs a) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= WT_ELEMS) return;
  const int n = gid / 512, k = gid % 512;
  float v = 0.f;
  if (n < 60) v = k < 256 ? params[a.w_off[L_P4] + k * 316 + 256 + n] : params[a.w_off[L_P0] + (k - 256) * 60 + n];
  wt[gid] = __float2bfloat16_rn(v);
}

// head activations backward (B.6) + the colour layer's input gradient + the dir_info bias gradient, HBM-bound:
// 256 B (u) read + 256 B (g_u) written per sample plus ~70 B of head values.
// A block owns HB_ROWS consecutive samples.  Phase 1: one thread per sample evaluates sigmoid' / abs' (fp32), leaves
// g_z in shared memory and feeds the two head-bias gradients (fp32 block sums: a bf16 column sum would cancel
// catastrophically on the sigma bias).  Phase 2: thread = (row r of 16, 8-column group cg of 16): 16-byte loads of u,
// g_u = relu'(u) * (g_z . W_c) (nerf.py:98-99), 16-byte stores; every thread keeps the fp32 column sums of what it
// produced (= db of dir_info), reduced over the 16 row-threads in shared memory -> one atomic per column per block.
constexpr int HB_ROWS = 256;
__global__ void __launch_bounds__(256) heads_backward_kernel(int64_t S, const float* __restrict__ rgb,
                                                             const float* __restrict__ zsig, const float* __restrict__ g_rgb,
                                                             const float* __restrict__ g_sigma, const bf16* __restrict__ U,
                                                             const float* __restrict__ Wc, bf16* __restrict__ Gz,
                                                             bf16* __restrict__ Gzs, float* __restrict__ gzsig,
                                                             bf16* __restrict__ Gu, float* __restrict__ db_col,
                                                             float* __restrict__ db_sig, float* __restrict__ db_dir) {
  __shared__ float s_gz[HB_ROWS][3];
  __shared__ float s_part[16][128];
  __shared__ float s_head[8][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * HB_ROWS;
  {
    const int64_t s = s0 + tid;
    float gz[3] = {0.f, 0.f, 0.f}, gs = 0.f;
    if (s < S) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float y = rgb[s * 3 + c];
        gz[c] = g_rgb[s * 3 + c] * y * (1.f - y);
      }
      const float z = zsig[s];
      gs = z > 0.f ? g_sigma[s] : (z < 0.f ? -g_sigma[s] : 0.f);
      gzsig[s] = gs;
      uint4 o;
      o.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gz[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gz[1])) << 16);
      o.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gz[2]));
      o.z = o.w = 0u;
      *reinterpret_cast<uint4*>(Gz + s * 8) = o;
      o.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gs));
      o.y = 0u;
      *reinterpret_cast<uint4*>(Gzs + s * 8) = o;
    }
    s_gz[tid][0] = gz[0];
    s_gz[tid][1] = gz[1];
    s_gz[tid][2] = gz[2];
    float r0 = gz[0], r1 = gz[1], r2 = gz[2], r3 = gs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r0 += __shfl_xor_sync(0xffffffffu, r0, o);
      r1 += __shfl_xor_sync(0xffffffffu, r1, o);
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      r3 += __shfl_xor_sync(0xffffffffu, r3, o);
    }
    if (lane == 0) {
      s_head[warp][0] = r0;
      s_head[warp][1] = r1;
      s_head[warp][2] = r2;
      s_head[warp][3] = r3;
    }
  }
  __syncthreads();
  if (tid < 4) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_head[w][tid];
    atomicAdd(tid < 3 ? db_col + tid : db_sig, t);
  }
  const int cg = tid & 15, r = tid >> 4;  // columns cg*8 .. +7, rows r, r+16, ...
  float wc[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int k = 0; k < 8; ++k) wc[c][k] = __ldg(Wc + c * 128 + cg * 8 + k);  // flat parameters are only 4-byte aligned
  }
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll 4
  for (int i = 0; i < HB_ROWS / 16; ++i) {
    const int rr = i * 16 + r;
    const int64_t s = s0 + rr;
    if (s >= S) break;
    const uint4 u4 = *reinterpret_cast<const uint4*>(U + s * 128 + cg * 8);
    const float z0 = s_gz[rr][0], z1 = s_gz[rr][1], z2 = s_gz[rr][2];
    const uint32_t uw[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t ow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 2 * e + h;
        const uint32_t ub = h ? (uw[e] >> 16) : (uw[e] & 0xffffu);
        float v = z0 * wc[0][k] + z1 * wc[1][k] + z2 * wc[2][k];
        if (!(ub != 0 && ub < 0x8000u)) v = 0.f;  // u > 0 as a bf16 bit pattern
        cs[k] += v;
        g[h] = v;
      }
      ow[e] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[1])) << 16);
    }
    *reinterpret_cast<uint4*>(Gu + s * 128 + cg * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_part[r][cg * 8 + k] = cs[k];
  __syncthreads();
  if (tid < 128) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += s_part[q][tid];
    atomicAdd(db_dir + tid, t);
  }
}

GemmTcEpi epi0() {
  GemmTcEpi e;
  memset(&e, 0, sizeof(e));
  return e;
}

}  // namespace

#define NT_TRY(x)                   \
  do {                              \
    int rc__ = (x);                 \
    if (rc__ != NT_OK) return rc__; \
  } while (0)

size_t nt_mlp_bf16_train_workspace_bytes(int64_t n, int p) { return carve(nullptr, n * p).bytes; }

int nt_mlp_bf16_train_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                              const float* params, const void* packed, float* rgb, float* sigma, void* ws, size_t ws_bytes,
                              cudaStream_t st) {
  const int64_t S = n * p;
  Ws w = carve(ws, S);
  if (ws_bytes < w.bytes) {
    nt_set_error("bf16 train workspace too small: have %zu need %zu", ws_bytes, w.bytes);
    return NT_ERR_WORKSPACE;
  }
  // the colour head's backward needs rgb: the caller keeps the forward's rgb buffer alive and passes it to the backward
  return nt_mlp_tc_forward_stash(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, &w.st, st);
}

int nt_mlp_bf16_train_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* params,
                               const float* rgb, const float* g_rgb, const float* g_sigma, float* G, float* g_t,
                               void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!rgb) {
    nt_set_error("bf16 backward needs the forward pass's rgb output");
    return NT_ERR_INVALID;
  }
  const int64_t S64 = n * p;
  if (S64 == 0) return NT_OK;
  if (S64 > (int64_t)1 << 30) {
    nt_set_error("bf16 train batch too large");
    return NT_ERR_INVALID;
  }
  const int S = (int)S64;
  Ws w = carve(ws, S);
  if (ws_bytes < w.bytes) {
    nt_set_error("bf16 train workspace too small for backward");
    return NT_ERR_WORKSPACE;
  }
  const LayerTable T = nt_layers();
  const float* P = params;
  bf16** H = reinterpret_cast<bf16**>(w.st.layer);  // H[0..7] trunk, H[8] point_info, H[9] u
  const bf16* ENC = (const bf16*)w.st.enc;
  const bf16* DENC = (const bf16*)w.st.denc;

  if (g_t) {
    TransposeArgs a;
    for (int i = 0; i < NT_N_LAYERS; ++i) {
      a.w_off[i] = (int)T.w[i];
      a.in_f[i] = kLayerIn[i];
    }
    pack_transposed_kernel<<<(WT_ELEMS + 255) / 256, 256, 0, st>>>(P, w.WT, a);
    NT_LAUNCH_CHECK(ctx);
  }
  heads_backward_kernel<<<(unsigned)((S + HB_ROWS - 1) / HB_ROWS), 256, 0, st>>>(S, rgb, w.st.zsig, g_rgb, g_sigma, H[9],
                                                                     P + T.w[L_COLOR], w.Gz, w.Gzs, w.gzsig, w.Gu,
                                                                     G + T.b[L_COLOR], G + T.b[L_SIGMA], G + T.b[L_DIR]);
  NT_LAUNCH_CHECK(ctx);

  // every weight gradient of this pass is queued and computed by ONE grouped tensor-core launch at the end
  NT_TRY(nt_dw_group_begin(S));
  auto dW = [&](const bf16* Gm, int ldg, int M, const bf16* Hm, int ldh, int N, float* dst, int ldc) {
    return nt_dw_group_add(Gm, ldg, M, Hm, ldh, N, dst, ldc);
  };
  // colour layer (128 -> 3)
  NT_TRY(dW(w.Gz, 8, 3, H[9], 128, 128, G + T.w[L_COLOR], 128));
  // dir_info on [dir_enc | point_info] (nerf.py:118)
  NT_TRY(dW(w.Gu, 128, 128, DENC, 64, 24, G + T.w[L_DIR], 280));
  NT_TRY(dW(w.Gu, 128, 128, H[8], 256, 256, G + T.w[L_DIR] + 24, 280));
  // fused backward-data chain: g_u -> g_info -> g_7 .. g_0 in one tcgen05 kernel (bwd_tc.cu); also the bias gradients
  {
    const void* WB = w.WB;
    if (ctx->share_wb && ctx->shared_wb) {
      WB = ctx->shared_wb;  // packed by the fine pass of this nt_render_backward call (same parameters, same stream)
    } else {
      NT_TRY(nt_bwd_tc_pack(ctx, P, w.WB, st));
      if (ctx->share_wb) ctx->shared_wb = w.WB;
    }
    void* outs[9] = {w.GI, w.GS[7], w.GS[6], w.GS[5], w.GS[4], w.GS[3], w.GS[2], w.GS[1], w.GS[0]};
    float* dbs[9] = {G + T.b[L_INFO], G + T.b[L_P7], G + T.b[L_P6], G + T.b[L_P5], G + T.b[L_P4],
                     G + T.b[L_P3],   G + T.b[L_P2], G + T.b[L_P1], G + T.b[L_P0]};
    const int lds[9] = {256, w.ldgs[7], w.ldgs[6], w.ldgs[5], w.ldgs[4], w.ldgs[3], w.ldgs[2], w.ldgs[1], w.ldgs[0]};
    // while a deferred weight-gradient launch of the previous pass is still running on the side stream, this MMA-bound
    // kernel takes only the SMs that launch leaves free
    const int cap = ctx->dw_pending ? ctx->sm_count - ctx->opt_dw_overlap_ctas : 0;
    NT_TRY(nt_bwd_tc_chain(ctx, S, w.Gu, outs, lds, WB, cap, w.st.bits, w.gzsig, P + T.w[L_SIGMA], dbs, st));
    if (ctx->defer_dw) NT_CUDA(cudaEventRecord(ctx->ev_fork, st));  // every operand of this pass's dW problems is complete here
  }
  // weight gradients (queued): point_info, sigma head, trunk
  NT_TRY(dW(w.GI, 256, 256, H[7], 256, 256, G + T.w[L_INFO], 256));
  NT_TRY(dW(w.Gzs, 8, 1, H[7], 256, 256, G + T.w[L_SIGMA], 256));
  for (int i = 7; i >= 1; --i) {
    NT_TRY(dW(w.GS[i], w.ldgs[i], 256, H[i - 1], 256, 256, G + T.w[i], kLayerIn[i]));
    if (i == 4) NT_TRY(dW(w.GS[4], w.ldgs[4], 256, ENC, 64, 60, G + T.w[L_P4] + 256, 316));
  }
  NT_TRY(dW(w.GS[0], w.ldgs[0], 256, ENC, 64, 60, G + T.w[L_P0], 60));
  if (g_t) {
    // fp32 gradient w.r.t. the xyz features, skip connection and first layer in ONE GEMM over [g_4 | g_0] (K = 512), then
    // down to t through the encoder (B.7)
    // the encoder backward runs in the GEMM's epilogue (each thread holds half of a sample's 60 feature gradients): the fp32
    // [S][64] feature gradient never goes to HBM (256 B written + read per sample and one launch less)
    GemmTcEpi e = epi0();
    e.enc_t = t;
    e.enc_rays = rays;
    e.enc_p = p;
    e.enc_gt = g_t;
    NT_CUDA(cudaMemsetAsync(g_t, 0, (size_t)S * sizeof(float), st));
    NT_TRY(nt_launch_gemm_tc(ctx, 0, S, 64, 512, w.GS[4], 512, w.WT, 512, e, st));
  }
  if (ctx->defer_dw) {
    // fork: the HBM-bound contraction of this (fine) pass runs on the side stream with a capped grid, next to the
    // coarse pass's backward on `st`; nt_join_deferred_dw() orders the rest of `st` behind it
    NT_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    NT_TRY(nt_dw_group_flush(ctx, ctx->side, ctx->opt_dw_overlap_ctas));
    NT_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
    ctx->dw_pending = 1;
    return NT_OK;
  }
  NT_TRY(nt_join_deferred_dw(ctx, st));  // one weight-gradient launch at a time: both are HBM-bound
  NT_TRY(nt_dw_group_flush(ctx, st, 0));
  return NT_OK;
}

int nt_join_deferred_dw(nt_ctx* ctx, cudaStream_t st) {
  if (ctx->dw_pending) {
    NT_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    ctx->dw_pending = 0;
  }
  return NT_OK;
}
