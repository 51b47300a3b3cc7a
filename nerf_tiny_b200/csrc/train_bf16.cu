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
  float* rgb;
  bf16 *GI, *GS[8], *Gu, *Gz, *Gzs, *WT;  // GI = d point_info, GS[i] = d pre-activation of trunk layer i
  float *gzsig, *genc;
  void* WB;  // transposed weights in the fused backward-data kernel's chunk format
  size_t bytes;
};

// transposed-weight pack (bf16, each [N][K] K-major for the dX GEMMs)
constexpr int WT_TRUNK = 0;                       // i = 1..7 : (W_i[:, :256])^T  [256][256] each
constexpr int WT_INFO = 7 * 65536;                // W_p^T [256][256]
constexpr int WT_DIRINFO = WT_INFO + 65536;       // (W_d[:, 24:280])^T [256][128]
constexpr int WT_ENC4 = WT_DIRINFO + 256 * 128;   // (W_4[:, 256:316])^T [64][256] (rows 60..63 zero)
constexpr int WT_ENC0 = WT_ENC4 + 64 * 256;       // W_0^T [64][256]
constexpr int WT_ELEMS = WT_ENC0 + 64 * 256;

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
  w.rgb = (float*)take(S * 12);
  w.GI = (bf16*)take(S * 256 * 2);
  for (int i = 0; i < 8; ++i) w.GS[i] = (bf16*)take(S * 256 * 2);
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
  float v = 0.f;
  if (gid < WT_INFO) {  // trunk layers 1..7: out[n = in col][k = out row]
    const int i = 1 + gid / 65536, r = gid % 65536, nn = r / 256, k = r % 256;
    v = params[a.w_off[i] + (int64_t)k * a.in_f[i] + nn];
  } else if (gid < WT_DIRINFO) {
    const int r = gid - WT_INFO, nn = r / 256, k = r % 256;
    v = params[a.w_off[L_INFO] + k * 256 + nn];
  } else if (gid < WT_ENC4) {
    const int r = gid - WT_DIRINFO, nn = r / 128, k = r % 128;  // n = info column, k = dir_info output
    v = params[a.w_off[L_DIR] + k * 280 + 24 + nn];
  } else if (gid < WT_ENC0) {
    const int r = gid - WT_ENC4, nn = r / 256, k = r % 256;
    v = nn < 60 ? params[a.w_off[L_P4] + k * 316 + 256 + nn] : 0.f;
  } else {
    const int r = gid - WT_ENC0, nn = r / 256, k = r % 256;
    v = nn < 60 ? params[a.w_off[L_P0] + k * 60 + nn] : 0.f;
  }
  wt[gid] = __float2bfloat16_rn(v);
}

// head activations backward (B.6) + the colour layer's input gradient + the dir_info bias gradient.
// Phase 1: one thread per sample (sigmoid' / abs', fp32 bias gradients of the two heads by warp reduction).
// Phase 2: the warp walks its 32 samples together, lane = 4 columns of the 128-wide row, so the u row is read and the
// g_u row written as one contiguous 256 B access per sample; each lane keeps the fp32 column sums of what it produced
// (= db of dir_info), reduced over the block's 4 warps in shared memory and added with one atomic per column per block.
__global__ void __launch_bounds__(128) heads_backward_kernel(int64_t S, const float* __restrict__ rgb,
                                                             const float* __restrict__ zsig, const float* __restrict__ g_rgb,
                                                             const float* __restrict__ g_sigma, const bf16* __restrict__ U,
                                                             const float* __restrict__ Wc, bf16* __restrict__ Gz,
                                                             bf16* __restrict__ Gzs, float* __restrict__ gzsig,
                                                             bf16* __restrict__ Gu, float* __restrict__ db_col,
                                                             float* __restrict__ db_sig, float* __restrict__ db_dir) {
  __shared__ float part[4][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = s < S;
  float gz[3] = {0.f, 0.f, 0.f};
  float gs = 0.f;
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = rgb[s * 3 + c];
      gz[c] = g_rgb[s * 3 + c] * y * (1.f - y);
    }
    const float z = zsig[s];
    gs = z > 0.f ? g_sigma[s] : (z < 0.f ? -g_sigma[s] : 0.f);
  }
  {  // bias gradients of the two heads from the fp32 values (the sums cancel heavily: do not round to bf16 first)
    float r0 = gz[0], r1 = gz[1], r2 = gz[2], r3 = gs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r0 += __shfl_xor_sync(0xffffffffu, r0, o);
      r1 += __shfl_xor_sync(0xffffffffu, r1, o);
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      r3 += __shfl_xor_sync(0xffffffffu, r3, o);
    }
    if (lane == 0) {
      atomicAdd(db_col + 0, r0);
      atomicAdd(db_col + 1, r1);
      atomicAdd(db_col + 2, r2);
      atomicAdd(db_sig, r3);
    }
  }
  if (live) {
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
  // g_u = relu'(u) * (g_z . W_c)   (nerf.py:98-99): this lane's 4 columns of the colour weights stay in registers
  float wc[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) wc[c][k] = __ldg(Wc + c * 128 + lane * 4 + k);  // flat parameters are only 4-byte aligned
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t s_warp = (int64_t)blockIdx.x * blockDim.x + warp * 32;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const float z0 = __shfl_sync(0xffffffffu, gz[0], i), z1 = __shfl_sync(0xffffffffu, gz[1], i),
                z2 = __shfl_sync(0xffffffffu, gz[2], i);
    const int64_t si = s_warp + i;
    if (si >= S) break;  // uniform across the warp
    const uint2 u2 = *reinterpret_cast<const uint2*>(U + si * 128 + lane * 4);
    const uint32_t uh[4] = {u2.x & 0xffffu, u2.x >> 16, u2.y & 0xffffu, u2.y >> 16};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g[k] = z0 * wc[0][k] + z1 * wc[1][k] + z2 * wc[2][k];
      if (!(uh[k] != 0 && uh[k] < 0x8000u)) g[k] = 0.f;  // u > 0 as a bf16 bit pattern
      cs[k] += g[k];
    }
    uint2 o2;
    o2.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[1])) << 16);
    o2.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g[3])) << 16);
    *reinterpret_cast<uint2*>(Gu + si * 128 + lane * 4) = o2;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) part[warp][lane * 4 + k] = cs[k];
  __syncthreads();
  const int col = threadIdx.x;
  atomicAdd(db_dir + col, (part[0][col] + part[1][col]) + (part[2][col] + part[3][col]));
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
  NT_TRY(nt_mlp_tc_forward_stash(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, &w.st, st));
  NT_CUDA(cudaMemcpyAsync(w.rgb, rgb, (size_t)S * 12, cudaMemcpyDeviceToDevice, st));
  return NT_OK;
}

int nt_mlp_bf16_train_backward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* params,
                               const float* rgb_unused, const float* g_rgb, const float* g_sigma, float* G, float* g_t,
                               void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)rgb_unused;
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

  {
    TransposeArgs a;
    for (int i = 0; i < NT_N_LAYERS; ++i) {
      a.w_off[i] = (int)T.w[i];
      a.in_f[i] = kLayerIn[i];
    }
    pack_transposed_kernel<<<(WT_ELEMS + 255) / 256, 256, 0, st>>>(P, w.WT, a);
    NT_LAUNCH_CHECK(ctx);
  }
  heads_backward_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(S, w.rgb, w.st.zsig, g_rgb, g_sigma, H[9],
                                                                     P + T.w[L_COLOR], w.Gz, w.Gzs, w.gzsig, w.Gu,
                                                                     G + T.b[L_COLOR], G + T.b[L_SIGMA], G + T.b[L_DIR]);
  NT_LAUNCH_CHECK(ctx);

  // every weight gradient of this pass is queued and computed by ONE grouped tensor-core launch at the end
  NT_TRY(nt_dw_group_begin(S));
  auto dW = [&](const bf16* Gm, int ldg, int M, const bf16* Hm, int ldh, int N, float* dst, int ldc) {
    return nt_dw_group_add(Gm, ldg, M, Hm, ldh, N, dst, ldc);
  };
  // dX also accumulates the column sums of what it stores = the bias gradient of the layer that produced `mask`
  auto dX = [&](const bf16* Gm, int ldg, int K, const bf16* WTm, int N, bf16* out, const bf16* mask, const float* r1_row,
                const float* r1_col, float* db) {
    GemmTcEpi e = epi0();
    e.C = out;
    e.ldc = N;
    e.mask = mask;
    e.ldmask = N;
    e.r1_row = r1_row;
    e.r1_col = r1_col;
    e.colsum = db;
    return nt_launch_gemm_tc(ctx, 0, S, N, K, Gm, ldg, WTm, K, e, st);
  };

  // colour layer (128 -> 3)
  NT_TRY(dW(w.Gz, 8, 3, H[9], 128, 128, G + T.w[L_COLOR], 128));
  // dir_info on [dir_enc | point_info] (nerf.py:118)
  NT_TRY(dW(w.Gu, 128, 128, DENC, 64, 24, G + T.w[L_DIR], 280));
  NT_TRY(dW(w.Gu, 128, 128, H[8], 256, 256, G + T.w[L_DIR] + 24, 280));
  // fused backward-data chain: g_u -> g_info -> g_7 .. g_0 in one tcgen05 kernel (bwd_tc.cu); also the bias gradients
  {
    NT_TRY(nt_bwd_tc_pack(ctx, P, w.WB, st));
    void* outs[9] = {w.GI, w.GS[7], w.GS[6], w.GS[5], w.GS[4], w.GS[3], w.GS[2], w.GS[1], w.GS[0]};
    float* dbs[9] = {G + T.b[L_INFO], G + T.b[L_P7], G + T.b[L_P6], G + T.b[L_P5], G + T.b[L_P4],
                     G + T.b[L_P3],   G + T.b[L_P2], G + T.b[L_P1], G + T.b[L_P0]};
    NT_TRY(nt_bwd_tc_chain(ctx, S, w.Gu, outs, w.WB, w.st.bits, w.gzsig, P + T.w[L_SIGMA], dbs, st));
  }
  // weight gradients (queued): point_info, sigma head, trunk
  NT_TRY(dW(w.GI, 256, 256, H[7], 256, 256, G + T.w[L_INFO], 256));
  NT_TRY(dW(w.Gzs, 8, 1, H[7], 256, 256, G + T.w[L_SIGMA], 256));
  for (int i = 7; i >= 1; --i) {
    bf16* cur = w.GS[i];
    NT_TRY(dW(cur, 256, 256, H[i - 1], 256, 256, G + T.w[i], kLayerIn[i]));
    if (i == 4) {
      NT_TRY(dW(cur, 256, 256, ENC, 64, 60, G + T.w[L_P4] + 256, 316));
      if (g_t) {  // fp32 gradient w.r.t. the xyz features through the skip connection
        GemmTcEpi e = epi0();
        e.C = w.genc;
        e.ldc = 64;
        e.store_f32 = 1;
        NT_TRY(nt_launch_gemm_tc(ctx, 0, S, 64, 256, cur, 256, w.WT + WT_ENC4, 256, e, st));
      }
    }
  }
  NT_TRY(dW(w.GS[0], 256, 256, ENC, 64, 60, G + T.w[L_P0], 60));
  if (g_t) {
    GemmTcEpi e = epi0();
    e.C = w.genc;
    e.ldc = 64;
    e.atomic_f32 = 1;  // accumulate onto the skip-connection part
    NT_TRY(nt_launch_gemm_tc(ctx, 0, S, 64, 256, w.GS[0], 256, w.WT + WT_ENC0, 256, e, st));
    NT_TRY(nt_launch_encode_backward(ctx, n, p, t, rays, w.genc, 64, g_t, st));
  }
  NT_TRY(nt_dw_group_flush(ctx, st));
  return NT_OK;
}
