// NT_PREC_TC32: the fp32-tolerance mode of the fused encode + 8x256 MLP forward ON THE TENSOR CORES (sm_100a).
// Reference: net_out nerf.py:200-219 -> Encoder.forward nerf.py:135-167 -> Network.forward nerf.py:101-124 (fp32).
//
// Every fp32 operand x is carried as two fp16 numbers  x = hi + 2^-12 * lo,  hi = fp16(x), lo = fp16((x - hi) * 2^12)
// (the residual is scaled so that it never falls into fp16's subnormal range; hi + 2^-12 lo reproduces x to ~2^-24).
// A layer is three tcgen05.mma.kind::f16 passes into TWO fp32 TMEM accumulators:
//     acc1  = A_hi . W_hi                       (exact fp16 products, fp32 accumulation)
//     acc2  = A_lo . W_hi + A_hi . W_lo         (the two cross terms; the dropped lo.lo term is ~2^-24 relative)
//     out   = acc1 + 2^-12 * acc2 + bias
// One persistent CTA per SM works on one 128-sample tile at a time: the hi and lo operand tiles take the shared memory the
// bf16 kernel uses for its two tiles, and acc1 / acc2 take its two TMEM accumulators (columns 0-255 / 256-511).  Weights
// are pre-split (nt_pack_weights) into [hi chunk | lo chunk] pairs and streamed through the same 2-stage TMA ring: a hi
// chunk feeds two MMA groups (A_hi and A_lo), the lo chunk one.  sin/cos features use the accurate sincosf; the sigma and
// colour heads are evaluated on CUDA cores from the fp32 values, as in the bf16 kernel.
// Cost: 3 MMAs per product -> the roofline of this mode is one third of the 16-bit tensor-pipe peak.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace mlpl;
using namespace tcptx;

constexpr int OFF_ACT_HI = 0;
constexpr int OFF_ACT_LO = ACT_BYTES;
constexpr int OFF_ENC_HI = 2 * ACT_BYTES;
constexpr int OFF_ENC_LO = OFF_ENC_HI + CHUNK_A_BYTES;
static_assert(OFF_ENC_LO + CHUNK_A_BYTES == OFF_W, "tc32 reuses the bf16 kernel's shared-memory budget");
constexpr int EPI_THREADS = 512;
enum { B_W_FULL = 0, B_W_EMPTY = 2, B_ACC_FULL = 4, B_ACT_READY = 5, B_AUX_FULL = 6 };
constexpr float LO_SCALE = 4096.f, LO_INV = 1.f / 4096.f;
constexpr int PACKED32_W_BYTES = 2 * PACKED_W_BYTES;
constexpr int PACKED32_BYTES = PACKED32_W_BYTES + AUX_FLOATS * 4;

struct Tc32Params {
  const float* t;
  const float* rays;
  const float* dir_enc;
  const uint8_t* packed;
  float* rgb;
  float* sigma;
  int64_t total;
  int p, p_shift;
  int num_tiles;
};


// x (two values) -> packed fp16 hi pair and packed fp16 scaled-residual pair
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16(x0, x1);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16((x0 - h.x) * LO_SCALE, (x1 - h.y) * LO_SCALE);
}

enum { EPI_RELU = 0, EPI_RELU_SIGMA = 1, EPI_LINEAR = 2, EPI_COLOUR = 3 };

// feature pairs 8*G .. 8*G+7 (pair index pi = c*10 + l) of one sample, accurate sincosf, split into hi / lo fp16 parts.  G is
// a template parameter so that every (c, l) is a compile-time constant after unrolling: straight-line independent chains
// instead of one branchy, serially dependent block per pair with run-time table look-ups (as in mlp_tc.cu::encode_group)
template <int G>
__device__ __forceinline__ void encode_group32(const float (&pos)[3], uint32_t (&fh)[8], uint32_t (&fl)[8]) {
  constexpr uint32_t kFreq[10] = NT_FREQ_POINT_INIT;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pi = G * 8 + i;
    if (pi < 30) {
      const int c = pi / 10, l = pi % 10;
      float sn, cs;
      sincosf(__fmul_rn(__uint_as_float(kFreq[l]), pos[c]), &sn, &cs);
      split2(sn, cs, fh[i], fl[i]);  // features (c*20+2l, c*20+2l+1)
    } else {
      fh[i] = fl[i] = 0u;  // K padded 60 -> 64
    }
  }
}

// epilogue of one layer for one thread: row `row` of the tile (= one TMEM lane), column quarter `quarter`
template <int KIND>
__device__ __forceinline__ void epilogue32(const Tc32Params& P, int quarter, uint32_t tmem_row, uint32_t act_hi, uint32_t act_lo,
                                           uint32_t aux_s, int quad_bar, const RowSwz sw, int64_t s, bool valid) {
  constexpr int NCB = KIND == EPI_COLOUR ? 1 : 2;  // 32-column blocks per quarter
  const int cb0 = quarter * NCB;
  float sig_acc = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCB; ++i) {
    const int cb = cb0 + i;
    uint32_t r1[32], r2[32];
    tmem_ld32_issue(tmem_row + cb * 32, r1);
    tmem_ld32_issue(tmem_row + 256 + cb * 32, r2);
    tmem_ld_wait(r1);
    tmem_ld_wait(r2);
    const uint32_t bias_s = aux_s + cb * 128;
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = lds128(bias_s + j * 16);
      v[4 * j + 0] = fmaf(__uint_as_float(r2[4 * j + 0]), LO_INV, __uint_as_float(r1[4 * j + 0])) + b4.x;
      v[4 * j + 1] = fmaf(__uint_as_float(r2[4 * j + 1]), LO_INV, __uint_as_float(r1[4 * j + 1])) + b4.y;
      v[4 * j + 2] = fmaf(__uint_as_float(r2[4 * j + 2]), LO_INV, __uint_as_float(r1[4 * j + 2])) + b4.z;
      v[4 * j + 3] = fmaf(__uint_as_float(r2[4 * j + 3]), LO_INV, __uint_as_float(r1[4 * j + 3])) + b4.w;
    }
    if (KIND != EPI_LINEAR) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);  // nerf.py:86-91, :98 (ReLU); point_info has none (:117)
    }
    if (KIND == EPI_RELU_SIGMA) {  // sigma head from the fp32 activations (nerf.py:94, :114)
      const uint32_t ws = aux_s + (AUX_EXTRA + cb * 32) * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w4 = lds128(ws + j * 16);
        sig_acc = fmaf(v[4 * j + 0], w4.x, sig_acc);
        sig_acc = fmaf(v[4 * j + 1], w4.y, sig_acc);
        sig_acc = fmaf(v[4 * j + 2], w4.z, sig_acc);
        sig_acc = fmaf(v[4 * j + 3], w4.w, sig_acc);
      }
    }
    if (KIND == EPI_COLOUR) {  // colour head (nerf.py:99, :119) on u = relu(.)
      const uint32_t wc = aux_s + (AUX_EXTRA + cb * 32) * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w0 = lds128(wc + j * 16), w1 = lds128(wc + 512 + j * 16), w2 = lds128(wc + 1024 + j * 16);
        c0 = fmaf(v[4 * j], w0.x, fmaf(v[4 * j + 1], w0.y, fmaf(v[4 * j + 2], w0.z, fmaf(v[4 * j + 3], w0.w, c0))));
        c1 = fmaf(v[4 * j], w1.x, fmaf(v[4 * j + 1], w1.y, fmaf(v[4 * j + 2], w1.z, fmaf(v[4 * j + 3], w1.w, c1))));
        c2 = fmaf(v[4 * j], w2.x, fmaf(v[4 * j + 1], w2.y, fmaf(v[4 * j + 2], w2.z, fmaf(v[4 * j + 3], w2.w, c2))));
      }
    } else {  // next layer's A operand (hi and lo tiles): K-chunk cb/2, 16-byte chunks (cb%2)*4 .. +3
      const uint32_t dh = act_hi + (cb >> 1) * CHUNK_A_BYTES, dl = act_lo + (cb >> 1) * CHUNK_A_BYTES;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(v[8 * qd + 2 * e], v[8 * qd + 2 * e + 1], h[e], l[e]);
        st_shared_v4(sw.addr(dh, (cb & 1) * 4 + qd), h[0], h[1], h[2], h[3]);
        st_shared_v4(sw.addr(dl, (cb & 1) * 4 + qd), l[0], l[1], l[2], l[3]);
      }
    }
  }
  // heads: the 4 column quarters of a row live in warps q, q+4, q+8, q+12 (same TMEM lanes): quarters 1-3 park their
  // partial sums in acc1 columns they have already drained, quarter 0 adds them up after a 128-thread named barrier
  if (KIND == EPI_RELU_SIGMA || KIND == EPI_COLOUR) {
    constexpr int QW = KIND == EPI_COLOUR ? 32 : 64;  // columns per quarter
    if (quarter != 0) tmem_st4(tmem_row + quarter * QW, sig_acc, c0, c1, c2);
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");
    tc_fence_after();
    if (quarter == 0) {
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        float o_s, o0, o1, o2;
        tmem_ld4(tmem_row + q * QW, o_s, o0, o1, o2);
        sig_acc += o_s;
        c0 += o0;
        c1 += o1;
        c2 += o2;
      }
      if (KIND == EPI_RELU_SIGMA) {
        if (valid) P.sigma[s] = fabsf(sig_acc + lds32(aux_s + AUX_SIG_B * 4));  // nerf.py:74
      } else if (valid) {
        P.rgb[s * 3 + 0] = 1.f / (1.f + expf(-(c0 + lds32(aux_s + AUX_COL_B * 4))));
        P.rgb[s * 3 + 1] = 1.f / (1.f + expf(-(c1 + lds32(aux_s + AUX_COL_B * 4 + 4))));
        P.rgb[s * 3 + 2] = 1.f / (1.f + expf(-(c2 + lds32(aux_s + AUX_COL_B * 4 + 8))));
      }
    }
  }
}

__global__ void __launch_bounds__(N_THREADS, 1) mlp_tc32_kernel(const __grid_constant__ Tc32Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();  // SWIZZLE_128B operands need a 1024-byte aligned base
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(bar(B_W_FULL + s), 1);
      mbar_init(bar(B_W_EMPTY + s), 1);
    }
    mbar_init(bar(B_ACC_FULL), 1);
    mbar_init(bar(B_ACT_READY), EPI_THREADS);
    mbar_init(bar(B_AUX_FULL), 1);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc_512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ===================== TMA producer: [hi | lo] chunk pairs, same order for every tile =====================
    if (lane == 0) {
      uint32_t q = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const uint8_t* src = P.packed;
        for (int L = 0; L < N_MMA_LAYERS; ++L) {
          const uint32_t bytes = chunk_bytes(L);
          for (int kc = 0; kc < 2 * layer_chunks(L); ++kc, ++q) {
            const uint32_t stage = q & 1;
            mbar_wait(bar(B_W_EMPTY + stage), ((q >> 1) & 1) ^ 1);
            mbar_expect_tx(bar(B_W_FULL + stage), bytes);
            tma_bulk_g2s(sbase + OFF_W + stage * W_STAGE_BYTES, src, bytes, bar(B_W_FULL + stage));
            src += bytes;
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer: three passes per K-chunk into acc1 (cols 0-255) / acc2 (cols 256-511) ============
    if (lane == 0) {
      uint32_t q = 0, lit = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        for (int L = 0; L < N_MMA_LAYERS; ++L, ++lit) {
          const int nch = layer_chunks(L);
          const uint32_t idesc = idesc_f16kind(TILE_M, layer_n(L), false);
          mbar_wait(bar(B_ACT_READY), lit & 1);  // hi/lo operand tiles written, both accumulators drained
          tc_fence_after();
          // every epilogue thread is done with the previous layer's record: stage this layer's biases / head weights
          mbar_expect_tx(bar(B_AUX_FULL), aux_bytes(L));
          tma_bulk_g2s(sbase + OFF_AUX, P.packed + PACKED32_W_BYTES + L * (AUX_REC_FLOATS * 4), aux_bytes(L), bar(B_AUX_FULL));
          for (int kc = 0; kc < nch; ++kc) {
            const bool from_enc = (L == 0) || (kc == 4);
            const uint32_t a_hi = from_enc ? sbase + OFF_ENC_HI : sbase + OFF_ACT_HI + kc * CHUNK_A_BYTES;
            const uint32_t a_lo = from_enc ? sbase + OFF_ENC_LO : sbase + OFF_ACT_LO + kc * CHUNK_A_BYTES;
            {  // W_hi chunk: acc1 += A_hi . W_hi ; acc2 += A_lo . W_hi
              const uint32_t stage = q & 1;
              mbar_wait(bar(B_W_FULL + stage), (q >> 1) & 1);
              tc_fence_after();
              const uint32_t b_addr = sbase + OFF_W + stage * W_STAGE_BYTES;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(tmem_base, desc_kmajor(a_hi + j * 32), desc_kmajor(b_addr + j * 32), idesc, (kc | j) != 0);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(tmem_base + 256, desc_kmajor(a_lo + j * 32), desc_kmajor(b_addr + j * 32), idesc, (kc | j) != 0);
              umma_commit(bar(B_W_EMPTY + stage));
              ++q;
            }
            {  // W_lo chunk: acc2 += A_hi . W_lo
              const uint32_t stage = q & 1;
              mbar_wait(bar(B_W_FULL + stage), (q >> 1) & 1);
              tc_fence_after();
              const uint32_t b_addr = sbase + OFF_W + stage * W_STAGE_BYTES;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(tmem_base + 256, desc_kmajor(a_hi + j * 32), desc_kmajor(b_addr + j * 32), idesc, 1u);
              umma_commit(bar(B_W_EMPTY + stage));
              ++q;
            }
          }
          umma_commit(bar(B_ACC_FULL));
        }
      }
    }
  } else {
    // ===================== 16 encode + epilogue warps: 4 lane quadrants x 4 column quarters =====================
    const int quad = warp & 3, quarter = warp >> 2;
    const int row = quad * 32 + lane;
    const int quad_bar = 1 + quad;
    RowSwz sw;
    sw.row_off = row * 128;
    sw.x4 = (row & 7) << 4;
    const uint32_t act_hi = sbase + OFF_ACT_HI, act_lo = sbase + OFF_ACT_LO;
    const uint32_t enc_hi = sbase + OFF_ENC_HI, enc_lo = sbase + OFF_ENC_LO;
    const uint32_t aux_s = sbase + OFF_AUX;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int64_t s = (int64_t)tile * TILE_M + row;
      const bool valid = s < P.total;
      const int64_t sc = valid ? s : P.total - 1;
      const int64_t ray = P.p_shift >= 0 ? (sc >> P.p_shift) : sc / P.p;
      {  // ---- positional encoding (nerf.py:200-216, 135-167): feature pairs 8*quarter .. +7 of this row, hi and lo tiles
        const float4* rp = reinterpret_cast<const float4*>(P.rays + ray * 16);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3);
        const float tt = __ldg(P.t + sc);
        const float pc0 = __fmul_rn(r0.x, tt), pc1 = __fmul_rn(r0.y, tt), pc2 = __fmul_rn(r0.z, tt);
        float pos[3];
        pos[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.w, pc0), __fmul_rn(r1.x, pc1)), __fmul_rn(r1.y, pc2)), r3.x);
        pos[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.z, pc0), __fmul_rn(r1.w, pc1)), __fmul_rn(r2.x, pc2)), r3.y);
        pos[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.y, pc0), __fmul_rn(r2.z, pc1)), __fmul_rn(r2.w, pc2)), r3.z);
        uint32_t fh[8], fl[8];
        switch (quarter) {  // warp-uniform; the group index must be a compile-time constant (see encode_group32)
          case 0: encode_group32<0>(pos, fh, fl); break;
          case 1: encode_group32<1>(pos, fh, fl); break;
          case 2: encode_group32<2>(pos, fh, fl); break;
          default: encode_group32<3>(pos, fh, fl); break;
        }
        st_shared_v4(sw.addr(enc_hi, quarter * 2), fh[0], fh[1], fh[2], fh[3]);
        st_shared_v4(sw.addr(enc_hi, quarter * 2 + 1), fh[4], fh[5], fh[6], fh[7]);
        st_shared_v4(sw.addr(enc_lo, quarter * 2), fl[0], fl[1], fl[2], fl[3]);
        st_shared_v4(sw.addr(enc_lo, quarter * 2 + 1), fl[4], fl[5], fl[6], fl[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(B_ACT_READY));

      for (int L = 0; L < N_MMA_LAYERS; ++L, ++it) {
        mbar_wait(bar(B_ACC_FULL), it & 1);
        tc_fence_after();
        mbar_wait(bar(B_AUX_FULL), it & 1);
        if (L == 7)
          epilogue32<EPI_RELU_SIGMA>(P, quarter, tmem_row, act_hi, act_lo, aux_s, quad_bar, sw, s, valid);
        else if (L == 8)
          epilogue32<EPI_LINEAR>(P, quarter, tmem_row, act_hi, act_lo, aux_s, quad_bar, sw, s, valid);
        else if (L == 9)
          epilogue32<EPI_COLOUR>(P, quarter, tmem_row, act_hi, act_lo, aux_s, quad_bar, sw, s, valid);
        else
          epilogue32<EPI_RELU>(P, quarter, tmem_row, act_hi, act_lo, aux_s, quad_bar, sw, s, valid);
        if (L == 4) {
          // all MMAs that read the xyz features have retired: the tiles now take the view-direction features
          // (24 features = 16-byte chunks 0-2, zero chunks 3-7), quarter 0 -> hi tile, quarter 1 -> lo tile
          if (quarter < 2) {
            const uint32_t enc = quarter == 0 ? enc_hi : enc_lo;
            const float4* de = reinterpret_cast<const float4*>(P.dir_enc + ray * 24);
            uint32_t f[12];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const float4 d4 = __ldg(de + j);
              uint32_t h0, l0, h1, l1;
              split2(d4.x, d4.y, h0, l0);
              split2(d4.z, d4.w, h1, l1);
              f[2 * j] = quarter == 0 ? h0 : l0;
              f[2 * j + 1] = quarter == 0 ? h1 : l1;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) st_shared_v4(sw.addr(enc, j), f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
#pragma unroll
            for (int j = 3; j < 8; ++j) st_shared_v4(sw.addr(enc, j), 0u, 0u, 0u, 0u);
          }
        }
        if (L != 9) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          tc_fence_before();
          mbar_arrive(bar(B_ACT_READY));
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc_512(tmem_base);
  }
}

}  // namespace

size_t nt_mlp_tc32_packed_bytes() { return PACKED32_BYTES; }

int nt_mlp_tc32_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                        const void* packed, float* rgb, float* sigma, cudaStream_t st) {
  if (!(ctx->attr_done & NT_ATTR_MLP_TC32)) {
    NT_CUDA(cudaFuncSetAttribute(mlp_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ctx->attr_done |= NT_ATTR_MLP_TC32;
  }
  Tc32Params P;
  memset(&P, 0, sizeof(P));
  P.t = t;
  P.rays = rays;
  P.dir_enc = dir_enc;
  P.packed = (const uint8_t*)packed;
  P.rgb = rgb;
  P.sigma = sigma;
  P.total = n * p;
  P.p = p;
  P.p_shift = -1;
  for (int b = 0; b < 30; ++b)
    if (p == (1 << b)) P.p_shift = b;
  const int64_t tiles = (P.total + TILE_M - 1) / TILE_M;
  if (tiles == 0) return NT_OK;
  if (tiles > 0x7fffffff) {
    nt_set_error("tc32: too many samples in one launch");
    return NT_ERR_INVALID;
  }
  P.num_tiles = (int)tiles;
  const int grid = ctx->sm_count < P.num_tiles ? ctx->sm_count : P.num_tiles;
  mlp_tc32_kernel<<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
