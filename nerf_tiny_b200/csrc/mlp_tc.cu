// NT_PREC_BF16: fused positional-encoding + 8x256 MLP forward on the 5th-gen tensor cores (sm_100a).
// Reference: net_out nerf.py:200-219 -> Encoder.forward nerf.py:135-167 -> Network.forward nerf.py:101-124.
//
// One persistent CTA per SM works on PAIRS of 128-sample tiles (A, B):
//   * sin/cos features are generated in registers by the tile's 4 epilogue warps and stored as bf16 straight into
//     the 128B-swizzled K-major A-operand layout in shared memory (never touch HBM);
//   * every layer's bf16 weights are pre-packed (nt_pack_weights) as ready-to-use swizzled [N x 64] K-chunks and
//     streamed by a TMA bulk-copy producer warp through a 2-stage shared-memory ring; each chunk is consumed by
//     BOTH tiles before it is released (weights cross L2->SMEM once per 256 samples);
//   * one elected thread issues tcgen05.mma (M=128, N=256|128, K=16, bf16 -> fp32) accumulating in TMEM
//     (tile A: columns 0-255, tile B: 256-511);
//   * the epilogue warps read the accumulator with tcgen05.ld (double-buffered), add bias, apply ReLU, round to
//     bf16 and write the next layer's A operand in place; the sigma head (256->1, abs) and colour head (128->3,
//     sigmoid) are evaluated on CUDA cores from the fp32 accumulators in the same pass.
// Skip (nerf.py:109) and view (nerf.py:118) concatenations are extra K-chunks accumulated into the same TMEM tile.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "mlp_layout.cuh"

namespace {

using namespace mlpl;

struct TcParams {
  const float* t;
  const float* rays;
  const float* dir_enc;
  const float* params;
  const uint8_t* packed;
  float* rgb;
  float* sigma;
  float* dbg;  // optional [S,256] fp32 dump of one layer's post-activation output
  int dbg_layer;
  int64_t total;  // samples
  int p;          // samples per ray
  int p_shift;    // log2(p) when p is a power of two, else -1
  int num_pairs;
  // lock-step kernel: work items.  Items below full_pairs are tile pairs (tiles 2i, 2i+1); the remaining items are SINGLE
  // tiles (tile 2*full_pairs + (i - full_pairs)) - used when the last wave of pairs would leave more than half of the CTAs
  // idle (the 1024-ray training batch: 512 fine-pass pairs = 3 waves of 148 + 68 pairs -> 136 single tiles on 136 CTAs)
  int num_items, full_pairs;
  // training stash (STASH instantiation): bf16 activations kept in HBM for the layer-major backward
  __nv_bfloat16* st[N_MMA_LAYERS];  // post-activation output of every tensor-core layer, [S][256] ([S][128] for dir_info)
  __nv_bfloat16* st_enc;            // xyz features [S][64] (60 + zero pad)
  __nv_bfloat16* st_denc;           // view features [S][64] (24 + zero pad)
  float* st_zsig;                   // sigma pre-activation [S] (abs' needs the sign)
  // v5 STASH instantiation: the operand tiles already sit in shared memory in the SWIZZLE_128B layout, so the stash
  // of layers 0..8, the xyz features and the view features is written by TMA tensor stores (box 64 cols x 128 rows),
  // not by the epilogue threads.  maps 0..8 = st[0..8] ([S][256]), 9 = st_enc, 10 = st_denc ([S][64]).
  CUtensorMap st_map[11];
  int use_tma_stash;
  uint32_t* st_bits;  // ReLU' bit masks of the 8 trunk layers, [8][S][8 words]: bit j of word cb = column cb*32+j > 0
};


// PTX wrappers: tc_ptx.cuh (shared with bwd_tc.cu / gemm_tc.cu / mlp_tc32.cu)
using namespace tcptx;

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return desc_kmajor(saddr); }
// instruction descriptor: D=f32, A=B=bf16 (F16 = false) or fp16, both K-major, M=128
template <bool F16 = false>
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return idesc_f16kind(TILE_M, n, !F16); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  umma_f16(d_tmem, adesc, bdesc, idesc, acc);
}

// sin/cos for the bf16 path: 2-term Cody-Waite reduction by 2*pi, then the SFU (abs err ~1e-6, far below bf16's 4e-3)
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
  // round-to-nearest by the 1.5 * 2^23 trick (|x / 2pi| < 2^22 here): two FADD-class ops instead of an XU-pipe FRND
  const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.f), -12582912.f);
  float r = fmaf(k, -6.2831854820251465f, x);
  r = fmaf(k, 1.7484555314695172e-07f, r);
  s = __sinf(r);
  c = __cosf(r);
}

// sin/cos feature pairs NP*G .. NP*G+NP-1 (pair index pi = c*10 + l, nerf.py:135-167) of one sample, packed to 16-bit.
// G is a template parameter so that after unrolling every (c, l) is a compile-time constant: the frequencies become
// immediates and the NP chains are independent straight-line code (with a run-time group index the compiler emitted one
// branchy, serially dependent block per pair with spilled index tables: ~3 000 clk per tile instead of ~1 000)
template <int NP, int G, bool F16>
__device__ __forceinline__ void encode_group(const float (&pos)[3], uint32_t (&f)[NP]) {
  constexpr uint32_t kFreq[10] = NT_FREQ_POINT_INIT;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int pi = G * NP + i;
    if (pi < 30) {
      const int c = pi / 10, l = pi % 10;
      float sn, cs;
      fast_sincos(__fmul_rn(__uint_as_float(kFreq[l]), pos[c]), sn, cs);
      f[i] = F16 ? pack_f16(sn, cs) : pack_bf16(sn, cs);  // features (c*20+2l, c*20+2l+1)
    } else {
      f[i] = 0u;  // K padded 60 -> 64
    }
  }
}

enum { EPI_RELU = 0, EPI_RELU_SIGMA = 1, EPI_LINEAR = 2, EPI_COLOUR = 3 };

// diagnostic timeline (debug instantiation only, dbg_layer >= 100): clock64 stamps of block 0's first 4 tile pairs,
// prof[(pair_local*10 + L)*16 + slot] as int64 inside the dbg buffer
#define TC_PROF(slot)                                                                                        \
  do {                                                                                                       \
    if (DBG && P.dbg_layer >= 100 && blockIdx.x == 0 && pair_local < 4)                                      \
      reinterpret_cast<long long*>(P.dbg)[(pair_local * 10 + L) * 16 + (slot)] = clock64();                  \
  } while (0)

// One layer's epilogue for one thread: the thread owns one sample (= one TMEM lane) and one half of the layer's
// output columns.  accumulator -> +bias -> activation -> bf16 A operand of the next layer (or the heads).
// 16 epilogue warps (4 per scheduler) hide the TMEM latency by thread-level parallelism; biases and head weights
// come from the layer's record in shared memory (broadcast LDS, staged by TMA while the MMAs run).
template <int KIND, bool DBG, bool STASH>
__device__ __forceinline__ void epilogue_layer(const TcParams& P, int L, int half, uint32_t tmem_row, uint32_t act,
                                               uint32_t aux_s, int pair_bar, const RowSwz sw, int64_t s, bool valid) {
  __nv_bfloat16* stp = STASH ? P.st[L] + s * (KIND == EPI_COLOUR ? 128 : 256) : nullptr;
  constexpr int NCB = KIND == EPI_COLOUR ? 2 : 4;  // 32-column blocks per half
  const int cb0 = half * NCB;
  float sig_acc = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  uint32_t mbits[4] = {0u, 0u, 0u, 0u};
  uint32_t buf[2][32];  // double-buffered accumulator blocks: the next TMEM load is in flight during the math
  tmem_ld32_issue(tmem_row + cb0 * 32, buf[0]);
#pragma unroll
  for (int i = 0; i < NCB; ++i) {
    const int cb = cb0 + i;
    uint32_t(&raw)[32] = buf[i & 1];
    tmem_ld_wait(raw);
    if (i + 1 < NCB) tmem_ld32_issue(tmem_row + (cb + 1) * 32, buf[(i + 1) & 1]);
    const uint32_t bias_s = aux_s + cb * 128;
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b4[j] = lds128(bias_s + j * 16);
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j + 0] = __uint_as_float(raw[4 * j + 0]) + b4[j].x;
      v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b4[j].y;
      v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b4[j].z;
      v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b4[j].w;
    }
    if (STASH && (KIND == EPI_RELU || KIND == EPI_RELU_SIGMA)) {
      // 1 bit per activation is all the backward-data kernel needs of this layer: funnel-shift the 32 sign bits of the
      // pre-activations into one word (one SHF each); element j ends up at bit 31-j, set = ReLU passes
      uint32_t sgn = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) sgn = __funnelshift_l(__float_as_uint(v[j]), sgn, 1);
      mbits[i] = ~sgn;
    }
    if (DBG) {
      if (P.dbg_layer == L && valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) P.dbg[s * 256 + cb * 32 + j] = (KIND == EPI_LINEAR) ? v[j] : fmaxf(v[j], 0.f);
      }
    }
    if (KIND == EPI_RELU_SIGMA) {  // sigma head from the fp32 activations (nerf.py:94, :114)
      const uint32_t ws = aux_s + (AUX_EXTRA + cb * 32) * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w4 = lds128(ws + j * 16);
        sig_acc = fmaf(fmaxf(v[4 * j + 0], 0.f), w4.x, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 1], 0.f), w4.y, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 2], 0.f), w4.z, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 3], 0.f), w4.w, sig_acc);
      }
    }
    if (KIND == EPI_COLOUR) {  // colour head (nerf.py:99, :119) on u = relu(.)
      const uint32_t wc = aux_s + (AUX_EXTRA + cb * 32) * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w0 = lds128(wc + j * 16), w1 = lds128(wc + 512 + j * 16), w2 = lds128(wc + 1024 + j * 16);
        const float u0 = fmaxf(v[4 * j], 0.f), u1 = fmaxf(v[4 * j + 1], 0.f), u2 = fmaxf(v[4 * j + 2], 0.f),
                    u3 = fmaxf(v[4 * j + 3], 0.f);
        c0 = fmaf(u0, w0.x, fmaf(u1, w0.y, fmaf(u2, w0.z, fmaf(u3, w0.w, c0))));
        c1 = fmaf(u0, w1.x, fmaf(u1, w1.y, fmaf(u2, w1.z, fmaf(u3, w1.w, c1))));
        c2 = fmaf(u0, w2.x, fmaf(u1, w2.y, fmaf(u2, w2.z, fmaf(u3, w2.w, c2))));
      }
      if (STASH && valid) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(stp + cb * 32 + qd * 8) =
              make_uint4(pack_bf16_relu(v[8 * qd], v[8 * qd + 1]), pack_bf16_relu(v[8 * qd + 2], v[8 * qd + 3]),
                         pack_bf16_relu(v[8 * qd + 4], v[8 * qd + 5]), pack_bf16_relu(v[8 * qd + 6], v[8 * qd + 7]));
      }
    } else {  // next layer's A operand: K-chunk cb/2, 16-byte chunks (cb%2)*4 .. +3
      const uint32_t dst = act + (cb >> 1) * CHUNK_A_BYTES;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = KIND == EPI_LINEAR ? pack_bf16(v[8 * qd + 2 * e], v[8 * qd + 2 * e + 1])
                                    : pack_bf16_relu(v[8 * qd + 2 * e], v[8 * qd + 2 * e + 1]);
        st_shared_v4(sw.addr(dst, (cb & 1) * 4 + qd), w[0], w[1], w[2], w[3]);
        if (STASH && !P.use_tma_stash && valid)
          *reinterpret_cast<uint4*>(stp + cb * 32 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  if (STASH && (KIND == EPI_RELU || KIND == EPI_RELU_SIGMA) && valid)
    *reinterpret_cast<uint4*>(P.st_bits + ((int64_t)L * P.total + s) * 8 + cb0) = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
  // ---- heads: the two column halves of a row live in warps w and w+8 (same TMEM lanes).  The upper half parks its
  // partial sums in accumulator columns it has already drained; the lower half picks them up after a 64-thread
  // named barrier.  Nothing overwrites those columns before both halves arrive on act_ready.
  if (KIND == EPI_RELU_SIGMA || KIND == EPI_COLOUR) {
    const uint32_t park = tmem_row + (KIND == EPI_COLOUR ? 64 : 128);
    if (half == 1) tmem_st4(park, sig_acc, c0, c1, c2);
    tc_fence_before();
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    tc_fence_after();
    if (half == 0) {
      float o_s, o0, o1, o2;
      tmem_ld4(park, o_s, o0, o1, o2);
      if (KIND == EPI_RELU_SIGMA) {
        const float z = sig_acc + o_s + lds32(aux_s + AUX_SIG_B * 4);
        if (valid) P.sigma[s] = fabsf(z);
        if (STASH && valid) P.st_zsig[s] = z;
      } else if (valid) {
        P.rgb[s * 3 + 0] = 1.f / (1.f + __expf(-(c0 + o0 + lds32(aux_s + AUX_COL_B * 4))));
        P.rgb[s * 3 + 1] = 1.f / (1.f + __expf(-(c1 + o1 + lds32(aux_s + AUX_COL_B * 4 + 4))));
        P.rgb[s * 3 + 2] = 1.f / (1.f + __expf(-(c2 + o2 + lds32(aux_s + AUX_COL_B * 4 + 8))));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
template <bool DBG, bool STASH>
__global__ void __launch_bounds__(N_THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();  // SWIZZLE_128B operands need a 1024-byte aligned base
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(bar(BAR_W_FULL + s), 1);
      mbar_init(bar(BAR_W_EMPTY + s), 1);
    }
    for (int tl = 0; tl < 2; ++tl) {
      mbar_init(bar(BAR_ACC_FULL + tl), 1);
      mbar_init(bar(BAR_ACT_READY + tl), 2 * TILE_M);
    }
    mbar_init(bar(BAR_AUX_FULL), 1);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc_512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ===================== TMA producer: weight K-chunks, same order for every tile pair =====================
    if (lane == 0) {
      uint32_t q = 0;
      const uint64_t pol_w = l2_policy_evict_last();  // STASH: weights stay L2-resident under the stash stream (tc_ptx.cuh)
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
        const uint8_t* src = P.packed;
        for (int L = 0; L < N_MMA_LAYERS; ++L) {
          const uint32_t bytes = chunk_bytes(L);
          for (int kc = 0; kc < layer_chunks(L); ++kc, ++q) {
            const uint32_t stage = q & 1;
            mbar_wait(bar(BAR_W_EMPTY + stage), ((q >> 1) & 1) ^ 1);
            mbar_expect_tx(bar(BAR_W_FULL + stage), bytes);
            if (STASH)
              tma_bulk_g2s_hint(sbase + OFF_W + stage * W_STAGE_BYTES, src, bytes, bar(BAR_W_FULL + stage), pol_w);
            else
              tma_bulk_g2s(sbase + OFF_W + stage * W_STAGE_BYTES, src, bytes, bar(BAR_W_FULL + stage));
            src += bytes;
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer: one thread drives the tensor core for both tiles =====================
    if (lane == 0) {
      uint32_t q = 0, lit = 0;
      int pair_local = 0;
      const uint64_t pol_st = l2_policy_evict_first();  // stash lines leave L2 first: they are read once, much later
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++pair_local) {
        // a single-tile item is always this CTA's LAST item, so tile B's barrier phases may fall behind `lit` from here on
        const bool single = item >= P.full_pairs;
        const int tile0 = single ? 2 * P.full_pairs + (item - P.full_pairs) : 2 * item;
        const int ntl = single ? 1 : 2;
        for (int L = 0; L < N_MMA_LAYERS; ++L, ++lit) {
          const int nch = layer_chunks(L);
          const uint32_t idesc = umma_idesc(layer_n(L));
          for (int kc = 0; kc < nch; ++kc, ++q) {
            const uint32_t stage = q & 1;
            mbar_wait(bar(BAR_W_FULL + stage), (q >> 1) & 1);
            tc_fence_after();
            const uint32_t b_addr = sbase + OFF_W + stage * W_STAGE_BYTES;
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
              if (tl >= ntl) continue;
              if (kc == 0) {  // A operand written + accumulator drained by the tile's epilogue warps
                mbar_wait(bar(BAR_ACT_READY + tl), lit & 1);
                tc_fence_after();
                TC_PROF(tl);
                if (STASH && P.use_tma_stash == 3 && L == 0) {
                  // (legacy scheme, NT_STASH_MODE=3) layer 0's operand = the xyz features: stash them from smem
                  const int row0 = (tile0 + tl) * TILE_M;
                  if (row0 < P.total) {
                    tma_store_2d_hint(&P.st_map[9], 0, row0, sbase + OFF_ENC + tl * CHUNK_A_BYTES, pol_st);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  }
                }
                if (tl == ntl - 1) {
                  // every epilogue thread is done with the previous layer's record: stage this layer's biases / head
                  // weights (1-2.5 KB) for the epilogue that follows these MMAs
                  mbar_expect_tx(bar(BAR_AUX_FULL), aux_bytes(L));
                  tma_bulk_g2s(sbase + OFF_AUX, P.packed + PACKED_W_BYTES + L * (AUX_REC_FLOATS * 4), aux_bytes(L),
                               bar(BAR_AUX_FULL));
                }
              }
              const bool from_enc = (L == 0) || (kc == 4);
              const uint32_t a_addr = from_enc ? sbase + OFF_ENC + tl * CHUNK_A_BYTES
                                               : sbase + OFF_ACT + tl * ACT_BYTES + kc * CHUNK_A_BYTES;
              const uint32_t d_tmem = tmem_base + tl * 256;
#pragma unroll
              for (int j = 0; j < 4; ++j)  // 4 x (K = 16) inside the 64-wide swizzled chunk: +32 B per step
                umma_bf16(d_tmem, umma_desc(a_addr + j * 32), umma_desc(b_addr + j * 32), idesc, (kc | j) != 0);
              if (STASH && P.use_tma_stash == 3 && L > 0) {
                // this K-chunk of the operand = 64 columns of the previous layer's bf16 output (chunk 4 of dir_info = the
                // view features): stash it from smem.  One 16 KB box per chunk step, not the whole tile at once, so the
                // weight loads of the next chunks are not queued behind 128 KB of stores in the TMA unit.
                const int row0 = (tile0 + tl) * TILE_M;
                if (row0 < P.total && (kc < 4 || L == 9)) {
                  if (kc < 4)
                    tma_store_2d_hint(&P.st_map[L - 1], kc * 64, row0, a_addr, pol_st);
                  else
                    tma_store_2d_hint(&P.st_map[10], 0, row0, a_addr, pol_st);
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              }
              if (kc == nch - 1) {
                // the epilogue that this commit releases overwrites the operand tiles: TMA stash reads must be done
                if (STASH && P.use_tma_stash == 3) tma_store_wait_read0();
                umma_commit(bar(BAR_ACC_FULL + tl));
              }
            }
            umma_commit(bar(BAR_W_EMPTY + stage));  // frees the ring slot once both tiles' MMAs retire
            if (kc == 0) TC_PROF(2);
            if (kc == nch - 1) TC_PROF(3);
          }
        }
      }
      if (STASH && P.use_tma_stash == 3) tma_store_wait0();
    }
  } else {
    // ===================== encode + epilogue warps: per tile 4 lane quadrants x 2 column halves =============
    const int tl = (warp >> 2) & 1;
    const int half = warp >> 3;
    const int row = (warp & 3) * 32 + lane;
    const int pair_bar = 1 + (warp & 7);  // named barrier shared by the two warps that own the same rows
    const uint32_t act = sbase + OFF_ACT + tl * ACT_BYTES;
    const uint32_t enc = sbase + OFF_ENC + tl * CHUNK_A_BYTES;
    const uint32_t aux_s = sbase + OFF_AUX;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + tl * 256;
    RowSwz sw;
    sw.row_off = row * 128;
    sw.x4 = (row & 7) << 4;
    uint32_t it = 0;
    int pair_local = 0;
    const int pslot = 4 + 3 * (warp >> 2);  // per warp-group stamp slots (lane 0 of warps 0, 4, 8, 12)
    // Training stash (use_tma_stash == 1): ONE elected thread per tile writes every operand tile to HBM with TMA tensor
    // stores right after the tile's 8 epilogue warps have produced it (256-thread named barrier), i.e. while the tensor
    // core multiplies it, and makes sure the stores have finished READING shared memory before the same warps overwrite
    // the tile in the next epilogue.  The MMA thread never waits on the TMA unit.
    const bool tma_stash = STASH && P.use_tma_stash == 1;
    const bool stasher = tma_stash && (warp & 3) == 0 && half == 0 && lane == 0;
    const int tile_bar = 9 + tl;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++pair_local) {
      const bool single = item >= P.full_pairs;
      if (single && tl == 1) break;  // single-tile item (this CTA's last): tile B's warps have nothing to do
      const int tile0 = single ? 2 * P.full_pairs + (item - P.full_pairs) : 2 * item;
      const int row0 = (tile0 + tl) * TILE_M;
      const int64_t s = ((int64_t)tile0 + tl) * TILE_M + row;
      const bool valid = s < P.total;
      const int64_t sc = valid ? s : P.total - 1;
      const int64_t ray = sc / P.p;
      // ---- positional encoding of this sample (nerf.py:200-216, 135-167) -> enc tile, bf16, swizzled ----
      // half 0 produces feature pairs 0..15 (16-byte chunks 0-3), half 1 pairs 16..29 + zero padding (chunks 4-7)
      {
        const float4* rp = reinterpret_cast<const float4*>(P.rays + ray * 16);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3);
        const float tt = __ldg(P.t + sc);
        const float pc0 = __fmul_rn(r0.x, tt), pc1 = __fmul_rn(r0.y, tt), pc2 = __fmul_rn(r0.z, tt);
        float pos[3];
        pos[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.w, pc0), __fmul_rn(r1.x, pc1)), __fmul_rn(r1.y, pc2)), r3.x);
        pos[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.z, pc0), __fmul_rn(r1.w, pc1)), __fmul_rn(r2.x, pc2)), r3.y);
        pos[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.y, pc0), __fmul_rn(r2.z, pc1)), __fmul_rn(r2.w, pc2)), r3.z);
        uint32_t f[16];
        if (half == 0)
          encode_group<16, 0, false>(pos, f);
        else
          encode_group<16, 1, false>(pos, f);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(sw.addr(enc, half * 4 + j), f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        if (STASH && !P.use_tma_stash && valid) {
          uint4* d = reinterpret_cast<uint4*>(P.st_enc + s * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) d[j] = make_uint4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
      fence_proxy_async();
      if (tma_stash) asm volatile("bar.sync %0, 256;" ::"r"(tile_bar) : "memory");
      tc_fence_before();
      mbar_arrive(bar(BAR_ACT_READY + tl));
      if (stasher && row0 < P.total) {  // the xyz features of the tile
        tma_store_2d(&P.st_map[9], 0, row0, enc);
        tma_store_commit();
      }

      for (int L = 0; L < N_MMA_LAYERS; ++L, ++it) {
        mbar_wait(bar(BAR_ACC_FULL + tl), it & 1);
        tc_fence_after();
        mbar_wait(bar(BAR_AUX_FULL), it & 1);
        if (tma_stash) {  // the stash stores of this tile's operands have left shared memory: it may be overwritten
          if (stasher) tma_store_wait_read0();
          asm volatile("bar.sync %0, 256;" ::"r"(tile_bar) : "memory");
        }
        if ((warp & 3) == 0 && lane == 0) TC_PROF(pslot);
        if (L == 7)
          epilogue_layer<EPI_RELU_SIGMA, DBG, STASH>(P, L, half, tmem_row, act, aux_s, pair_bar, sw, s, valid);
        else if (L == 8)
          epilogue_layer<EPI_LINEAR, DBG, STASH>(P, L, half, tmem_row, act, aux_s, pair_bar, sw, s, valid);
        else if (L == 9)
          epilogue_layer<EPI_COLOUR, DBG, STASH>(P, L, half, tmem_row, act, aux_s, pair_bar, sw, s, valid);
        else
          epilogue_layer<EPI_RELU, DBG, STASH>(P, L, half, tmem_row, act, aux_s, pair_bar, sw, s, valid);
        if (L == 4) {
          // all MMAs that read the xyz features have retired: reuse the tile for the view-direction features
          // (24 features = 16-byte chunks 0-2 by half 0; zero chunks 3-7 split between the halves)
          if (half == 0) {
            const float4* de = reinterpret_cast<const float4*>(P.dir_enc + ray * 24);
            uint32_t f[12];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const float4 d4 = __ldg(de + j);
              f[2 * j] = pack_bf16(d4.x, d4.y);
              f[2 * j + 1] = pack_bf16(d4.z, d4.w);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) st_shared_v4(sw.addr(enc, j), f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            st_shared_v4(sw.addr(enc, 3), 0u, 0u, 0u, 0u);
            if (STASH && !P.use_tma_stash && valid) {
              uint4* d = reinterpret_cast<uint4*>(P.st_denc + s * 64);
#pragma unroll
              for (int j = 0; j < 3; ++j) d[j] = make_uint4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
#pragma unroll
              for (int j = 3; j < 8; ++j) d[j] = make_uint4(0u, 0u, 0u, 0u);
            }
          } else {
#pragma unroll
            for (int j = 4; j < 8; ++j) st_shared_v4(sw.addr(enc, j), 0u, 0u, 0u, 0u);
          }
        }
        if ((warp & 3) == 0 && lane == 0) TC_PROF(pslot + 1);
        if (L != 9) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          if (tma_stash) asm volatile("bar.sync %0, 256;" ::"r"(tile_bar) : "memory");
          tc_fence_before();
          mbar_arrive(bar(BAR_ACT_READY + tl));
          if (stasher && row0 < P.total) {
            // this layer's bf16 output = the operand the tensor core reads next: 4 boxes of 64 columns, at most two in
            // flight so that the weight ring's loads are never queued behind more than 32 KB of stores in the TMA unit
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tma_store_2d(&P.st_map[L], c * 64, row0, act + c * CHUNK_A_BYTES);
              tma_store_commit();
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            if (L == 4) {  // the view features just written over the xyz features
              tma_store_2d(&P.st_map[10], 0, row0, enc);
              tma_store_commit();
            }
          }
        }
        if ((warp & 3) == 0 && lane == 0) TC_PROF(pslot + 2);
      }
    }
    if (stasher) tma_store_wait0();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc_512(tmem_base);
  }
}


// =========================================================================================================
// v7: tile-STAGGERED schedule with CTA-PAIR MMAs (tcgen05 cta_group::2).
// As v6 the tensor cores alternate between the two tiles one whole layer at a time so that one tile's epilogue hides
// behind the other tile's MMAs, and each tile makes its own pass over the layer's weight chunks.  v6 showed that a
// single SM cannot ingest a full 32 KB chunk per 512 MMA-cycles; here the two CTAs of a cluster form one 256-row MMA
// (M = 2 x 128 rows, one tile from each CTA) whose B operand is SPLIT across the pair: each CTA streams only its
// N-half of every chunk (16 KB, 4-stage ring).  The leader CTA's elected thread issues tcgen05.mma.cta_group::2; the
// peer CTA's otherwise idle MMA warp relays "operand ready" / "weights landed" to the leader with remote mbarrier
// arrives; completion (tcgen05.commit.cta_group::2 multicast) wakes the epilogue warps and producers of both CTAs.
// =========================================================================================================
namespace v7 {

constexpr int NST = 4;                                // ring stages of 16 KB (this CTA's N-half of a chunk)
// chunk stream of a layer = the loads the producer issues, in ring order.  1- and 4-chunk layers: every chunk is loaded
// once and used by tile A, then tile B.  5-chunk layers (skip / view concatenation): [X, Y0, Y1, Y2, Y3, X] where X is the
// encoding chunk (K-chunk 4): tile A consumes loads 0-4 (X first, released at once), tile B loads 1-5 (X last), so four
// slots still suffice and only X is streamed twice.
__host__ __device__ constexpr int layer_loads(int L) { return layer_chunks(L) == 5 ? 6 : layer_chunks(L); }
__host__ __device__ constexpr int load_chunk(int L, int li) {
  return layer_chunks(L) == 5 ? ((li == 0 || li == 5) ? 4 : li - 1) : li;
}
constexpr int HALF_STAGE = W_STAGE_BYTES / 2;
constexpr int OFF_BIAS = OFF_BAR + 256;               // 2 x 1 KB double-buffered bias rows (layer parity)
constexpr int SMEM7_BYTES = OFF_BIAS + 2048;
enum { B_W_FULL = 0, B_W_EMPTY = 4, B_ACC_FULL = 8, B_ACT_READY = 10, B_BIAS_FULL = 12 };
constexpr int EPI_THREADS = 512;
constexpr int BAR_ID_EPI_ALL = 5;                     // named barrier over all epilogue threads (ids 1-4: row quads)

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// M = 256 (both CTAs' 128 rows), D = f32, A = B = bf16, K-major
template <bool F16>
__host__ __device__ constexpr uint32_t umma2_idesc(int n) { return idesc_f16kind(256, n, !F16); }
// 16-bit operand packing of the schedule's precision: bf16 (NT_PREC_BF16) or saturating fp16 (NT_PREC_FP16)
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) { return F16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
template <bool F16>
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi) { return F16 ? pack_f16_relu(lo, hi) : pack_bf16_relu(lo, hi); }
// shared-memory matrix descriptor (K-major, SWIZZLE_128B, SBO 1024 B) split into its constant high word and the
// address word: advancing K by 16 bf16 (32 B) is +2 on the low word
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | (1u << 16); }
__device__ __forceinline__ void umma2_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI)
      : "memory");
}
// relaxed: the relayed facts (TMA bytes landed / operand rows written and proxy-fenced) were already acquired by the
// relaying thread from its own CTA's barrier; a release.cluster arrive costs ~700 clk per relay (measured)
__device__ __forceinline__ void remote_arrive_raw(uint32_t raddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t local_bar, uint32_t target_cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_bar), "r"(target_cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tmem2_alloc_512(uint32_t smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_slot) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc_512(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}

// epilogue of one tile-layer for one thread: row `row` of the tile, column quarter `quarter`
template <int KIND, bool F16, bool TL = false>
__device__ __forceinline__ void epilogue_q(const TcParams& P, int L, int quarter, uint32_t tmem_row, uint32_t act,
                                           uint32_t bias_s, const float* __restrict__ aux_g, int quad_bar, const RowSwz sw,
                                           int64_t s, bool valid) {
  constexpr int NCB = KIND == EPI_COLOUR ? 1 : 2;  // 32-column blocks per quarter
  const int cb0 = quarter * NCB;
  const int xf = TL ? P.dbg_layer - 100 : 0;  // timeline build only: experiment flags (1: no exchange, 2: no weight loads)
  float sig_acc = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  uint32_t buf[2][32];
  tmem_ld32_issue(tmem_row + cb0 * 32, buf[0]);
#pragma unroll
  for (int i = 0; i < NCB; ++i) {
    const int cb = cb0 + i;
    uint32_t(&raw)[32] = buf[i & 1];
    tmem_ld_wait(raw);
    if (i + 1 < NCB) tmem_ld32_issue(tmem_row + (cb + 1) * 32, buf[(i + 1) & 1]);
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b4[j] = lds128(bias_s + cb * 128 + j * 16);
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // packed fp32 adds (FADD2): half the instructions of the bias add
      add_f32x2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(raw[4 * j + 0]), __uint_as_float(raw[4 * j + 1]), b4[j].x, b4[j].y);
      add_f32x2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]), b4[j].z, b4[j].w);
    }
    if (KIND == EPI_RELU_SIGMA) {  // sigma head (nerf.py:94, :114); weights from the fp32 side block in L1/L2
      const float4* __restrict__ ws = reinterpret_cast<const float4*>(aux_g + 7 * AUX_REC_FLOATS + AUX_EXTRA + cb * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w4 = (TL && (xf & 2)) ? make_float4(0.5f, 0.25f, 0.125f, 1.f) : __ldg(ws + j);
        sig_acc = fmaf(fmaxf(v[4 * j + 0], 0.f), w4.x, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 1], 0.f), w4.y, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 2], 0.f), w4.z, sig_acc);
        sig_acc = fmaf(fmaxf(v[4 * j + 3], 0.f), w4.w, sig_acc);
      }
    }
    if (KIND == EPI_COLOUR) {  // colour head (nerf.py:99, :119) on u = relu(.)
      const float4* __restrict__ wc = reinterpret_cast<const float4*>(aux_g + 9 * AUX_REC_FLOATS + AUX_EXTRA + cb * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool nw = TL && (xf & 2);
        const float4 w0 = nw ? make_float4(0.5f, 0.25f, 0.125f, 1.f) : __ldg(wc + j);
        const float4 w1 = nw ? make_float4(0.25f, 0.5f, 0.125f, 1.f) : __ldg(wc + 32 + j);
        const float4 w2 = nw ? make_float4(0.125f, 0.25f, 0.5f, 1.f) : __ldg(wc + 64 + j);
        const float u0 = fmaxf(v[4 * j], 0.f), u1 = fmaxf(v[4 * j + 1], 0.f), u2 = fmaxf(v[4 * j + 2], 0.f),
                    u3 = fmaxf(v[4 * j + 3], 0.f);
        c0 = fmaf(u0, w0.x, fmaf(u1, w0.y, fmaf(u2, w0.z, fmaf(u3, w0.w, c0))));
        c1 = fmaf(u0, w1.x, fmaf(u1, w1.y, fmaf(u2, w1.z, fmaf(u3, w1.w, c1))));
        c2 = fmaf(u0, w2.x, fmaf(u1, w2.y, fmaf(u2, w2.z, fmaf(u3, w2.w, c2))));
      }
    } else {  // next layer's A operand: K-chunk cb/2 (= this thread's quarter), 16-byte chunks (cb%2)*4 .. +3
      const uint32_t dst = act + (cb >> 1) * CHUNK_A_BYTES;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = KIND == EPI_LINEAR ? pack16<F16>(v[8 * qd + 2 * e], v[8 * qd + 2 * e + 1])
                                    : pack16_relu<F16>(v[8 * qd + 2 * e], v[8 * qd + 2 * e + 1]);
        st_shared_v4(sw.addr(dst, (cb & 1) * 4 + qd), w[0], w[1], w[2], w[3]);
      }
    }
  }
  // heads: the 4 column quarters of a row live in warps q, q+4, q+8, q+12 (same TMEM lanes): quarters 1-3 park their
  // partial sums in accumulator columns they have already drained, quarter 0 adds them up after a 128-thread barrier
  if (KIND == EPI_RELU_SIGMA || KIND == EPI_COLOUR) {
    constexpr int QW = KIND == EPI_COLOUR ? 32 : 64;  // columns per quarter
    const bool nx = TL && (xf & 1);
    if (quarter != 0 && !nx) tmem_st4(tmem_row + quarter * QW, sig_acc, c0, c1, c2);
    if (!nx) {
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");
      tc_fence_after();
    }
    if (quarter == 0) {
      float o[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (!nx) tmem_ld4x3(tmem_row + QW, tmem_row + 2 * QW, tmem_row + 3 * QW, o);
#pragma unroll
      for (int q = 0; q < 3; ++q) {  // same order of additions as before
        sig_acc += o[4 * q];
        c0 += o[4 * q + 1];
        c1 += o[4 * q + 2];
        c2 += o[4 * q + 3];
      }
      if (KIND == EPI_RELU_SIGMA) {
        const float z = sig_acc + __ldg(aux_g + 7 * AUX_REC_FLOATS + AUX_SIG_B);
        if (valid) P.sigma[s] = fabsf(z);
      } else if (valid) {
        const float* cbias = aux_g + 9 * AUX_REC_FLOATS + AUX_COL_B;
        P.rgb[s * 3 + 0] = 1.f / (1.f + __expf(-(c0 + __ldg(cbias))));
        P.rgb[s * 3 + 1] = 1.f / (1.f + __expf(-(c1 + __ldg(cbias + 1))));
        P.rgb[s * 3 + 2] = 1.f / (1.f + __expf(-(c2 + __ldg(cbias + 2))));
      }
    }
  }
}

template <bool F16, bool TL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N_THREADS, 1) mlp_tc7_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 200);
  const uint8_t* aux_bytes_g = P.packed + PACKED_W_BYTES;
  const int iters = (P.num_pairs + (int)gridDim.x - 1) / (int)gridDim.x;  // same trip count in both CTAs of a cluster

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();
    // the leader's "full" / "ready" barriers also count one remote arrive from the peer's MMA warp (its N-half of the
    // chunk has landed / its half of the 256-row operand is ready), so the issuing thread waits on ONE barrier per event
    const uint32_t relay = crank == 0 ? 1u : 0u;
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar(B_W_FULL + s), 1 + relay);
      mbar_init(bar(B_W_EMPTY + s), 1);  // the leader's tcgen05.commit.cta_group::2, multicast to both CTAs
    }
    for (int tl = 0; tl < 2; ++tl) {
      mbar_init(bar(B_ACC_FULL + tl), 1);
      mbar_init(bar(B_ACT_READY + tl), EPI_THREADS + relay);
      mbar_init(bar(B_BIAS_FULL + tl), 1);
    }
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem2_alloc_512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ===================== TMA producer: my N-half of every load of the layer's chunk stream =====================
    if (lane == 0) {
      uint32_t q = 0;
      for (int itp = 0; itp < iters; ++itp) {
        const uint8_t* src = P.packed;
        for (int L = 0; L < N_MMA_LAYERS; ++L) {
          const uint32_t bytes = chunk_bytes(L), hbytes = bytes / 2;
          const int nl = layer_loads(L);
          for (int li = 0; li < nl; ++li, ++q) {
            const int kc = load_chunk(L, li);
            const uint32_t stage = q % NST;
            // polled: for the peer CTA the leader's multicast commit is a remote arrive, which does not wake a thread that is
            // suspended inside try_wait (it would sleep out the ~700 clk time limit before reloading the slot)
            mbar_spin(bar(B_W_EMPTY + stage), ((q / NST) & 1) ^ 1);
            mbar_expect_tx(bar(B_W_FULL + stage), hbytes);
            tma_bulk_g2s(sbase + OFF_W + stage * HALF_STAGE, src + kc * bytes + crank * hbytes, hbytes,
                         bar(B_W_FULL + stage));
          }
          src += layer_chunks(L) * bytes;
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA warp: the leader issues the pair's MMAs, the peer relays readiness ==========
    // Everything per-layer is a compile-time constant (the layer loop is fully unrolled) and descriptors are 32-bit
    // adds: the issuing thread shares its scheduler with four busy epilogue warps, so every instruction between the
    // last MMA of one pass and the first MMA of the next is exposed once the ~4-deep MMA queue has drained.
    if (lane == 0 && crank != 0) {
      uint32_t r_act, r_wfull;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_act) : "r"(bar(B_ACT_READY)), "r"(0));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_wfull) : "r"(bar(B_W_FULL)), "r"(0));
      uint32_t q = 0, lit = 0;
      for (int itp = 0; itp < iters; ++itp) {
#pragma unroll
        for (int L = 0; L < N_MMA_LAYERS; ++L, ++lit) {
          constexpr int dummy = 0;
          (void)dummy;
          const int nch = layer_chunks(L);
          const bool five = nch == 5;
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
            mbar_wait(bar(B_ACT_READY + tl), lit & 1);  // my half of the 256-row operand written, accumulator drained
            remote_arrive_raw(r_act + 8u * tl);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              if (i >= nch || (tl == 1 && !(five && i == 4))) continue;
              const uint32_t qq = q + i + ((five && tl == 1) ? 1 : 0), stage = qq % NST;
              mbar_wait(bar(B_W_FULL + stage), (qq / NST) & 1);   // my N-half of the chunk has landed
              remote_arrive_raw(r_wfull + 8u * stage);
            }
          }
          q += layer_loads(L);
        }
      }
    } else if (lane == 0) {
      const uint32_t bar_act = bar(B_ACT_READY), bar_wf = bar(B_W_FULL), bar_we = bar(B_W_EMPTY), bar_acc = bar(B_ACC_FULL);
      const uint32_t lo_act = desc_lo(sbase + OFF_ACT), lo_enc = desc_lo(sbase + OFF_ENC), lo_w = desc_lo(sbase + OFF_W);
      uint32_t q = 0, lit = 0;
      for (int itp = 0; itp < iters; ++itp) {
        const bool stamp = TL && blockIdx.x == 0 && itp < 4;
#pragma unroll
        for (int L = 0; L < N_MMA_LAYERS; ++L, ++lit) {
          const int nch = layer_chunks(L);
          const bool five = nch == 5;
          const uint32_t idesc = umma2_idesc<F16>(layer_n(L));
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
            // loads used by this pass: tile A = loads 0..nch-1 (all first uses), tile B = the resident chunks again and,
            // in 5-chunk layers, the re-streamed extra chunk (load 5) last
            const uint32_t q0 = q + ((five && tl == 1) ? 1 : 0);
            // poll every barrier of the pass at once (independent test_waits overlap their latency); in steady state the
            // weights have landed long ago and the 16-20 MMAs go out back to back
            uint32_t rdy = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              if (i >= nch) continue;
              const bool need = tl == 0 || (five && i == 4);
              const uint32_t qq = q0 + i;
              rdy |= (need ? mbar_try(bar_wf + 8u * (qq % NST), (qq / NST) & 1) : 1u) << i;
            }
            mbar_spin(bar_act + 8u * tl, lit & 1);  // operand written + accumulator drained in BOTH CTAs
            tc_fence_after();
            if (stamp) reinterpret_cast<long long*>(P.dbg)[(itp * 10 + L) * 16 + tl * 2] = clock64();
            const uint32_t d_tmem = tmem_base + tl * 256;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              if (i >= nch) continue;
              const uint32_t qq = q0 + i, stage = qq % NST;
              if (!((rdy >> i) & 1)) {
                mbar_spin(bar_wf + 8u * stage, (qq / NST) & 1);
                tc_fence_after();
              }
              // K-chunk of the A operand behind this load: 5-chunk layers put the extra (encoding) chunk first for tile A
              const int kc = five ? (tl == 0 ? (i == 0 ? 4 : i - 1) : i) : i;
              const bool from_enc = (L == 0) || (kc == 4);
              const uint32_t a_lo = from_enc ? lo_enc + tl * (CHUNK_A_BYTES >> 4)
                                             : lo_act + tl * (ACT_BYTES >> 4) + kc * (CHUNK_A_BYTES >> 4);
              const uint32_t b_lo = lo_w + stage * (HALF_STAGE >> 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma2_bf16_lo(d_tmem, a_lo + 2 * j, b_lo + 2 * j, idesc, (i | j) != 0);
              if (tl == 1 || (five && i == 0)) umma2_commit_mc(bar_we + 8u * stage, (uint16_t)3);  // slot free in BOTH CTAs
            }
            umma2_commit_mc(bar_acc + 8u * tl, (uint16_t)3);     // wakes both CTAs' epilogue warps
            if (stamp) reinterpret_cast<long long*>(P.dbg)[(itp * 10 + L) * 16 + tl * 2 + 1] = clock64();
          }
          q += layer_loads(L);
        }
      }
    }
  } else {
    // ===================== 16 encode + epilogue warps, serving tile A and tile B alternately ==================
    const int quad = warp & 3, quarter = warp >> 2;
    const int row = quad * 32 + lane;
    const int quad_bar = 1 + quad;
    const float* __restrict__ aux_g = reinterpret_cast<const float*>(aux_bytes_g);
    RowSwz sw;
    sw.row_off = row * 128;
    sw.x4 = (row & 7) << 4;
    const uint32_t tmem_q = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t it = 0;        // per-tile layer counter (acc_full / act_ready parity)
    uint32_t bias_use[2] = {0, 0};
    const bool loader = threadIdx.x == 0;
    if (loader) {  // biases of layers 0 and 1 (1 KB each) into the two buffers
      for (int b = 0; b < 2; ++b) {
        mbar_expect_tx(bar(B_BIAS_FULL + b), 1024);
        tma_bulk_g2s(sbase + OFF_BIAS + b * 1024, aux_bytes_g + b * (AUX_REC_FLOATS * 4), 1024, bar(B_BIAS_FULL + b));
      }
    }
    for (int itp = 0; itp < iters; ++itp) {
      const int pair = blockIdx.x + itp * (int)gridDim.x;
      int64_t s_t[2];
      bool valid_t[2];
      int64_t ray_t[2];
      // ---- positional encoding of both tiles' samples: this thread = one row, feature pairs 8*quarter .. +7 ----
      // the inputs of BOTH tiles are requested first (one exposed memory latency, not two), and the next pair's lines
      // are pulled into L2 a whole pair ahead
      float4 rin[2][4];
      float tin[2];
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) {
        const int64_t s = ((int64_t)pair * 2 + tl) * TILE_M + row;
        const bool valid = pair < P.num_pairs && s < P.total;
        const int64_t sc = valid ? s : P.total - 1;
        const int64_t ray = P.p_shift >= 0 ? (sc >> P.p_shift) : sc / P.p;
        s_t[tl] = s;
        valid_t[tl] = valid;
        ray_t[tl] = ray;
        const float4* rp = reinterpret_cast<const float4*>(P.rays + ray * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) rin[tl][j] = __ldg(rp + j);
        tin[tl] = __ldg(P.t + sc);
      }
      if (quarter == 0 && lane == 0) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          const int64_t sn = ((int64_t)(pair + (int)gridDim.x) * 2 + tl) * TILE_M + quad * 32;
          if (sn < P.total) {
            prefetch_l2(P.t + sn);
            prefetch_l2(P.rays + (P.p_shift >= 0 ? (sn >> P.p_shift) : sn / P.p) * 16);
          }
        }
      }
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) {
        const int64_t s = s_t[tl];
        const bool valid = valid_t[tl];
        const float4 r0 = rin[tl][0], r1 = rin[tl][1], r2 = rin[tl][2], r3 = rin[tl][3];
        const float tt = tin[tl];
        const float pc0 = __fmul_rn(r0.x, tt), pc1 = __fmul_rn(r0.y, tt), pc2 = __fmul_rn(r0.z, tt);
        float pos[3];
        pos[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.w, pc0), __fmul_rn(r1.x, pc1)), __fmul_rn(r1.y, pc2)), r3.x);
        pos[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.z, pc0), __fmul_rn(r1.w, pc1)), __fmul_rn(r2.x, pc2)), r3.y);
        pos[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.y, pc0), __fmul_rn(r2.z, pc1)), __fmul_rn(r2.w, pc2)), r3.z);
        uint32_t f[8];
        switch (quarter) {  // warp-uniform
          case 0: encode_group<8, 0, F16>(pos, f); break;
          case 1: encode_group<8, 1, F16>(pos, f); break;
          case 2: encode_group<8, 2, F16>(pos, f); break;
          default: encode_group<8, 3, F16>(pos, f); break;
        }
        const uint32_t enc = sbase + OFF_ENC + tl * CHUNK_A_BYTES;
        st_shared_v4(sw.addr(enc, quarter * 2), f[0], f[1], f[2], f[3]);
        st_shared_v4(sw.addr(enc, quarter * 2 + 1), f[4], f[5], f[6], f[7]);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar(B_ACT_READY + tl));
      }

      for (int L = 0; L < N_MMA_LAYERS; ++L, ++it) {
        const int bb = L & 1;
        const uint32_t bias_s = sbase + OFF_BIAS + bb * 1024;
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          const uint32_t act = sbase + OFF_ACT + tl * ACT_BYTES;
          const uint32_t enc = sbase + OFF_ENC + tl * CHUNK_A_BYTES;
          const uint32_t tmem_row = tmem_q + tl * 256;
          mbar_wait(bar(B_ACC_FULL + tl), it & 1);
          tc_fence_after();
          if (tl == 0) mbar_wait(bar(B_BIAS_FULL + bb), bias_use[bb] & 1);
          if (TL && blockIdx.x == 0 && itp < 4 && threadIdx.x == 0)
            reinterpret_cast<long long*>(P.dbg)[(itp * 10 + L) * 16 + 4 + tl * 2] = clock64();
          if (L == 7)
            epilogue_q<EPI_RELU_SIGMA, F16, TL>(P, L, quarter, tmem_row, act, bias_s, aux_g, quad_bar, sw, s_t[tl], valid_t[tl]);
          else if (L == 8)
            epilogue_q<EPI_LINEAR, F16, TL>(P, L, quarter, tmem_row, act, bias_s, aux_g, quad_bar, sw, s_t[tl], valid_t[tl]);
          else if (L == 9)
            epilogue_q<EPI_COLOUR, F16, TL>(P, L, quarter, tmem_row, act, bias_s, aux_g, quad_bar, sw, s_t[tl], valid_t[tl]);
          else
            epilogue_q<EPI_RELU, F16, TL>(P, L, quarter, tmem_row, act, bias_s, aux_g, quad_bar, sw, s_t[tl], valid_t[tl]);
          if (L == 4) {
            // the xyz features of this tile are dead (all its MMAs retired): store the view features in their place
            if (quarter == 0) {
              const float4* de = reinterpret_cast<const float4*>(P.dir_enc + ray_t[tl] * 24);
              uint32_t f[12];
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                const float4 d4 = __ldg(de + j);
                f[2 * j] = pack16<F16>(d4.x, d4.y);
                f[2 * j + 1] = pack16<F16>(d4.z, d4.w);
              }
#pragma unroll
              for (int j = 0; j < 3; ++j) st_shared_v4(sw.addr(enc, j), f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              st_shared_v4(sw.addr(enc, 3), 0u, 0u, 0u, 0u);
            } else if (quarter == 1) {
#pragma unroll
              for (int j = 4; j < 8; ++j) st_shared_v4(sw.addr(enc, j), 0u, 0u, 0u, 0u);
            }
          }
          if (L != 9) {
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar(B_ACT_READY + tl));
          }
          if (TL && blockIdx.x == 0 && itp < 4 && threadIdx.x == 0)
            reinterpret_cast<long long*>(P.dbg)[(itp * 10 + L) * 16 + 5 + tl * 2] = clock64();
        }
        // both tiles are done with this layer's bias row: refill the buffer with layer L+2's (next pair's for L = 8, 9)
        ++bias_use[bb];
        asm volatile("bar.sync %0, %1;" ::"r"(BAR_ID_EPI_ALL), "r"(EPI_THREADS) : "memory");
        if (loader) {
          const int Ln = (L + 2) % N_MMA_LAYERS;
          const bool more = (L + 2 < N_MMA_LAYERS) || (itp + 1 < iters);
          if (more) {
            fence_proxy_async();
            mbar_expect_tx(bar(B_BIAS_FULL + bb), 1024);
            tma_bulk_g2s(sbase + OFF_BIAS + bb * 1024, aux_bytes_g + Ln * (AUX_REC_FLOATS * 4), 1024, bar(B_BIAS_FULL + bb));
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still multicast into / arrive on this CTA
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem2_dealloc_512(tmem_base);
  }
}

}  // namespace v7

// ---------------------------------------------------------------------------------------------------------
// weight packing: nn.Linear (out,in) fp32 -> bf16 [N x 64] SW128 K-chunks in consumption order
// ---------------------------------------------------------------------------------------------------------
struct PackLayer {
  int w_off, ld, n;  // flat weight offset, row length (in_features), rows
};
struct PackParams {
  PackLayer layer[N_MMA_LAYERS];
  int bias_off[N_MMA_LAYERS];
  int sig_w, sig_b, col_w, col_b;
};

// mode 0: bf16, 1: fp16 (same layout), 2: split fp16 for NT_PREC_TC32 — every chunk is followed by the chunk of its scaled
// residuals, lo = fp16((w - hi) * 2^12), and the fp32 side block sits behind the doubled image
__global__ void pack_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed, PackParams pp, int mode) {
  const int w_bytes = mode == 2 ? 2 * PACKED_W_BYTES : PACKED_W_BYTES;
  // one thread per 16-byte chunk (8 bf16) of the packed image, then one per float of the fp32 side block
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= PACKED_W_BYTES / 16) {
    const int a = gid - PACKED_W_BYTES / 16;
    if (a >= AUX_FLOATS) return;
    float v = 0.f;
    const int L = a / AUX_REC_FLOATS, i = a % AUX_REC_FLOATS;
    if (i < layer_n(L))
      v = params[pp.bias_off[L] + i];
    else if (L == 9 && i >= AUX_COL_B && i < AUX_COL_B + 3)
      v = params[pp.col_b + (i - AUX_COL_B)];
    else if (L == 7 && i >= AUX_EXTRA && i < AUX_EXTRA + 256)
      v = params[pp.sig_w + (i - AUX_EXTRA)];
    else if (L == 7 && i == AUX_SIG_B)
      v = params[pp.sig_b];
    else if (L == 9 && i >= AUX_EXTRA)
      v = params[pp.col_w + (i - AUX_EXTRA)];
    reinterpret_cast<float*>(packed + w_bytes)[a] = v;
    return;
  }
  int byte = gid * 16, L = 0, kc = 0, base = 0;
  for (L = 0; L < N_MMA_LAYERS; ++L) {
    const int lb = layer_chunks(L) * chunk_bytes(L);
    if (byte < lb) break;
    byte -= lb;
    base += lb;
  }
  kc = byte / chunk_bytes(L);
  byte -= kc * chunk_bytes(L);
  const int n = byte / 128;
  const int jphys = (byte % 128) / 16;
  const int j = jphys ^ (n & 7);  // logical 16-byte chunk stored at this swizzled position
  const PackLayer pl = pp.layer[L];
  uint32_t out[4], out_lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = j * 8 + e * 2 + h;  // K index inside the chunk
      int col = -1;                      // source column in the (out,in) matrix
      if (L == 0)
        col = k < 60 ? k : -1;
      else if (L == 4)
        col = kc < 4 ? kc * 64 + k : (k < 60 ? 256 + k : -1);  // [hidden | xyz enc] (nerf.py:109)
      else if (L == 9)
        col = kc < 4 ? 24 + kc * 64 + k : (k < 24 ? k : -1);   // [dir enc | point_info] (nerf.py:118), info first in K
      else
        col = kc * 64 + k;
      v[h] = col >= 0 ? params[pl.w_off + (int64_t)n * pl.ld + col] : 0.f;
    }
    out[e] = mode ? pack_f16(v[0], v[1]) : pack_bf16(v[0], v[1]);
    if (mode == 2) {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&out[e]));
      out_lo[e] = pack_f16((v[0] - h.x) * 4096.f, (v[1] - h.y) * 4096.f);
    }
  }
  if (mode == 2) {
    uint8_t* dst = packed + 2 * (size_t)base + (size_t)kc * 2 * chunk_bytes(L) + byte;
    *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<uint4*>(dst + chunk_bytes(L)) = make_uint4(out_lo[0], out_lo[1], out_lo[2], out_lo[3]);
  } else {
    *reinterpret_cast<uint4*>(packed + (size_t)gid * 16) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

const int kMmaLayerIndex[N_MMA_LAYERS] = {L_P0, L_P1, L_P2, L_P3, L_P4, L_P5, L_P6, L_P7, L_INFO, L_DIR};

}  // namespace

size_t nt_mlp_tc_packed_bytes() { return PACKED_BYTES; }

int nt_mlp_tc_pack(nt_ctx* ctx, const float* params, void* packed, int mode, cudaStream_t st) {
  const LayerTable T = nt_layers();
  PackParams pp;
  for (int i = 0; i < N_MMA_LAYERS; ++i) {
    const int li = kMmaLayerIndex[i];
    pp.layer[i].w_off = (int)T.w[li];
    pp.layer[i].ld = kLayerIn[li];
    pp.layer[i].n = kLayerOut[li];
    pp.bias_off[i] = (int)T.b[li];
  }
  pp.sig_w = (int)T.w[L_SIGMA];
  pp.sig_b = (int)T.b[L_SIGMA];
  pp.col_w = (int)T.w[L_COLOR];
  pp.col_b = (int)T.b[L_COLOR];
  const int threads = PACKED_W_BYTES / 16 + AUX_FLOATS;
  pack_weights_kernel<<<(threads + 255) / 256, 256, 0, st>>>(params, (uint8_t*)packed, pp, mode);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

static float* g_fwd_prof = nullptr;
// diagnostics only (tools/fwd_stash_timeline.py): device buffer of 4 * 10 * 16 int64 that the next stash-forward launches stamp
extern "C" void nt_debug_set_fwd_prof(void* p) { g_fwd_prof = (float*)p; }

static int mlp_tc_launch(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                         const float* params, const void* packed, float* rgb, float* sigma, float* dbg, int dbg_layer,
                         const TcStash* stash, int fp16, cudaStream_t st) {
  if (fp16 && (stash || (dbg && dbg_layer < 100))) {
    nt_set_error("NT_PREC_FP16 operands are a rendering mode: training / the per-layer dump run on the bf16 kernels");
    return NT_ERR_UNSUPPORTED;
  }
  if (!(ctx->attr_done & NT_ATTR_MLP_TC5)) {
    NT_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NT_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NT_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ctx->attr_done |= NT_ATTR_MLP_TC5;
  }
  TcParams P;
  memset(&P, 0, sizeof(P));
  if (stash) {
    for (int i = 0; i < N_MMA_LAYERS; ++i) P.st[i] = (__nv_bfloat16*)stash->layer[i];
    P.st_enc = (__nv_bfloat16*)stash->enc;
    P.st_denc = (__nv_bfloat16*)stash->denc;
    P.st_zsig = stash->zsig;
    const int64_t S = n * p;
    for (int i = 0; i < 9; ++i) {
      int rc = nt_make_map_bf16(&P.st_map[i], stash->layer[i], S, 256, 256, 64, TILE_M);
      if (rc != NT_OK) return rc;
    }
    int rc = nt_make_map_bf16(&P.st_map[9], stash->enc, S, 64, 64, 64, TILE_M);
    if (rc != NT_OK) return rc;
    rc = nt_make_map_bf16(&P.st_map[10], stash->denc, S, 64, 64, 64, TILE_M);
    if (rc != NT_OK) return rc;
    // NT_STASH_MODE: 3 (default) TMA stash stores issued by the MMA thread, one 16 KB box after each K-chunk's MMAs; 1: issued
    // by an elected epilogue thread per tile right after the epilogue (measured 8 % slower: 448 vs 415 us per 1024-ray step,
    // the four boxes arrive at the TMA unit as a burst in front of the weight ring's loads); 0: plain global stores from the
    // epilogue registers; 2: no stash stores at all (timing experiment only: 322 us)
    const char* sm = getenv("NT_STASH_MODE");
    P.use_tma_stash = sm ? atoi(sm) : 3;
    P.st_bits = stash->bits;
  }
  P.t = t;
  P.rays = rays;
  P.dir_enc = dir_enc;
  P.params = params;
  P.packed = (const uint8_t*)packed;
  P.rgb = rgb;
  P.sigma = sigma;
  P.dbg = dbg;
  P.dbg_layer = dbg_layer;
  P.total = n * p;
  P.p = p;
  P.p_shift = -1;
  for (int b = 0; b < 30; ++b)
    if (p == (1 << b)) P.p_shift = b;
  const int64_t tiles = (P.total + TILE_M - 1) / TILE_M;
  P.num_pairs = (int)((tiles + 1) / 2);
  if (P.num_pairs == 0) return NT_OK;
  // lock-step kernel: tile pairs, except that a last wave which would occupy at most half of the CTAs runs as single tiles
  P.full_pairs = P.num_pairs;
  P.num_items = P.num_pairs;
  {
    const int64_t G = ctx->sm_count, waves = tiles / (2 * G), rest = tiles - 2 * G * waves;
    if (rest > 0 && rest <= G && !getenv("NT_NO_SINGLE_TILES")) {
      P.full_pairs = (int)(G * waves);
      P.num_items = (int)(G * waves + rest);
    }
  }
  int grid = ctx->sm_count < P.num_items ? ctx->sm_count : P.num_items;
  // schedule variants (NT_OPT_MLP_TC_VERSION / env NT_MLP_TC_VERSION): 5 = tile pair in lock-step (the training
  // instantiation and the per-layer debug dump always use it), 6 = staggered tiles + 2-CTA weight multicast,
  // 7 = staggered tiles + cta_group::2 MMAs with layer-stationary weights (default for rendering, fastest)
  int ver = ctx->opt_tc_version;
  if (ver == 0) {
    const char* e = getenv("NT_MLP_TC_VERSION");
    ver = e ? atoi(e) : 7;
  }
  const bool use_v7 = ver == 7;
  if ((use_v7 || fp16) && (!dbg || dbg_layer >= 100) && !stash) {
    if (!(ctx->attr_done & NT_ATTR_MLP_TC7)) {
      NT_CUDA(cudaFuncSetAttribute(v7::mlp_tc7_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::SMEM7_BYTES));
      NT_CUDA(cudaFuncSetAttribute(v7::mlp_tc7_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::SMEM7_BYTES));
      NT_CUDA(cudaFuncSetAttribute(v7::mlp_tc7_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::SMEM7_BYTES));
      NT_CUDA(cudaFuncSetAttribute(v7::mlp_tc7_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::SMEM7_BYTES));
      ctx->attr_done |= NT_ATTR_MLP_TC7;
    }
    int g7 = ctx->sm_count & ~1;                       // whole 2-CTA clusters
    const int need = ((P.num_pairs + 1) / 2) * 2;
    if (g7 > need) g7 = need;
    if (dbg && fp16)
      v7::mlp_tc7_kernel<true, true><<<g7, N_THREADS, v7::SMEM7_BYTES, st>>>(P);
    else if (dbg)
      v7::mlp_tc7_kernel<false, true><<<g7, N_THREADS, v7::SMEM7_BYTES, st>>>(P);
    else if (fp16)
      v7::mlp_tc7_kernel<true, false><<<g7, N_THREADS, v7::SMEM7_BYTES, st>>>(P);
    else
      v7::mlp_tc7_kernel<false, false><<<g7, N_THREADS, v7::SMEM7_BYTES, st>>>(P);
    NT_LAUNCH_CHECK(ctx);
    return NT_OK;
  }
  if (stash && g_fwd_prof) {
    NT_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    P.dbg = g_fwd_prof;
    P.dbg_layer = 100;
    mlp_tc_kernel<true, true><<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  } else if (stash)
    mlp_tc_kernel<false, true><<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  else if (dbg)
    mlp_tc_kernel<true, false><<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  else
    mlp_tc_kernel<false, false><<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

int nt_mlp_tc_forward_dbg(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                          const float* params, const void* packed, float* rgb, float* sigma, float* dbg, int dbg_layer,
                          cudaStream_t st) {
  return mlp_tc_launch(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, dbg, dbg_layer, nullptr, 0, st);
}
int nt_mlp_tc_forward(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                      const float* params, const void* packed, float* rgb, float* sigma, int fp16, cudaStream_t st) {
  return mlp_tc_launch(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, nullptr, -1, nullptr, fp16, st);
}
int nt_mlp_tc_forward_stash(nt_ctx* ctx, int64_t n, int p, const float* t, const float* rays, const float* dir_enc,
                            const float* params, const void* packed, float* rgb, float* sigma, const TcStash* stash,
                            cudaStream_t st) {
  return mlp_tc_launch(ctx, n, p, t, rays, dir_enc, params, packed, rgb, sigma, nullptr, -1, stash, 0, st);
}
