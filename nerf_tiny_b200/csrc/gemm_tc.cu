// Generic bf16 tcgen05 GEMM with fp32 accumulation in TMEM — the backbone of the NT_PREC_BF16 training path
// (autograd of Network.forward, nerf.py:101-124 / loss.backward() nerf.py:473).
//
//   D[M, N] = A (.) B          M = rows of the CTA tile (128), N = BN (64 | 128 | 256), K streamed in 64-wide blocks
//   K_MAJOR : A stored [M][K], B stored [N][K]           (dX = G . W : B is the pre-transposed weight)
//   MN_MAJOR: A stored [K][M], B stored [K][N]           (dW = G^T . H : K is the SAMPLE axis, split over CTAs)
// Operands are fetched with TMA tensor maps (cp.async.bulk.tensor, SWIZZLE_128B) into a 4-stage mbarrier ring by
// one producer warp; one elected thread issues tcgen05.mma (M=128, N=BN, K=16); two TMEM accumulator stages let the
// 8 epilogue warps drain tile i while tile i+1 is being multiplied.  With b_resident the (small) B operand is loaded
// once per CTA and kept in shared memory while M-tiles of A stream past it.
// Epilogues: bf16 store with optional bias / ReLU / ReLU'-mask (mask = stored activation > 0), or fp32 vector
// atomics (split-K weight gradients).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NSTAGE = 4;
__host__ __device__ constexpr int gemm_stages(int bn) { return bn == 64 ? 8 : NSTAGE; }
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int GEMM_THREADS = 320;  // warp 0 = TMA, warp 1 = MMA, warps 2-9 = epilogue

struct GemmTcArgs {
  int M, N, K;        // problem (N <= BN, one N tile)
  int n_valid;        // columns actually written (<= N)
  int split_k;        // CTAs along K (atomic epilogue only)
  int kb_per_split;   // 64-wide K blocks per split
  int b_resident;     // keep all K blocks of B in smem (requires K <= NSTAGE*64, split_k == 1)
  int epi;            // 0 = bf16 store, 1 = fp32 atomic add, 2 = fp32 store
  void* C;
  int ldc;
  const float* bias;  // [N] or null (bf16 store only)
  int relu;
  const __nv_bfloat16* mask;  // [M][ldmask] or null: out = mask > 0 ? acc : 0
  int ldmask;
  const float* r1_row;  // rank-1 term v += r1_row[m] * r1_col[n] (sigma-head gradient joining point_info's)
  const float* r1_col;
  float* colsum;        // optional [N]: += column sums of the stored matrix (bias gradient of the layer below)
  const float* enc_t;   // epi == 3: fused encoder backward (GemmTcEpi)
  const float* enc_rays;
  int enc_p;
  float* enc_gt;
};

// epi == 3: this thread holds g_enc[m][n0 .. n0+31] (n0 = 0 or 32) of sample m; features f = c*20 + 2l + {sin, cos}
__device__ __forceinline__ void enc_backward_part(const GemmTcArgs& g, int m, int n0, const float (&v)[32]) {
  constexpr uint32_t kFreq[10] = NT_FREQ_POINT_INIT;
  const float* ray = g.enc_rays + (int64_t)(m / g.enc_p) * 16;
  float r[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(ray) + q);
    r[4 * q] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
  }
  const float t = __ldg(g.enc_t + m);
  // p_cam = fl(d_cam * t); p_wrd = ((R0 p0 + R1 p1) + R2 p2) + T, no FMA (nerf.py:200-216); d_wrd likewise (nerf.py:211)
  const float pc0 = __fmul_rn(r[0], t), pc1 = __fmul_rn(r[1], t), pc2 = __fmul_rn(r[2], t);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // (c, l) pairs whose two features fall into this half: half 0 -> c = 0 (all l), c = 1 (l < 6); half 1 -> c = 1 (l >= 6), c = 2
    const int l_lo = n0 == 0 ? 0 : (c == 1 ? 6 : 0), l_hi = n0 == 0 ? (c == 0 ? 10 : (c == 1 ? 6 : 0)) : (c == 0 ? 0 : 10);
    if (l_lo >= l_hi) continue;
    const float pos = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3 + c * 3], pc0), __fmul_rn(r[4 + c * 3], pc1)),
                                          __fmul_rn(r[5 + c * 3], pc2)), r[12 + c]);
    const float dw = __fadd_rn(__fadd_rn(__fmul_rn(r[3 + c * 3], r[0]), __fmul_rn(r[4 + c * 3], r[1])), __fmul_rn(r[5 + c * 3], r[2]));
    float gp = 0.f;
#pragma unroll
    for (int l = 0; l < 10; ++l) {
      if (l < l_lo || l >= l_hi) continue;
      const float w = __uint_as_float(kFreq[l]);
      const float x = __fmul_rn(w, pos);
      const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.f), -12582912.f);
      float rr = fmaf(k, -6.2831854820251465f, x);
      rr = fmaf(k, 1.7484555314695172e-07f, rr);
      const float sn = __sinf(rr), cs = __cosf(rr);
      const int f = c * 20 + 2 * l - n0;  // compile-time after unrolling (n0 selects one of two instantiations below)
      gp += w * (v[f] * cs - v[f + 1] * sn);
    }
    acc += gp * dw;
  }
  atomicAdd(g.enc_gt + m, acc);
}

// PTX wrappers: tc_ptx.cuh
using namespace tcptx;
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool mn_major) { return idesc_f16kind(BM, n, true, mn_major); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  umma_f16(d_tmem, adesc, bdesc, idesc, acc);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { return pack_bf16(lo, hi); }

// fp32 atomic accumulation of 32 consecutive values: 16-byte vector reductions on the aligned middle, scalar atomics on
// the (at most 3 + 3) unaligned head/tail elements — the flat parameter layout is only 4-byte aligned after the
// 1-wide sigma bias
__device__ __forceinline__ void red_add_32(float* c, const uint32_t (&raw)[32], int n_ok) {
  const int head = (4 - (int)((reinterpret_cast<uintptr_t>(c) >> 2) & 3)) & 3;
  if (head == 0 && n_ok >= 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + 4 * j), "f"(__uint_as_float(raw[4 * j])),
                   "f"(__uint_as_float(raw[4 * j + 1])), "f"(__uint_as_float(raw[4 * j + 2])),
                   "f"(__uint_as_float(raw[4 * j + 3]))
                   : "memory");
    return;
  }
  // head is warp-uniform in practice (same row pitch for every lane), so these branches do not diverge
#pragma unroll
  for (int hh = 1; hh <= 3; ++hh) {
    if (head == hh) {
#pragma unroll
      for (int j = 0; j < hh; ++j)
        if (j < n_ok) atomicAdd(c + j, __uint_as_float(raw[j]));
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (hh + 4 * j + 4 <= n_ok)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + hh + 4 * j),
                       "f"(__uint_as_float(raw[hh + 4 * j])), "f"(__uint_as_float(raw[hh + 4 * j + 1])),
                       "f"(__uint_as_float(raw[hh + 4 * j + 2])), "f"(__uint_as_float(raw[hh + 4 * j + 3]))
                       : "memory");
        else {
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2)
            if (hh + 4 * j + k2 < n_ok) atomicAdd(c + hh + 4 * j + k2, __uint_as_float(raw[hh + 4 * j + k2]));
        }
      }
#pragma unroll
      for (int j = hh + 28; j < 32; ++j)
        if (j < n_ok) atomicAdd(c + j, __uint_as_float(raw[j]));
    }
  }
  if (head == 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_ok) atomicAdd(c + j, __uint_as_float(raw[j]));
  }
}

template <int BN, bool MN_MAJOR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmTcArgs g) {
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int NST = gemm_stages(BN);  // ring depth: 8 stages for the narrow (BN = 64) shape, 4 otherwise
  constexpr int OFF_A = 0;
  constexpr int OFF_B = NST * A_STAGE_BYTES;
  constexpr int OFF_BAR = OFF_B + NST * B_STAGE_BYTES;
  constexpr int B_EMPTY = NST, B_ACCF = 2 * NST, B_ACCE = 2 * NST + 2, B_BFULL = 2 * NST + 4;  // barrier slots
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // barriers: full[NST], empty[NST], acc_full[2], acc_empty[2], b_full
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * (B_BFULL + 1));

  const int m_tiles = (g.M + BM - 1) / BM;
  const int kb_total = g.K / BK;
  const int items = m_tiles * g.split_k;

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar(s), 1);
      mbar_init(bar(B_EMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_ACCF + s), 1);
      mbar_init(bar(B_ACCE + s), 256);
    }
    mbar_init(bar(B_BFULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t q = 0;  // ring counter
      bool b_loaded = false;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int mt = item % m_tiles, sp = item / m_tiles;
        const int kb0 = sp * g.kb_per_split;
        const int kb1 = min(kb_total, kb0 + g.kb_per_split);
        if (g.b_resident && !b_loaded) {
          mbar_expect_tx(bar(B_BFULL), kb_total * B_STAGE_BYTES);
          for (int kb = 0; kb < kb_total; ++kb) {
            const uint32_t dst = sbase + OFF_B + kb * B_STAGE_BYTES;
            if (MN_MAJOR) {
              for (int b = 0; b < BN / 64; ++b) tma_load_2d(dst + b * 8192, &map_b, b * 64, kb * BK, bar(B_BFULL));
            } else {
              tma_load_2d(dst, &map_b, kb * BK, 0, bar(B_BFULL));
            }
          }
          b_loaded = true;
        }
        for (int kb = kb0; kb < kb1; ++kb, ++q) {
          const uint32_t s = q % NST;
          mbar_wait(bar(B_EMPTY + s), ((q / NST) & 1) ^ 1);
          mbar_expect_tx(bar(s), A_STAGE_BYTES + (g.b_resident ? 0 : B_STAGE_BYTES));
          const uint32_t da = sbase + OFF_A + s * A_STAGE_BYTES;
          if (MN_MAJOR) {
            tma_load_2d(da, &map_a, mt * BM, kb * BK, bar(s));
            tma_load_2d(da + 8192, &map_a, mt * BM + 64, kb * BK, bar(s));
          } else {
            tma_load_2d(da, &map_a, kb * BK, mt * BM, bar(s));
          }
          if (!g.b_resident) {
            const uint32_t db = sbase + OFF_B + s * B_STAGE_BYTES;
            if (MN_MAJOR) {
              for (int b = 0; b < BN / 64; ++b) tma_load_2d(db + b * 8192, &map_b, b * 64, kb * BK, bar(s));
            } else {
              tma_load_2d(db, &map_b, kb * BK, 0, bar(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BN, MN_MAJOR);
      uint32_t q = 0, it = 0;
      bool b_ready = false;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int sp = item / m_tiles;
        const int kb0 = sp * g.kb_per_split;
        const int kb1 = min(kb_total, kb0 + g.kb_per_split);
        const uint32_t as = it & 1;  // accumulator stage
        mbar_wait(bar(B_ACCE + as), ((it >> 1) & 1) ^ 1);  // epilogue drained this TMEM stage
        tc_fence_after();
        if (g.b_resident && !b_ready) {
          mbar_wait(bar(B_BFULL), 0);
          b_ready = true;
        }
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = kb0; kb < kb1; ++kb, ++q) {
          const uint32_t s = q % NST;
          mbar_wait(bar(s), (q / NST) & 1);
          tc_fence_after();
          const uint32_t a_addr = sbase + OFF_A + s * A_STAGE_BYTES;
          const uint32_t b_addr = sbase + OFF_B + (g.b_resident ? kb : s) * B_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint64_t ad, bd;
            if (MN_MAJOR) {  // 16 K-rows = 2 KB per step inside each 64-element MN block (blocks 8 KB apart)
              ad = desc_mnmajor(a_addr + j * 2048, 8192);
              bd = desc_mnmajor(b_addr + j * 2048, 8192);
            } else {
              ad = desc_kmajor(a_addr + j * 32);
              bd = desc_kmajor(b_addr + j * 32);
            }
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
          }
          umma_commit(bar(B_EMPTY + s));
        }
        umma_commit(bar(B_ACCF + as));
      }
    }
  } else {
    // ======================= epilogue: 8 warps = 4 lane quadrants x 2 column halves =======================
    const int e = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = e >> 2;
    const int row_in_tile = quad * 32 + lane;
    constexpr int NCB = BN / 64;  // 32-column blocks per half
    uint32_t it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int mt = item % m_tiles;
      const uint32_t as = it & 1;
      const int m = mt * BM + row_in_tile;
      // ReLU' masks of this thread's row come from HBM: fetch them while the MMAs of this tile are still running
      uint4 mpre[NCB][4];
      const bool mask_fast = g.mask != nullptr && g.epi == 0 && m < g.M && (half + 1) * NCB * 32 <= g.n_valid &&
                             ((reinterpret_cast<uintptr_t>(g.mask + (int64_t)m * g.ldmask) & 15) == 0);
      if (mask_fast) {
        const uint4* mk = reinterpret_cast<const uint4*>(g.mask + (int64_t)m * g.ldmask + half * NCB * 32);
#pragma unroll
        for (int i = 0; i < NCB; ++i)
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) mpre[i][qd] = __ldg(mk + i * 4 + qd);
      }
      mbar_wait(bar(B_ACCF + as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 256;
#pragma unroll
      for (int i = 0; i < NCB; ++i) {
        const int n0 = (half * NCB + i) * 32;
        uint32_t raw[32];
        tmem_ld32(trow + n0, raw);
        float cs[32];
        if (g.colsum) {
#pragma unroll
          for (int j = 0; j < 32; ++j) cs[j] = 0.f;
        }
        if (m < g.M && n0 < g.n_valid) {
          if (g.epi == 1) {
            float* c = reinterpret_cast<float*>(g.C) + (int64_t)m * g.ldc + n0;
            red_add_32(c, raw, g.n_valid - n0);
          } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
            if (g.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(g.bias + n0 + j);
            }
            if (g.r1_row) {
              const float rr = __ldg(g.r1_row + m);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaf(rr, __ldg(g.r1_col + n0 + j), v[j]);
            }
            if (g.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (g.mask && ((reinterpret_cast<uintptr_t>(g.mask + (int64_t)m * g.ldmask + n0) & 15) != 0 || n0 + 32 > g.n_valid)) {
              const __nv_bfloat16* mk = g.mask + (int64_t)m * g.ldmask + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.n_valid && !(__bfloat162float(mk[j]) > 0.f)) v[j] = 0.f;
            } else if (g.mask) {
              const uint4* mk = reinterpret_cast<const uint4*>(g.mask + (int64_t)m * g.ldmask + n0);
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                const uint4 mm = mask_fast ? mpre[i][qd] : __ldg(mk + qd);
                const uint32_t w[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  // bf16 > 0  <=>  sign clear and not zero
                  const uint32_t lo = w[k2] & 0xffffu, hi = w[k2] >> 16;
                  if (!(lo != 0 && lo < 0x8000u)) v[qd * 8 + k2 * 2] = 0.f;
                  if (!(hi != 0 && hi < 0x8000u)) v[qd * 8 + k2 * 2 + 1] = 0.f;
                }
              }
            }
            if (g.epi == 3) {
              if (n0 == 0)
                enc_backward_part(g, m, 0, v);
              else if (n0 == 32)
                enc_backward_part(g, m, 32, v);
              continue;
            }
            if (g.epi == 2) {
              float* cf = reinterpret_cast<float*>(g.C) + (int64_t)m * g.ldc + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.n_valid) cf[j] = v[j];
              continue;
            }
            if (g.colsum) {
#pragma unroll
              for (int j = 0; j < 32; ++j) cs[j] = v[j];
            }
            __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(g.C) + (int64_t)m * g.ldc + n0;
            if (n0 + 32 <= g.n_valid && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
#pragma unroll
              for (int qd = 0; qd < 4; ++qd)
                *reinterpret_cast<uint4*>(c + qd * 8) =
                    make_uint4(pack2(v[qd * 8], v[qd * 8 + 1]), pack2(v[qd * 8 + 2], v[qd * 8 + 3]),
                               pack2(v[qd * 8 + 4], v[qd * 8 + 5]), pack2(v[qd * 8 + 6], v[qd * 8 + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.n_valid) c[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
        if (g.colsum && n0 < g.n_valid) {
          // column sums over this warp's 32 rows by recursive halving: after 5 exchange steps lane L holds column L
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const float send = up ? cs[j] : cs[j + o];
              const float keep = up ? cs[j + o] : cs[j];
              cs[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          if (n0 + lane < g.n_valid) atomicAdd(g.colsum + n0 + lane, cs[0]);
        }
      }
      tc_fence_before();
      mbar_arrive(bar(B_ACCE + as));
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMM  dW[256 x N] += G^T . H  over the sample axis, both 128-row halves of the output
// in ONE CTA so that G and H each cross HBM exactly once:  A = G stored [S][256] (MN-major, M = 256 = 4 boxes of 64),
// B = H stored [S][N<=256].  Split-K over the samples, 3-stage TMA ring of 64-sample blocks, two TMEM accumulators
// (columns 0-255 / 256-511), fp32 vector atomics into the flat gradient buffer.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_STAGES = 3;
constexpr int DW_A_BYTES = 4 * 8192;  // 256 output rows = 4 x (64 elements x 64 samples)

struct DwArgs {
  int S;           // samples (K), multiple of 64
  int m_valid;     // output rows actually present (<= 256)
  int n_valid;     // output columns written
  int kb_per_cta;  // 64-sample blocks per CTA
  float* C;
  int ldc;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    dw_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const DwArgs g) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = DW_A_BYTES + B_BYTES;
  constexpr int OFF_BAR = DW_STAGES * STAGE;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };  // full[3] 0-2, empty[3] 3-5, acc_full 6
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
  const int kb_total = (g.S + BK - 1) / BK;  // rows past S are zero-filled by TMA: a partial last block adds nothing
  const int kb0 = blockIdx.x * g.kb_per_cta;
  const int kb1 = min(kb_total, kb0 + g.kb_per_cta);

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(bar(s), 1);
      mbar_init(bar(3 + s), 1);
    }
    mbar_init(bar(6), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t q = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++q) {
        const uint32_t s = q % DW_STAGES;
        mbar_wait(bar(3 + s), ((q / DW_STAGES) & 1) ^ 1);
        mbar_expect_tx(bar(s), STAGE);
        const uint32_t da = sbase + s * STAGE, db = da + DW_A_BYTES;
#pragma unroll
        for (int b = 0; b < 4; ++b) tma_load_2d(da + b * 8192, &map_a, b * 64, kb * BK, bar(s));
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) tma_load_2d(db + b * 8192, &map_b, b * 64, kb * BK, bar(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BN, true);
      uint32_t q = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++q) {
        const uint32_t s = q % DW_STAGES;
        mbar_wait(bar(s), (q / DW_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = sbase + s * STAGE, b_addr = a_addr + DW_A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t bd = desc_mnmajor(b_addr + j * 2048, 8192);
          const uint32_t acc = (kb > kb0 || j > 0) ? 1u : 0u;
          umma_bf16(tmem_base, desc_mnmajor(a_addr + j * 2048, 8192), bd, idesc, acc);                 // rows 0-127
          umma_bf16(tmem_base + 256, desc_mnmajor(a_addr + 16384 + j * 2048, 8192), bd, idesc, acc);   // rows 128-255
        }
        umma_commit(bar(3 + s));
      }
      umma_commit(bar(6));
    }
  } else if (kb1 > kb0) {
    const int e = warp - 2, quad = warp & 3, half = e >> 2;
    constexpr int NCB = BN / 64;
    mbar_wait(bar(6), 0);
    tc_fence_after();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int m = mt * 128 + quad * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + mt * 256;
#pragma unroll
      for (int i = 0; i < NCB; ++i) {
        const int n0 = (half * NCB + i) * 32;
        uint32_t raw[32];
        tmem_ld32(trow + n0, raw);
        if (m < g.m_valid && n0 < g.n_valid) {
          float* c = g.C + (int64_t)m * g.ldc + n0;
          red_add_32(c, raw, g.n_valid - n0);
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Grouped weight-gradient GEMM: every dW of one backward pass (14 problems: 10 tensor-core layers, the two
// concatenation tails and the two heads) in ONE launch.  Each problem gets a slice of the grid proportional to its
// HBM traffic and is split over the sample axis inside its slice, so a [256 x 256] gradient receives ~14 atomic
// partial sums instead of 148 and the launch runs at HBM speed instead of atomic speed.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DWG_MAX = 16;
struct DwGroup {
  CUtensorMap map_a[DWG_MAX];
  CUtensorMap map_b[DWG_MAX];
  float* C[DWG_MAX];
  int ldc[DWG_MAX], m_valid[DWG_MAX], n_valid[DWG_MAX];
  int cta_begin[DWG_MAX + 1];  // grid slice of every problem
  int kb_per_cta[DWG_MAX];
  int n_problems;
  int S;
};

// Persistent: a CTA walks work items blockIdx.x, blockIdx.x + gridDim.x, ... (item = one problem's K-split); the TMA
// producer runs ahead into the next item's sample blocks while the epilogue warps drain the accumulator, and the grid can
// be capped so that the launch shares the GPU with an MMA-bound kernel on another stream (train_bf16.cu).
__global__ void __launch_bounds__(GEMM_THREADS, 1) dw_grouped_kernel(const __grid_constant__ DwGroup g) {
  constexpr int BN = 256;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = DW_A_BYTES + B_BYTES;
  constexpr int OFF_BAR = DW_STAGES * STAGE;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };  // full[3] 0-2, empty[3] 3-5, acc_full 6, acc_empty 7
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
  const int n_items = g.cta_begin[g.n_problems];
  const int kb_total = (g.S + BK - 1) / BK;  // rows past S are zero-filled by TMA: a partial last block adds nothing
  constexpr int N_EPI_WARPS = GEMM_THREADS / 32 - 2;

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(bar(s), 1);
      mbar_init(bar(3 + s), 1);
    }
    mbar_init(bar(6), 1);
    mbar_init(bar(7), N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (problem, K range); evaluated identically by every role
  auto locate = [&](int item, int& pr, int& kb0, int& kb1) {
    pr = 0;
    while (pr + 1 < g.n_problems && item >= g.cta_begin[pr + 1]) ++pr;
    kb0 = (item - g.cta_begin[pr]) * g.kb_per_cta[pr];
    kb1 = min(kb_total, kb0 + g.kb_per_cta[pr]);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t q = 0;
      const uint64_t pol_ld = l2_policy_evict_first();
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int pr, kb0, kb1;
        locate(item, pr, kb0, kb1);
        const int a_boxes = (g.m_valid[pr] + 63) / 64, b_boxes = (g.n_valid[pr] + 63) / 64;  // boxes not entirely out of range
        for (int kb = kb0; kb < kb1; ++kb, ++q) {
          const uint32_t s = q % DW_STAGES;
          mbar_wait(bar(3 + s), ((q / DW_STAGES) & 1) ^ 1);
          mbar_expect_tx(bar(s), (a_boxes + b_boxes) * 8192);
          const uint32_t da = sbase + s * STAGE, db = da + DW_A_BYTES;
          // the stash is read exactly once: evict_first keeps it from displacing the fp32 gradient the epilogues reduce into
#ifdef NT_DW_NO_HINT
          for (int b = 0; b < a_boxes; ++b) tma_load_2d(da + b * 8192, &g.map_a[pr], b * 64, kb * BK, bar(s));
          for (int b = 0; b < b_boxes; ++b) tma_load_2d(db + b * 8192, &g.map_b[pr], b * 64, kb * BK, bar(s));
#else
          for (int b = 0; b < a_boxes; ++b) tma_load_2d_hint(da + b * 8192, &g.map_a[pr], b * 64, kb * BK, bar(s), pol_ld);
          for (int b = 0; b < b_boxes; ++b) tma_load_2d_hint(db + b * 8192, &g.map_b[pr], b * 64, kb * BK, bar(s), pol_ld);
#endif
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t q = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int pr, kb0, kb1;
        locate(item, pr, kb0, kb1);
        const int b_boxes = (g.n_valid[pr] + 63) / 64;
        const bool two_m = g.m_valid[pr] > 128;
        // N of the instruction = the columns that exist (multiple of 64): columns beyond were never loaded
        const uint32_t idesc = idesc_bf16(b_boxes * 64, true);
        mbar_wait(bar(7), (it & 1) ^ 1);  // the previous item's accumulator has been drained
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++q) {
          const uint32_t s = q % DW_STAGES;
          mbar_wait(bar(s), (q / DW_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = sbase + s * STAGE, b_addr = a_addr + DW_A_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t bd = desc_mnmajor(b_addr + j * 2048, 8192);
            const uint32_t acc = (kb > kb0 || j > 0) ? 1u : 0u;
            umma_bf16(tmem_base, desc_mnmajor(a_addr + j * 2048, 8192), bd, idesc, acc);
            if (two_m) umma_bf16(tmem_base + 256, desc_mnmajor(a_addr + 16384 + j * 2048, 8192), bd, idesc, acc);
          }
          umma_commit(bar(3 + s));
        }
        umma_commit(bar(6));
      }
    }
  } else {
    const int e = warp - 2, quad = warp & 3, half = e >> 2;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int pr, kb0, kb1;
      locate(item, pr, kb0, kb1);
      const int m_valid = g.m_valid[pr], n_valid = g.n_valid[pr];
      const int b_boxes = (n_valid + 63) / 64;
      const bool two_m = m_valid > 128;
      mbar_wait(bar(6), it & 1);
      tc_fence_after();
      float* Cp = g.C[pr];
      const int ldc = g.ldc[pr];
      for (int mt = 0; mt < (two_m ? 2 : 1); ++mt) {
        const int m = mt * 128 + quad * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + mt * 256;
        for (int i = 0; i < 4; ++i) {
          const int n0 = (half * 4 + i) * 32;
          if (n0 >= b_boxes * 64) continue;  // warp-uniform: these columns were not computed
          uint32_t raw[32];
          tmem_ld32(trow + n0, raw);
          if (m < m_valid && n0 < n_valid) red_add_32(Cp + (int64_t)m * ldc + n0, raw, n_valid - n0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(7));
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host side: tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [rows][cols] (cols contiguous, row pitch ld elements); box = box_cols x box_rows, SWIZZLE_128B
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    nt_set_error("cuTensorMapEncodeTiled entry point not available");
    return NT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nt_set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d base=%p", (int)r, (long long)rows,
                 (long long)cols, (long long)ld, box_cols, box_rows, base);
    return NT_ERR_CUDA;
  }
  return NT_OK;
}

template <int BN, bool MN>
int launch(nt_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const GemmTcArgs& g, cudaStream_t st) {
  constexpr int smem = gemm_stages(BN) * (A_STAGE_BYTES + BN * 128) + 256;
  static bool set[NT_MAX_DEVICES] = {};  // the opt-in is per device (and per template instantiation)
  if (!set[ctx->device]) {
    NT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set[ctx->device] = true;
  }
  const int items = ((g.M + BM - 1) / BM) * g.split_k;
  const int grid = items < ctx->sm_count ? items : ctx->sm_count;
  gemm_tc_kernel<BN, MN><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, g);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

}  // namespace

// C = A (.) B with bf16 operands.  mn_major = 0: A [M][K] (lda), B [N][K] (ldb); 1: A [K][M] (lda), B [K][N] (ldb).
// N is rounded up to 64/128/256 internally (TMA zero-fills out-of-range rows/cols); n_valid columns are written.
int nt_launch_gemm_tc(nt_ctx* ctx, int mn_major, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                      const GemmTcEpi& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return NT_OK;
  if (K % BK != 0 || N > 256) {
    nt_set_error("gemm_tc: K must be a multiple of 64 and N <= 256 (K=%d N=%d)", K, N);
    return NT_ERR_INVALID;
  }
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  GemmTcArgs g;
  g.M = M;
  g.N = BN;
  g.K = K;
  g.n_valid = N;
  g.epi = epi.enc_gt ? 3 : (epi.atomic_f32 ? 1 : (epi.store_f32 ? 2 : 0));
  if (epi.enc_gt && (N != 64 || mn_major || epi.atomic_f32)) {
    nt_set_error("gemm_tc: the fused encoder backward needs N = 64, K-major operands and no split-K");
    return NT_ERR_INVALID;
  }
  g.enc_t = epi.enc_t;
  g.enc_rays = epi.enc_rays;
  g.enc_p = epi.enc_p;
  g.enc_gt = epi.enc_gt;
  g.r1_row = epi.r1_row;
  g.r1_col = epi.r1_col;
  g.colsum = epi.colsum;
  g.C = epi.C;
  g.ldc = epi.ldc;
  g.bias = epi.bias;
  g.relu = epi.relu;
  g.mask = reinterpret_cast<const __nv_bfloat16*>(epi.mask);
  g.ldmask = epi.ldmask;
  const int kb_total = K / BK;
  const int m_tiles = (M + BM - 1) / BM;
  g.split_k = 1;
  if (epi.atomic_f32) {
    int want = (2 * ctx->sm_count) / (m_tiles > 0 ? m_tiles : 1);
    if (want < 1) want = 1;
    if (want > kb_total) want = kb_total;
    g.split_k = want;
  }
  g.kb_per_split = (kb_total + g.split_k - 1) / g.split_k;
  g.split_k = (kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.b_resident = (!epi.atomic_f32 && kb_total <= gemm_stages(BN) && m_tiles > ctx->sm_count) ? 1 : 0;
  CUtensorMap ma, mb;
  int rc;
  if (mn_major) {
    if ((rc = make_map(&ma, A, K, M, lda, 64, BK)) != NT_OK) return rc;
    if ((rc = make_map(&mb, B, K, N, ldb, 64, BK)) != NT_OK) return rc;
  } else {
    if ((rc = make_map(&ma, A, M, K, lda, BK, BM)) != NT_OK) return rc;
    if ((rc = make_map(&mb, B, N, K, ldb, BK, BN)) != NT_OK) return rc;
  }
  if (mn_major) {
    if (BN == 64) return launch<64, true>(ctx, ma, mb, g, st);
    if (BN == 128) return launch<128, true>(ctx, ma, mb, g, st);
    return launch<256, true>(ctx, ma, mb, g, st);
  }
  if (BN == 64) return launch<64, false>(ctx, ma, mb, g, st);
  if (BN == 128) return launch<128, false>(ctx, ma, mb, g, st);
  return launch<256, false>(ctx, ma, mb, g, st);
}

// dW[m_valid x n_valid] += G^T . H : G bf16 [S][ldg] (m_valid <= 256 columns used), H bf16 [S][ldh] (n_valid <= 256)
int nt_launch_dw_gemm(nt_ctx* ctx, int S, const void* G, int ldg, int m_valid, const void* H, int ldh, int n_valid,
                      float* C, int ldc, cudaStream_t st) {
  if (S <= 0 || m_valid <= 0 || n_valid <= 0) return NT_OK;
  if (S % BK != 0 || m_valid > 256 || n_valid > 256) {
    nt_set_error("dw_gemm: S must be a multiple of 64, M,N <= 256");
    return NT_ERR_INVALID;
  }
  const int BN = n_valid <= 64 ? 64 : (n_valid <= 128 ? 128 : 256);
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, G, S, m_valid, ldg, 64, BK)) != NT_OK) return rc;
  if ((rc = make_map(&mb, H, S, n_valid, ldh, 64, BK)) != NT_OK) return rc;
  DwArgs g;
  g.S = S;
  g.m_valid = m_valid;
  g.n_valid = n_valid;
  g.C = C;
  g.ldc = ldc;
  const int kb_total = S / BK;
  int ctas = ctx->sm_count < kb_total ? ctx->sm_count : kb_total;
  g.kb_per_cta = (kb_total + ctas - 1) / ctas;
  ctas = (kb_total + g.kb_per_cta - 1) / g.kb_per_cta;
#define DW_LAUNCH(BN_)                                                                                              \
  {                                                                                                                 \
    constexpr int smem = DW_STAGES * (DW_A_BYTES + BN_ * 128) + 128;                                                \
    static bool set[NT_MAX_DEVICES] = {};                                                                           \
    if (!set[ctx->device]) {                                                                                        \
      NT_CUDA(cudaFuncSetAttribute(dw_gemm_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
      set[ctx->device] = true;                                                                                      \
    }                                                                                                               \
    dw_gemm_kernel<BN_><<<ctas, GEMM_THREADS, smem, st>>>(ma, mb, g);                                               \
  }
  if (BN == 64) DW_LAUNCH(64) else if (BN == 128) DW_LAUNCH(128) else DW_LAUNCH(256)
#undef DW_LAUNCH
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

// ---- grouped dW: problems are queued during the backward pass and flushed in one launch ---------------------------
struct DwQueue {
  DwGroup g;
  int weight[DWG_MAX];
};
static thread_local DwQueue g_dwq;

int nt_dw_group_begin(int S) {
  if (S <= 0) {
    nt_set_error("dw_group: S must be positive");
    return NT_ERR_INVALID;
  }
  g_dwq.g.n_problems = 0;
  g_dwq.g.S = S;
  return NT_OK;
}

int nt_dw_group_add(const void* G, int ldg, int m_valid, const void* H, int ldh, int n_valid, float* C, int ldc) {
  DwGroup& g = g_dwq.g;
  if (g.n_problems >= DWG_MAX || m_valid > 256 || n_valid > 256 || m_valid <= 0 || n_valid <= 0) {
    nt_set_error("dw_group: too many problems or bad shape");
    return NT_ERR_INVALID;
  }
  const int i = g.n_problems;
  int rc;
  if ((rc = make_map(&g.map_a[i], G, g.S, m_valid, ldg, 64, BK)) != NT_OK) return rc;
  if ((rc = make_map(&g.map_b[i], H, g.S, n_valid, ldh, 64, BK)) != NT_OK) return rc;
  g.C[i] = C;
  g.ldc[i] = ldc;
  g.m_valid[i] = m_valid;
  g.n_valid[i] = n_valid;
  g_dwq.weight[i] = ((m_valid + 63) / 64 + (n_valid + 63) / 64);  // 64-wide boxes moved per sample block
  g.n_problems = i + 1;
  return NT_OK;
}

int nt_dw_group_flush(nt_ctx* ctx, cudaStream_t st, int max_ctas) {
  DwGroup& g = g_dwq.g;
  if (g.n_problems == 0) return NT_OK;
  const int kb_total = (g.S + BK - 1) / BK;  // rows past S are zero-filled by TMA: a partial last block adds nothing
  int wsum = 0;
  for (int i = 0; i < g.n_problems; ++i) wsum += g_dwq.weight[i];
  const int budget = 2 * ctx->sm_count;  // two CTAs' worth of work per SM keeps the tail short
  int begin = 0;
  for (int i = 0; i < g.n_problems; ++i) {
    int ctas = (int)((long long)budget * g_dwq.weight[i] / wsum);
    if (ctas < 1) ctas = 1;
    if (ctas > kb_total) ctas = kb_total;
    g.kb_per_cta[i] = (kb_total + ctas - 1) / ctas;
    ctas = (kb_total + g.kb_per_cta[i] - 1) / g.kb_per_cta[i];
    g.cta_begin[i] = begin;
    begin += ctas;
  }
  g.cta_begin[g.n_problems] = begin;
  constexpr int smem = DW_STAGES * (DW_A_BYTES + 256 * 128) + 128;
  if (!(ctx->attr_done & NT_ATTR_DW_GROUPED)) {
    NT_CUDA(cudaFuncSetAttribute(dw_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_done |= NT_ATTR_DW_GROUPED;
  }
  int grid = ctx->sm_count;  // persistent: one CTA per SM walks the ~2 items per SM
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (grid > begin) grid = begin;
  dw_grouped_kernel<<<grid, GEMM_THREADS, smem, st>>>(g);
  NT_LAUNCH_CHECK(ctx);
  g.n_problems = 0;
  return NT_OK;
}

// 2-D bf16 tensor map [rows][cols] (row pitch ld elements), box = box_cols x box_rows, SWIZZLE_128B (shared with mlp_tc.cu)
int nt_make_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  return make_map(map, base, rows, cols, ld, box_cols, box_rows);
}
