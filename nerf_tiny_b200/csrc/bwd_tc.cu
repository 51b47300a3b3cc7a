// Fused backward-data chain of the 8x256 MLP on tcgen05 (NT_PREC_BF16 training): the mirror image of mlp_tc.cu.
// Autograd of Network.forward (nerf.py:101-124) w.r.t. the activations, SURVEY.md B.6/B.7:
//   step 0 : g_info = g_u . W_d[:, 24:]                                   (dir_info -> point_info, no activation)
//   step 1 : g_7    = relu'(h7) * (g_info . W_p + g_zsigma (x) w_sigma)    (point_info and the sigma head join here)
//   step k : g_{8-k} = relu'(h_{8-k}) * (g_{9-k} . W_{9-k}[:, :256])       k = 2..8  (trunk layers 7..1)
// One persistent CTA per SM, pairs of 128-sample tiles in lock-step exactly like the forward kernel: the gradient tile
// stays in shared memory as the bf16 SWIZZLE_128B A operand of the next step, transposed weights stream through a
// 2-stage TMA ring in pre-packed K-chunks, accumulators live in TMEM.  The epilogue applies the ReLU' BIT mask that the
// forward pass stashed (32 B per sample-layer instead of re-reading 512 B of activations), accumulates the bias
// gradient (column sums by shuffle recursive-halving) and writes the next operand in place; every step's gradient
// tile is also stashed to HBM by TMA tensor stores for the grouped weight-gradient GEMM (gemm_tc.cu).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_A_BYTES = TILE_M * 128;
constexpr int ACT_BYTES = 4 * CHUNK_A_BYTES;
constexpr int W_STAGE_BYTES = 256 * 128;
constexpr int N_STAGES = 3;  // 3 x 32 KB: without the forward kernel's feature tiles there is room for a third slot
constexpr int N_STEPS = 9;
constexpr int OFF_ACT = 0;
constexpr int OFF_W = 2 * ACT_BYTES;
constexpr int OFF_BAR = OFF_W + N_STAGES * W_STAGE_BYTES;
constexpr int OFF_WSIG = OFF_BAR + 256;  // w_sigma row, fp32 [256]
constexpr int SMEM_BYTES = OFF_WSIG + 1024;
constexpr int N_THREADS = 640;  // 16 epilogue warps + TMA producer + MMA issuer + 2 column-sum warps (one per tile)
constexpr int WARP_TMA = 16, WARP_MMA = 17, WARP_CS = 18;
enum { BAR_W_FULL = 0, BAR_W_EMPTY = 3, BAR_ACC_FULL = 6, BAR_ACT_READY = 8, BAR_A_FULL = 10, BAR_CS_DONE = 12 };

__host__ __device__ constexpr int step_chunks(int st) { return st == 0 ? 2 : 4; }
constexpr int total_chunks() {
  int c = 0;
  for (int st = 0; st < N_STEPS; ++st) c += step_chunks(st);
  return c;
}
constexpr int PACKED_BWD_BYTES = total_chunks() * W_STAGE_BYTES;  // 34 x 32 KB

struct BwdParams {
  CUtensorMap map_gu;             // load:  g_u  [S][128]
  CUtensorMap map_out[N_STEPS];   // store: g_info, g_7 .. g_0  [S][256]
  const uint8_t* packed;          // transposed bf16 weights in consumption order
  const uint32_t* bits;           // ReLU' bit masks [8][S][8] (layer, sample, 256 bits; column cb*32+j = bit 31-j of word cb)
  const float* gzsig;             // [S]
  const float* wsig;              // [256] fp32 (flat parameter buffer)
  float* db[N_STEPS];             // bias gradients: point_info, layer 7 .. layer 0
  int64_t total;
  int num_pairs;
  int num_items, full_pairs;      // work items: pairs below full_pairs, then single tiles (see mlp_tc.cu TcParams)
  long long* prof;                // optional clock64 timeline of block 0's first pairs (tools/chain_timeline.py)
};

// PTX wrappers: tc_ptx.cuh
using namespace tcptx;
// diagnostic timeline: prof[(pair_local * 9 + step) * 16 + slot], block 0, first 4 pairs
#define BW_PROF(pl_, st_, slot_)                                                          \
  do {                                                                                    \
    if (P.prof && blockIdx.x == 0 && (pl_) < 4) P.prof[((pl_) * 9 + (st_)) * 16 + (slot_)] = clock64(); \
  } while (0)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return desc_kmajor(saddr); }
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return idesc_f16kind(TILE_M, n, true); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  umma_f16(d_tmem, adesc, bdesc, idesc, acc);
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THREADS, 1) bwd_tc_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 128);

  if (threadIdx.x == 0) {
    if (sbase & 1023) __trap();
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(bar(BAR_W_FULL + s), 1);
      mbar_init(bar(BAR_W_EMPTY + s), 1);
    }
    for (int tl = 0; tl < 2; ++tl) {
      mbar_init(bar(BAR_ACC_FULL + tl), 1);
      mbar_init(bar(BAR_ACT_READY + tl), 2 * TILE_M);
      mbar_init(bar(BAR_A_FULL + tl), 1);
      mbar_init(bar(BAR_CS_DONE + tl), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // w_sigma row for the rank-1 term of step 1
  if (threadIdx.x < 256) reinterpret_cast<float*>(smem + OFF_WSIG)[threadIdx.x] = P.wsig[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ===================== weight producer =====================
    // TMA bulk copies with an evict_last L2 policy: ~9 KB of stash per sample stream through L2 while every CTA re-reads the
    // same 1.1 MB of transposed weights; without the hint the stash evicted them and a 32 KB chunk took 2 000 - 8 500 clk to
    // arrive (clock64 timeline), with it the weights stay L2-resident.
    if (lane == 0) {
      uint32_t q = 0;
      int pl = 0;
      const uint64_t pol_w = l2_policy_evict_last();
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++pl) {
        const uint8_t* src = P.packed;
        for (int st = 0; st < N_STEPS; ++st)
          for (int kc = 0; kc < step_chunks(st); ++kc, ++q) {
            const uint32_t stage = q % N_STAGES;
            mbar_wait(bar(BAR_W_EMPTY + stage), ((q / N_STAGES) & 1) ^ 1);
            if (kc == 1) BW_PROF(pl, st, 8);
            if (kc == 3) BW_PROF(pl, st, 9);
            mbar_expect_tx(bar(BAR_W_FULL + stage), W_STAGE_BYTES);
            tma_bulk_g2s_hint(sbase + OFF_W + stage * W_STAGE_BYTES, src, W_STAGE_BYTES, bar(BAR_W_FULL + stage), pol_w);
            src += W_STAGE_BYTES;
          }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer (+ the g_u operand loads) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(256);
      uint32_t q = 0, lit = 0, pl = 0;
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++pl) {
        // a single-tile item is always this CTA's last one: tile B's barrier phases may fall behind `lit` from there on
        const bool single = item >= P.full_pairs;
        const int tile0 = single ? 2 * P.full_pairs + (item - P.full_pairs) : 2 * item;
        const int ntl = single ? 1 : 2;
        for (int st = 0; st < N_STEPS; ++st, ++lit) {
          const int nch = step_chunks(st);
          for (int kc = 0; kc < nch; ++kc, ++q) {
            const uint32_t stage = q % N_STAGES;
            mbar_wait(bar(BAR_W_FULL + stage), (q / N_STAGES) & 1);
            tc_fence_after();
            const uint32_t b_addr = sbase + OFF_W + stage * W_STAGE_BYTES;
            BW_PROF(pl, st, 2 + kc);
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
              if (tl >= ntl) continue;
              const uint32_t act = sbase + OFF_ACT + tl * ACT_BYTES;
              if (kc == 0) {
                mbar_wait(bar(BAR_ACT_READY + tl), lit & 1);  // epilogue done: operand in place, accumulator drained
                tc_fence_after();
                BW_PROF(pl, st, tl);
                if (st == 0) {
                  // fetch this pair's g_u tile over the previous pair's g_0 tile: its stash stores have left shared memory
                  // (the epilogue's storing threads wait for that before the arrive of the last step) and its column sums
                  // have been taken
                  const int row0 = (tile0 + tl) * TILE_M;
                  if (lit > 0) mbar_wait(bar(BAR_CS_DONE + tl), (lit - 1) & 1);
                  mbar_expect_tx(bar(BAR_A_FULL + tl), 2 * CHUNK_A_BYTES);
                  tma_load_2d(act, &P.map_gu, 0, row0, bar(BAR_A_FULL + tl));
                  tma_load_2d(act + CHUNK_A_BYTES, &P.map_gu, 64, row0, bar(BAR_A_FULL + tl));
                  mbar_wait(bar(BAR_A_FULL + tl), pl & 1);
                }
              }
              const uint32_t a_addr = act + kc * CHUNK_A_BYTES;
              const uint32_t d_tmem = tmem_base + tl * 256;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_bf16(d_tmem, umma_desc(a_addr + j * 32), umma_desc(b_addr + j * 32), idesc, (kc | j) != 0);
              if (kc == nch - 1) {
                umma_commit(bar(BAR_ACC_FULL + tl));
                BW_PROF(pl, st, 6 + tl);
              }
            }
            umma_commit(bar(BAR_W_EMPTY + stage));
          }
        }
      }
    }
  } else if (warp >= WARP_CS) {
    // ===================== bias gradients: column sums of every step's gradient tile, straight from shared memory ==========
    // db_l = sum over samples of g_l.  The epilogue used to reduce its fp32 values over the warp's 32 rows with 124 shuffles
    // per thread and step (three quarters of its instructions); here one otherwise idle warp per tile reads the finished
    // bf16 operand tile (the very values the weight-gradient GEMM multiplies) while the tensor core consumes it, and adds the
    // 128 rows in fp32: lane = one 16-byte chunk (8 columns) of the 512-byte row, one atomic per column and tile.
    const int tl = warp - WARP_CS;
    const uint32_t act = sbase + OFF_ACT + tl * ACT_BYTES;
    const uint32_t base = act + (lane >> 3) * CHUNK_A_BYTES;
    const int j = lane & 7;
    uint32_t it = 0;
    mbar_wait(bar(BAR_ACT_READY + tl), 0);  // completion #0 is the epilogue warps' start-up arrive (no data behind it)
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
      if (item >= P.full_pairs && tl == 1) break;  // single-tile item: no tile B
      for (int st = 0; st < N_STEPS; ++st, ++it) {
        mbar_wait(bar(BAR_ACT_READY + tl), (it + 1) & 1);  // completion #it+1: the output of step `it` is in place
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int r8 = 0; r8 < TILE_M / 8; ++r8) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(base + (r8 * 8 + r) * 128 + ((j ^ r) << 4)));
            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[2 * e] += __uint_as_float(w[e] << 16);
              acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
            }
          }
        }
        fence_proxy_async();  // these generic-proxy reads precede the next pair's TMA load into the same tile
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_CS_DONE + tl));  // the epilogue / the next pair's load may overwrite the tile
        if (lane == 0) BW_PROF((int)(it / N_STEPS), st, 14 + tl);
        float* __restrict__ db = P.db[st] + lane * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(db + e, acc[e]);
      }
    }
  } else {
    // ===================== epilogue warps: per tile 4 lane quadrants x 2 column halves =====================
    // Every step's gradient tile is also the weight-gradient GEMM's operand: the 128 threads of a (tile, column half) group
    // hand each 64-column chunk to the TMA unit as soon as they have written it (one 16 KB tensor store per chunk, issued by
    // the group's elected thread after a 128-thread named barrier), so the stores drain to HBM during the rest of the
    // epilogue AND the next step's MMAs instead of holding the MMA thread before its commit (2 800 of 10 300 clk per step
    // in the clock64 timeline when the MMA thread issued and awaited them).
    const int tl = (warp >> 2) & 1;
    const int half = warp >> 3;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t act = sbase + OFF_ACT + tl * ACT_BYTES;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + tl * 256;
    const uint32_t row_off = row * 128, x4 = (row & 7) << 4;
    const uint32_t wsig_s = sbase + OFF_WSIG;
    const int grp_bar = 1 + tl * 2 + half;                 // named barrier of this (tile, half) group
    const bool storer = (warp & 3) == 0 && lane == 0;      // the group's elected thread
    const uint64_t pol_st = l2_policy_evict_first();       // stash lines leave L2 first (they are read once, much later)
    uint32_t it = 0;
    int pl = 0;
    const bool stamper = (warp & 3) == 0 && half == 0 && lane == 0;
    mbar_arrive(bar(BAR_ACT_READY + tl));  // nothing to protect before the first pair
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x, ++pl) {
      const bool single = item >= P.full_pairs;
      if (single && tl == 1) break;  // single-tile item (this CTA's last): tile B's warps have nothing to do
      const int row0 = ((single ? 2 * P.full_pairs + (item - P.full_pairs) : 2 * item) + tl) * TILE_M;
      const int64_t s = (int64_t)row0 + row;
      const bool valid = s < P.total;
      const float gz = valid ? __ldg(P.gzsig + s) : 0.f;
      for (int st = 0; st < N_STEPS; ++st, ++it) {
        // ReLU' bits of this thread's 128 columns (4 words), fetched while the MMAs run
        uint4 mw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (st >= 1) {
          const int ml = 8 - st;
          mw = valid ? __ldg(reinterpret_cast<const uint4*>(P.bits + ((int64_t)ml * P.total + s) * 8 + half * 4))
                     : make_uint4(0u, 0u, 0u, 0u);
        }
        const uint32_t mword[4] = {mw.x, mw.y, mw.z, mw.w};
        mbar_wait(bar(BAR_ACC_FULL + tl), it & 1);
        tc_fence_after();
        if (it > 0) mbar_wait(bar(BAR_CS_DONE + tl), (it - 1) & 1);  // the tile about to be overwritten has been column-summed
        // ... and the previous step's stash stores of this group's two chunks have finished reading it (normally long ago)
        if (storer) tma_store_wait_read0();
        asm volatile("bar.sync %0, 128;" ::"r"(grp_bar) : "memory");
        if (stamper) BW_PROF(pl, st, 10 + tl);
        uint32_t buf[2][32];
        tmem_ld32_issue(tmem_row + half * 128, buf[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cb = half * 4 + i;
          uint32_t(&raw)[32] = buf[i & 1];
          tmem_ld_wait(raw);
          if (i + 1 < 4) tmem_ld32_issue(tmem_row + (cb + 1) * 32, buf[(i + 1) & 1]);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (st == 1) {  // + g_zsigma (x) w_sigma : the sigma head's path into h7 (nerf.py:114)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = lds128(wsig_s + (cb * 32 + 4 * j) * 4);
              v[4 * j + 0] = fmaf(gz, w4.x, v[4 * j + 0]);
              v[4 * j + 1] = fmaf(gz, w4.y, v[4 * j + 1]);
              v[4 * j + 2] = fmaf(gz, w4.z, v[4 * j + 2]);
              v[4 * j + 3] = fmaf(gz, w4.w, v[4 * j + 3]);
            }
          }
          const uint32_t m32 = mword[i];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((int)(m32 << j) < 0) ? v[j] : 0.f;  // element j sits at bit 31-j
          // next operand, in place (K-chunk cb/2, 16-byte chunks (cb%2)*4 .. +3)
          const uint32_t dst = act + (cb >> 1) * CHUNK_A_BYTES + row_off;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            st_shared_v4(dst + (x4 ^ (uint32_t)(((cb & 1) * 4 + qd) << 4)), pack_bf16(v[8 * qd], v[8 * qd + 1]),
                         pack_bf16(v[8 * qd + 2], v[8 * qd + 3]), pack_bf16(v[8 * qd + 4], v[8 * qd + 5]),
                         pack_bf16(v[8 * qd + 6], v[8 * qd + 7]));
          if (i & 1) {
            // K-chunk half*2 + i/2 of the new tile is complete: make it visible to the async proxy and stash it
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(grp_bar) : "memory");
            if (storer && row0 < P.total) {
              const int kc = half * 2 + (i >> 1);
              tma_store_2d_hint(&P.map_out[st], kc * 64, row0, act + kc * CHUNK_A_BYTES, pol_st);
              tma_store_commit();
            }
          }
        }
        // the next pair's g_u load overwrites chunks 0-1 of this tile right after the arrive of the LAST step
        if (st == N_STEPS - 1 && storer) tma_store_wait_read0();
        if (st == N_STEPS - 1) asm volatile("bar.sync %0, 128;" ::"r"(grp_bar) : "memory");
        tc_fence_before();
        mbar_arrive(bar(BAR_ACT_READY + tl));
        if (stamper) BW_PROF(pl, st, 12 + tl);
      }
    }
    if (storer) tma_store_wait0();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- transposed weight pack: [256 (n = input column of the layer)][64 (k = output row)] bf16 chunks, SW128 ------------
struct PackBwdArgs {
  int w_off[NT_N_LAYERS];
  int in_f[NT_N_LAYERS];
};

__global__ void pack_bwd_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed, PackBwdArgs a) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  if (gid >= PACKED_BWD_BYTES / 16) return;
  int byte = gid * 16, st = 0;
  for (st = 0; st < N_STEPS; ++st) {
    const int sb = step_chunks(st) * W_STAGE_BYTES;
    if (byte < sb) break;
    byte -= sb;
  }
  const int kc = byte / W_STAGE_BYTES;
  byte -= kc * W_STAGE_BYTES;
  const int n = byte / 128;
  const int j = ((byte % 128) / 16) ^ (n & 7);
  uint32_t out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = kc * 64 + j * 8 + e * 2 + h;  // output row of the forward layer
      if (st == 0)
        v[h] = params[a.w_off[L_DIR] + k * 280 + 24 + n];       // (W_d[:, 24:])^T, k < 128
      else if (st == 1)
        v[h] = params[a.w_off[L_INFO] + k * 256 + n];
      else {
        const int i = 9 - st;                                     // trunk layer 7 .. 1
        v[h] = params[a.w_off[i] + (int64_t)k * a.in_f[i] + n];   // first 256 input columns (hidden part of the skip)
      }
    }
    __nv_bfloat162 b = __floats2bfloat162_rn(v[0], v[1]);
    out[e] = *reinterpret_cast<uint32_t*>(&b);
  }
  *reinterpret_cast<uint4*>(packed + (size_t)gid * 16) = make_uint4(out[0], out[1], out[2], out[3]);
}

}  // namespace

size_t nt_bwd_tc_packed_bytes() { return PACKED_BWD_BYTES; }

static long long* g_chain_prof = nullptr;
// diagnostics only (tools/chain_timeline.py): device buffer of 4 * 9 * 16 int64 that the next chain launches stamp
extern "C" void nt_debug_set_chain_prof(void* p) { g_chain_prof = (long long*)p; }

int nt_bwd_tc_pack(nt_ctx* ctx, const float* params, void* packed, cudaStream_t st) {
  const LayerTable T = nt_layers();
  PackBwdArgs a;
  for (int i = 0; i < NT_N_LAYERS; ++i) {
    a.w_off[i] = (int)T.w[i];
    a.in_f[i] = kLayerIn[i];
  }
  const int threads = PACKED_BWD_BYTES / 16;
  pack_bwd_kernel<<<(threads + 255) / 256, 256, 0, st>>>(params, (uint8_t*)packed, a);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}

// g_u bf16 [S][128] -> outs[0] = g_info, outs[1..8] = g_7 .. g_0 (bf16 [S][256] each); db[k] accumulates colsum(outs[k])
int nt_bwd_tc_chain(nt_ctx* ctx, int64_t S, const void* g_u, void* const outs[9], const int out_ld[9], const void* packed,
                    int max_ctas, const uint32_t* bits,
                    const float* gzsig, const float* wsig, float* const db[9], cudaStream_t st) {
  if (S <= 0) return NT_OK;
  if (!(ctx->attr_done & NT_ATTR_BWD_TC)) {
    NT_CUDA(cudaFuncSetAttribute(bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ctx->attr_done |= NT_ATTR_BWD_TC;
  }
  BwdParams P;
  memset(&P, 0, sizeof(P));
  int rc = nt_make_map_bf16(&P.map_gu, g_u, S, 128, 128, 64, TILE_M);
  if (rc != NT_OK) return rc;
  for (int i = 0; i < N_STEPS; ++i) {
    rc = nt_make_map_bf16(&P.map_out[i], outs[i], S, 256, out_ld[i], 64, TILE_M);
    if (rc != NT_OK) return rc;
    P.db[i] = db[i];
  }
  P.packed = (const uint8_t*)packed;
  P.bits = bits;
  P.gzsig = gzsig;
  P.wsig = wsig;
  P.total = S;
  P.prof = g_chain_prof;
  const int64_t tiles = (S + TILE_M - 1) / TILE_M;
  P.num_pairs = (int)((tiles + 1) / 2);
  int G = ctx->sm_count;
  if (max_ctas > 0 && max_ctas < G) G = max_ctas;  // leave SMs to a concurrent launch on another stream
  // tile pairs, except that a last wave which would occupy at most half of the CTAs runs as single tiles
  P.full_pairs = P.num_pairs;
  P.num_items = P.num_pairs;
  {
    const int64_t waves = tiles / (2 * (int64_t)G), rest = tiles - 2 * (int64_t)G * waves;
    if (rest > 0 && rest <= G && !getenv("NT_NO_SINGLE_TILES")) {
      P.full_pairs = (int)(G * waves);
      P.num_items = (int)(G * waves + rest);
    }
  }
  const int grid = G < P.num_items ? G : P.num_items;
  bwd_tc_kernel<<<grid, N_THREADS, SMEM_BYTES, st>>>(P);
  NT_LAUNCH_CHECK(ctx);
  return NT_OK;
}
