// Packed-weight layout and shared-memory geometry shared by the fused encode+MLP tensor-core kernels
// (mlp_tc.cu: bf16 / fp16 single pass; mlp_tc32.cu: 3-pass split-fp16).  Layer order L = 0..9: the 8 trunk layers
// (nerf.py:85-91), point_info (:96), dir_info (:98); the sigma and colour heads ride in the fp32 side block.
#pragma once
#include <stdint.h>

namespace mlpl {

constexpr int TILE_M = 128;
constexpr int CHUNK_A_BYTES = TILE_M * 128;      // [128 rows x 64 bf16], one swizzle row per sample
constexpr int ACT_BYTES = 4 * CHUNK_A_BYTES;     // 256-wide activations
constexpr int W_STAGE_BYTES = 256 * 128;         // [256 x 64] bf16
constexpr int N_STAGES = 2;
constexpr int OFF_ACT = 0;
constexpr int OFF_ENC = 2 * ACT_BYTES;
constexpr int OFF_W = OFF_ENC + 2 * CHUNK_A_BYTES;
constexpr int OFF_BAR = OFF_W + N_STAGES * W_STAGE_BYTES;
constexpr int AUX_REC_FLOATS = 640;                // per-layer fp32 record: bias[256] | head weights[384]
constexpr int OFF_AUX = OFF_BAR + 128;             // single-buffered stage for the current layer's record
constexpr int SMEM_BYTES = OFF_AUX + AUX_REC_FLOATS * 4;
constexpr int N_MMA_LAYERS = 10;  // L0..L7, point_info, dir_info
constexpr int N_EPI_WARPS = 16;   // per tile: 4 lane quadrants x 2 column halves
constexpr int WARP_TMA = 16, WARP_MMA = 17;
constexpr int N_THREADS = 576;    // 16 epilogue warps + TMA producer + MMA issuer

// barrier slots (8 B each) inside the OFF_BAR block
enum { BAR_W_FULL = 0, BAR_W_EMPTY = 2, BAR_ACC_FULL = 4, BAR_ACT_READY = 6, BAR_AUX_FULL = 8, BAR_COUNT = 9 };

__host__ __device__ constexpr int layer_chunks(int L) { return L == 0 ? 1 : ((L == 4 || L == 9) ? 5 : 4); }
__host__ __device__ constexpr int layer_n(int L) { return L == 9 ? 128 : 256; }
__host__ __device__ constexpr int chunk_bytes(int L) { return layer_n(L) * 128; }
constexpr int total_packed_bytes() {
  int b = 0;
  for (int L = 0; L < N_MMA_LAYERS; ++L) b += layer_chunks(L) * chunk_bytes(L);
  return b;
}
constexpr int PACKED_W_BYTES = total_packed_bytes();  // 1 196 032
// fp32 side block appended to the packed weights, 16-byte aligned (the flat parameter buffer is not: the
// 1-wide sigma bias shifts everything after it by one float)
// one record per tensor-core layer, TMA-copied into shared memory right before the layer's epilogue:
//   [0,256)  bias (dir_info: 128 biases, then the 3 colour biases at 128..130)
//   [256,640) head weights: layer 7 -> sigma weights [256] + sigma bias at 512; layer 9 -> colour weights [3][128]
constexpr int AUX_EXTRA = 256;
constexpr int AUX_SIG_B = 512;
constexpr int AUX_COL_B = 128;
constexpr int AUX_FLOATS = N_MMA_LAYERS * AUX_REC_FLOATS;
constexpr int PACKED_BYTES = PACKED_W_BYTES + AUX_FLOATS * 4;
__host__ __device__ constexpr int aux_bytes(int L) { return (L == 7 || L == 9) ? AUX_REC_FLOATS * 4 : 1024; }


}  // namespace mlpl
