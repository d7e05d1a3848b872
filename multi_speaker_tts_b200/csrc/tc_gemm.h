// Hand-written tcgen05 GEMM over pre-tiled bf16 operands (tc_gemm.cu).
#pragma once
#include "common.cuh"

struct TcGemmParams {
  const uint8_t* A;   // [Mt][Kb][hi | lo][16 KB]
  const uint8_t* B;   // [Nt][Kb][hi | lo][32 KB]
  float* C;           // fp32 output (mode 0), row-major with ldc
  int M, Mt, Nt, Kb, ldc;
  // fused-epilogue operands (WaveGlow modes)
  const float *bias0, *bias1;
  float *out_f32, *skip;
  uint8_t* out_img;
  const float* g_in;
  int T, Tp, dil, first, lastl;
  // split-K tail: the first n_full tiles run whole; each of the remaining n_split tiles is computed as two K halves on two
  // CTAs of the same round (the writer parks its fp32 accumulator in `partial`, the finisher adds it in its epilogue)
  int n_full, n_split;
  float* partial;     // [n_split][256 cols][128 rows]
  unsigned* flags;    // [n_split], zeroed before the launch; the 8 writer warps each add 1
};
// scratch bytes a caller must provide for the split-K tail (partials + flags)
constexpr size_t kTcGemmScratchBytes = (size_t)74 * 256 * 128 * 4 + 1024;

template <int MODE>
__device__ __forceinline__ void tc_gemm_epilogue(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part);

// WaveGlow WN layer, fused epilogues (tc_gemm.cu):
//   gate : A = im2col image [3 taps x 512 | mel 640] (K = 2176, hi and lo tiles), B columns permuted so that n-tile j
//          holds tanh channels 128j.. and the matching sigmoid channels; epilogue = bias + tanh * sigmoid -> the operand
//          image of the res/skip GEMM (K = 512)
//   res  : A = that image, B = res/skip kernel (N = 1024, or 512 in the last layer); epilogue = residual onto the gated
//          activation -> the NEXT layer's im2col taps (dilation dil), and skip accumulation
int tc_gemm_wn_gate(cudaStream_t s, const void* A1, const void* B1, int M, const float* b_in, const float* b_cond, float* g_f32, void* A2,
                    int T, int Tp, void* scratch);
int tc_gemm_wn_res(cudaStream_t s, const void* A2, const void* B2, int M, const float* b_res, const float* g_f32, float* skip, void* A1_next,
                   int T, int Tp, int dil_next, int first, int lastl, void* scratch);
constexpr int kWnK1 = 3 * 512 + 640;  // 2176: three taps of the dilated conv + the conditioning
constexpr int kWnK2 = 512;
// byte offset of the 16-byte chunk (row m, K index k with k % 8 == 0, hl = 0 hi tile / 1 lo tile) inside an A image with Kb
// k-blocks per m-tile
__host__ __device__ inline size_t tc_a_chunk_offset(int m, int k, int hl, int Kb) {
  return (((size_t)(m >> 7) * Kb + (k >> 6)) * 2 + hl) * 16384 + (size_t)((m & 127) >> 3) * 1024 + (size_t)((k & 63) >> 3) * 128 +
         (size_t)(m & 7) * 16;
}
#ifdef __CUDACC__
// One (row, 8-channel) piece of a layer input h in the im2col image of a dilated k=3 conv: the value goes to the three rows
// that read it (tap k of row m - (k-1) d), hi and lo tiles; the row's own taps whose source falls outside the utterance are zeroed.
__device__ __forceinline__ void wn_store_taps(uint8_t* img, int m, int t, int T, int d, int ch, const float (&h)[8]) {
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2bfloat16_rn(h[j]);
    lo[j] = __float2bfloat16_rn(h[j] - __bfloat162float(hi[j]));
  }
  const uint4 hv = *reinterpret_cast<const uint4*>(hi), lv = *reinterpret_cast<const uint4*>(lo), zv = make_uint4(0u, 0u, 0u, 0u);
  constexpr int Kb = kWnK1 / 64;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int tp = t - (k - 1) * d;  // time of the row that reads this value through tap k
    if (tp >= 0 && tp < T) {
      const int md = m - (k - 1) * d;
      *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(md, k * 512 + ch, 0, Kb)) = hv;
      *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(md, k * 512 + ch, 1, Kb)) = lv;
    }
    const int ts = t + (k - 1) * d;  // source time of this row's own tap k
    if (ts < 0 || ts >= T) {
      *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(m, k * 512 + ch, 0, Kb)) = zv;
      *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(m, k * 512 + ch, 1, Kb)) = zv;
    }
  }
}
// 8 values of row m at K index k (k % 8 == 0) into an A image with Kb k-blocks: hi tile and lo tile
__device__ __forceinline__ void wn_store_hl(uint8_t* img, int m, int Kb, int k, const float (&x)[8]) {
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2bfloat16_rn(x[j]);
    lo[j] = __float2bfloat16_rn(x[j] - __bfloat162float(hi[j]));
  }
  *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(m, k, 0, Kb)) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(img + tc_a_chunk_offset(m, k, 1, Kb)) = *reinterpret_cast<const uint4*>(lo);
}
#endif

int tc_gemm_plain(cudaStream_t s, const void* A_tiled, const void* B_tiled, float* C, int M, int N, int K, int ldc, void* scratch);
