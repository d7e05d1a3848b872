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
  // general front-end (MODE 3, tc_gemm_f32 / tc_gemm_hl): `batch` independent products with byte strides between the operand
  // images and an element stride between the outputs; N valid output columns (the last n-tile may be partial); C = acc +
  // beta * C.  ksplit > 1: work item = (tile, K slice); slice s of batch b stores its fp32 accumulator to the slab
  // slabs[(s * batch + b)][Mt * 128][Nt * 256] and a second kernel sums the slabs in a fixed order (deterministic split-K).
  int batch, N, ksplit;
  size_t a_bstride, b_bstride, c_bstride;
  float beta;
  float* slabs;
  int stream_l2;      // 1: operand images are loaded evict_first (a product running BESIDE a persistent loop must not displace
                      // the weight image the loop keeps in L2)
};
// scratch bytes a caller must provide for the split-K tail (partials + flags)
constexpr size_t kTcGemmScratchBytes = (size_t)74 * 256 * 128 * 4 + 1024;

// work item -> (tile, K-block range, role): role 0 = whole tile (or one K slice of a ksplit product), 1 = writer (first K
// half of a tail tile), 2 = finisher (second half)
struct TcItem {
  int tile, kb0, kb1, role, split;
  int b, mt, nt;  // batch index, m-tile, n-tile
};
template <int MODE>
__device__ __forceinline__ void tc_gemm_epilogue(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part,
                                                 const TcItem& w);

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

// ---- general row-major front-ends (every dense product outside the recurrent loops; no library GEMM anywhere) ----
// C_b[M,N] = op(A_b) op(B_b) + beta C_b for b in [0, batch), all row-major; op(A) is M x K (A stored K x M when transA),
// op(B) is K x N (B stored N x K when transB); sA / sB / sC are element strides between batches (sB = 0 shares B).
// The operands are packed into the tensor core's tile images (hi + lo bf16, zero padded to 128 / 256 rows and 64-wide
// k-blocks) by a pack kernel, multiplied as bf16x3 by tc_gemm_kernel<3>, and (for few-tile / long-K shapes) summed over
// K slices by a deterministic reduce kernel.  Scratch comes from the library's stream-ordered pool (scratch_pool.h).
// Precision levels.  The fp32 accumulator of the tensor core truncates, so the error of one accumulation chain grows linearly
// with its length; the operand split bounds the error per product.
//   TC_FAST    bf16x3 (operands split hi + lo, three products), one chain over the whole K: ~1e-5 of max at K ~ 10^3
//   TC_CHAINED bf16x3, chains of <= 8 k-blocks (K = 512) summed in double by the reduce kernel
//   TC_PRECISE operands split THREE ways (h + m + l, 24 mantissa bits), six partial products (hh + hm + mh + mm + hl + lh, as
//              four K segments of the same kernel), chains of <= 4 k-blocks: fp32-SGEMM accuracy at 4x the tensor work -- for
//              the products whose fp32 result is small against the magnitude of its terms (weight gradients of narrow layers)
enum { TC_FAST = 0, TC_CHAINED = 1, TC_PRECISE = 2 };
int tc_gemm_f32(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda, long long sA, const float* B,
                int ldb, long long sB, float* C, int ldc, long long sC, float beta, int batch, int prec);
// building blocks for call sites that share a packed operand between products or assemble one from several matrices:
// an image holds ceil(R / TR) tile rows (TR = 128 for the A side, 256 for the B side) x Kb k-blocks of [hi | lo] tiles
static inline size_t tc_image_bytes(int R, int K, int TR, bool precise = false) {
  return (size_t)((R + TR - 1) / TR) * ((K + 63) / 64) * (precise ? 4 : 1) * 2 * (size_t)TR * 128;
}
// pack R rows x K (element (r, k) at src[r * ld + k], or src[k * ld + r] when trans) at tile row rt0 / k-block kb0 of an
// image with Kb_total k-blocks; rows up to the next multiple of TR and k up to the next multiple of 64 are zero filled
int tc_pack_f32(cudaStream_t s, const float* src, int ld, bool trans, int R, int K, int TR, int Kb_total, void* img, int rt0, int kb0,
                bool precise = false);
// While alive, caps the grid of every launch this thread issues through the front-end (pack, product, reduce) at `ctas` SMs'
// worth of CTAs: products that run on a side stream BESIDE a persistent loop, on the SMs the loop leaves idle, must never
// hold more SMs than that or the loop's cooperative launch could not become resident.
struct TcGridCap {
  int prev;
  explicit TcGridCap(int ctas);
  ~TcGridCap();
};
// same from a matrix already split into bf16 hi + lo parts (hi and lo share the leading dimension)
int tc_pack_hl(cudaStream_t s, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, bool trans, int R, int K, int TR, int Kb_total, void* img,
               int rt0, int kb0);
// C[M,N] = A_img . B_img^T + beta C  (K = the images' k extent)
int tc_gemm_images(cudaStream_t s, const void* A_img, const void* B_img, int M, int N, int K, float* C, int ldc, float beta,
                   int prec = TC_FAST);
// same with operands already split into bf16 hi + lo row-major matrices (hi and lo share the leading dimension / strides)
int tc_gemm_hl(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
               int lda, long long sA, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, long long sB, float* C, int ldc,
               long long sC, float beta, int batch);
