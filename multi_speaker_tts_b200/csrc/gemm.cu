// Dense products outside the recurrent loops (hoisted prenet / projection / memory-layer products, weight gradients,
// WaveGlow reverse pass, convolutions).  Row-major in, row-major out.  Every entry point runs on the hand-written tcgen05
// kernel of tc_gemm.cu (bf16x3: operands split hi + lo, three partial products accumulated in fp32 in TMEM); this
// library links no vendor GEMM.
#include "gemm.h"

#include "tc_gemm.h"

int gemm_rowmajor_p(cudaStream_t s, int prec, bool transA, bool transB, int M, int N, int K, const float* A, int lda, const float* B,
                    int ldb, float* C, int ldc, float beta) {
  return tc_gemm_f32(s, transA, transB, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, beta, 1, prec);
}

int gemm_rowmajor_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                     const float* B, int ldb, float* C, int ldc, float beta) {
  return tc_gemm_f32(s, transA, transB, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, beta, 1, TC_FAST);
}

int gemm_rowmajor_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                          long long sA, const float* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta,
                          int batch) {
  return tc_gemm_f32(s, transA, transB, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, beta, batch, TC_FAST);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ src, size_t rows, size_t cols, size_t ld, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
  const size_t n = rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    const float x = src[r * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// 4 elements per thread: one 16-byte load, two 8-byte stores (cols, ld multiples of 4, 16-byte aligned pointers)
__global__ void split_bf16_vec4_kernel(const float* __restrict__ src, size_t rows, size_t cols4, size_t ld, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo) {
  const size_t n = rows * cols4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols4, c = (i - r * cols4) * 4;
    const float4 x = *reinterpret_cast<const float4*>(src + r * ld + c);
    __align__(8) __nv_bfloat16 h[4], l[4];
    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(xv[j]);
      l[j] = __float2bfloat16_rn(xv[j] - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint2*>(hi + i * 4) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i * 4) = *reinterpret_cast<const uint2*>(l);
  }
}

// stacked operands for a single-call bf16x3 product (the three partial products folded into K):
//   A side: dst[r][0:C] = hi, [C:2C] = lo, [2C:3C] = hi          B side: dst rows [0:R] = hi, [R:2R] = hi, [2R:3R] = lo
__global__ void split_bf16_stack_kernel(const float* __restrict__ src, size_t rows, size_t cols, size_t ld, __nv_bfloat16* __restrict__ dst,
                                        int b_side) {
  const size_t n = rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    const float x = src[r * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    if (!b_side) {
      __nv_bfloat16* d = dst + r * 3 * cols + c;
      d[0] = h;
      d[cols] = l;
      d[2 * cols] = h;
    } else {
      __nv_bfloat16* d = dst + r * cols + c;
      d[0] = h;
      d[rows * cols] = h;
      d[2 * rows * cols] = l;
    }
  }
}

int split_bf16_stack(cudaStream_t s, const float* src, size_t rows, size_t cols, size_t ld, __nv_bfloat16* dst, bool b_side) {
  const size_t n = rows * cols;
  if (n == 0) return MSTTS_OK;
  size_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  split_bf16_stack_kernel<<<(int)g, 256, 0, s>>>(src, rows, cols, ld, dst, b_side ? 1 : 0);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

int split_bf16_matrix(cudaStream_t s, const float* src, size_t rows, size_t cols, size_t ld, Bf16Pair dst) {
  const size_t n = rows * cols;
  if (n == 0) return MSTTS_OK;
  size_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (cols % 4 == 0 && ld % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst.hi & 7) == 0 && ((uintptr_t)dst.lo & 7) == 0) {
    size_t g4 = (n / 4 + 255) / 256;
    if (g4 > 148 * 16) g4 = 148 * 16;
    split_bf16_vec4_kernel<<<(int)g4, 256, 0, s>>>(src, rows, cols / 4, ld, dst.hi, dst.lo);
    MSTTS_CUDA(cudaGetLastError());
    return MSTTS_OK;
  }
  split_bf16_kernel<<<(int)g, 256, 0, s>>>(src, rows, cols, ld, dst.hi, dst.lo);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

int gemm_rowmajor_x3(cudaStream_t s, bool transA, bool transB, int M, int N, int K, Bf16Pair A, int lda, Bf16Pair B, int ldb,
                     float* C, int ldc, float beta) {
  return tc_gemm_hl(s, transA, transB, M, N, K, A.hi, A.lo, lda, 0, B.hi, B.lo, ldb, 0, C, ldc, 0, beta, 1);
}

int gemm_hl_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, Bf16Pair A, int lda, long long sA, Bf16Pair B, int ldb,
                    long long sB, float* C, int ldc, long long sC, float beta, int batch) {
  return tc_gemm_hl(s, transA, transB, M, N, K, A.hi, A.lo, lda, sA, B.hi, B.lo, ldb, sB, C, ldc, sC, beta, batch);
}
