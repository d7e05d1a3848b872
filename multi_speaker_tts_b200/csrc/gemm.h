// Dense products that sit outside the recurrent loops (hoisted prenet / projection / weight-gradient products).
// Row-major fp32 in, row-major fp32 out, on the hand-written tcgen05 kernel (tc_gemm.h): bf16x3 (operands split hi + lo, three
// partial products, fp32 accumulation in TMEM).  gemm_rowmajor_p takes a precision level (tc_gemm.h: TC_FAST / TC_CHAINED /
// TC_PRECISE) for the products whose result is small against its terms (weight gradients of the narrow layers).
#pragma once
#include "common.cuh"
#include "tc_gemm.h"

// C[M,N] = op(A) * op(B) + beta * C, all row-major with leading dimensions lda/ldb/ldc.
// op(A) is M x K (A stored K x M when transA), op(B) is K x N (B stored N x K when transB).
int gemm_rowmajor_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                     const float* B, int ldb, float* C, int ldc, float beta);

static inline int gemm_rowmajor(cudaStream_t s, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, float beta) {
  return gemm_rowmajor_ex(s, false, false, M, N, K, A, lda, B, ldb, C, ldc, beta);
}


int gemm_rowmajor_p(cudaStream_t s, int prec, bool transA, bool transB, int M, int N, int K, const float* A, int lda, const float* B,
                    int ldb, float* C, int ldc, float beta);

// batched: for i in [0,batch): C_i = op(A_i) op(B_i) + beta C_i with element strides sA/sB/sC between batches
int gemm_rowmajor_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                          long long sA, const float* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta,
                          int batch);

// ---- products over operands already split into bf16 hi + bf16 lo (C = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi, fp32
// accumulation in TMEM, ~16 mantissa bits per operand): the same hand-written kernel, the pack step skips the split
struct Bf16Pair {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
};
// dst.{hi,lo}[r*cols + c] = split(src[r*ld + c])
int split_bf16_matrix(cudaStream_t s, const float* src, size_t rows, size_t cols, size_t ld, Bf16Pair dst);
// stacked single-buffer split for a one-call bf16x3 product: A side [rows, 3*cols] = [hi|lo|hi], B side [3*rows, cols] = [hi;hi;lo]
int split_bf16_stack(cudaStream_t s, const float* src, size_t rows, size_t cols, size_t ld, __nv_bfloat16* dst, bool b_side);
int gemm_rowmajor_x3(cudaStream_t s, bool transA, bool transB, int M, int N, int K, Bf16Pair A, int lda, Bf16Pair B, int ldb,
                     float* C, int ldc, float beta);

// strided-batched form over hi/lo pairs (element strides between batches; sB = 0 shares B)
int gemm_hl_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, Bf16Pair A, int lda, long long sA, Bf16Pair B, int ldb,
                    long long sB, float* C, int ldc, long long sC, float beta, int batch);
