// Plain dense GEMMs that sit outside the recurrent loop (hoisted prenet / projection / weight-gradient
// products).  Row-major in, row-major out, fp32.
#pragma once
#include "common.cuh"

// C[M,N] = op(A) * op(B) + beta * C, all row-major with leading dimensions lda/ldb/ldc.
// op(A) is M x K (A stored K x M when transA), op(B) is K x N (B stored N x K when transB).
int gemm_rowmajor_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                     const float* B, int ldb, float* C, int ldc, float beta);

static inline int gemm_rowmajor(cudaStream_t s, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, float beta) {
  return gemm_rowmajor_ex(s, false, false, M, N, K, A, lda, B, ldb, C, ldc, beta);
}


// batched: for i in [0,batch): C_i = op(A_i) op(B_i) + beta C_i with element strides sA/sB/sC between batches
int gemm_rowmajor_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                          long long sA, const float* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta,
                          int batch);

// ---- bf16x3 library GEMMs (same numerics contract as the recurrent tcgen05 kernels) -------------------------
// An fp32 matrix is split once into bf16 hi + bf16 lo; C = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi on the tensor cores
// (three cublasGemmEx calls, fp32 accumulation), ~16 mantissa bits per operand.
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

// one bf16 x bf16 -> fp32 tensor-core GEMM (row-major, no transposes): C = A B + beta C.  Callers that fold the three
// bf16x3 partial products into K (A rows [hi|lo|hi], B rows [hi;hi;lo]) get the bf16x3 result from a single call.
int gemm_rowmajor_bf16(cudaStream_t s, int M, int N, int K, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                       float* C, int ldc, float beta);
// general form: C[M,N] = op(A) op(B) + beta C with op(A) M x K (stored K x M when transA), op(B) K x N (stored N x K when transB)
int gemm_bf16_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B,
                 int ldb, float* C, int ldc, float beta);
// strided-batched form (element strides sA / sB / sC between batches; sB = 0 shares B)
int gemm_bf16_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A, int lda, long long sA,
                      const __nv_bfloat16* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta, int batch);
