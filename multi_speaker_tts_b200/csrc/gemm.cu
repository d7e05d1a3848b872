// cuBLAS SGEMM wrapper (library GEMM for the plain, non-recurrent products; full fp32 math so the
// parity mode stays at fp32 accuracy).  One handle per (thread, device).
#include <cublas_v2.h>

#include "gemm.h"

static cublasHandle_t get_handle(int* rc) {
  static thread_local cublasHandle_t handles[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    mstts_set_error("gemm: cudaGetDevice failed");
    *rc = MSTTS_E_CUDA;
    return nullptr;
  }
  if (!handles[dev]) {
    cublasStatus_t st = cublasCreate(&handles[dev]);
    if (st != CUBLAS_STATUS_SUCCESS) {
      mstts_set_error("gemm: cublasCreate failed (%d)", (int)st);
      *rc = MSTTS_E_CUDA;
      handles[dev] = nullptr;
      return nullptr;
    }
    cublasSetMathMode(handles[dev], CUBLAS_PEDANTIC_MATH);
  }
  *rc = MSTTS_OK;
  return handles[dev];
}

int gemm_rowmajor_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                     const float* B, int ldb, float* C, int ldc, float beta) {
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  const float alpha = 1.f;
  // row-major C = op(A) op(B)  <=>  column-major C^T = op(B)^T op(A)^T
  st = cublasSgemm(h, transB ? CUBLAS_OP_T : CUBLAS_OP_N, transA ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &alpha, B, ldb, A,
                   lda, &beta, C, ldc);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSgemm(M=%d,N=%d,K=%d) failed (%d)", M, N, K, (int)st);
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}

int gemm_rowmajor_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda,
                          long long sA, const float* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta,
                          int batch) {
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  const float alpha = 1.f;
  st = cublasSgemmStridedBatched(h, transB ? CUBLAS_OP_T : CUBLAS_OP_N, transA ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &alpha,
                                 B, ldb, sB, A, lda, sA, &beta, C, ldc, sC, batch);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSgemmStridedBatched(M=%d,N=%d,K=%d,batch=%d) failed (%d)", M, N, K, batch, (int)st);
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}
