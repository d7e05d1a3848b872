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
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);  // the handle is pedantic for the fp32 path
  const float one = 1.f;
  const cublasOperation_t opB = transB ? CUBLAS_OP_T : CUBLAS_OP_N, opA = transA ? CUBLAS_OP_T : CUBLAS_OP_N;
  // small terms first, the dominant hi.hi product last
  const __nv_bfloat16* a_ops[3] = {A.lo, A.hi, A.hi};
  const __nv_bfloat16* b_ops[3] = {B.hi, B.lo, B.hi};
  for (int i = 0; i < 3; ++i) {
    const float bt = i == 0 ? beta : 1.f;
    st = cublasGemmEx(h, opB, opA, N, M, K, &one, b_ops[i], CUDA_R_16BF, ldb, a_ops[i], CUDA_R_16BF, lda, &bt, C, CUDA_R_32F, ldc,
                      CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
    if (st != CUBLAS_STATUS_SUCCESS) {
      cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
      mstts_set_error("gemm: cublasGemmEx bf16 (M=%d,N=%d,K=%d) failed (%d)", M, N, K, (int)st);
      return MSTTS_E_CUDA;
    }
  }
  cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
  return MSTTS_OK;
}

int gemm_rowmajor_bf16(cudaStream_t s, int M, int N, int K, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                       float* C, int ldc, float beta) {
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);  // the handle is pedantic for the fp32 path
  const float one = 1.f;
  st = cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &one, B, CUDA_R_16BF, ldb, A, CUDA_R_16BF, lda, &beta, C, CUDA_R_32F, ldc,
                    CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasGemmEx bf16 (M=%d,N=%d,K=%d) failed (%d)", M, N, K, (int)st);
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}

int gemm_bf16_ex(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B,
                 int ldb, float* C, int ldc, float beta) {
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
  const float one = 1.f;
  st = cublasGemmEx(h, transB ? CUBLAS_OP_T : CUBLAS_OP_N, transA ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &one, B, CUDA_R_16BF, ldb, A,
                    CUDA_R_16BF, lda, &beta, C, CUDA_R_32F, ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasGemmEx bf16 (M=%d,N=%d,K=%d,tA=%d,tB=%d) failed (%d)", M, N, K, (int)transA, (int)transB, (int)st);
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}

int gemm_bf16_batched(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A, int lda, long long sA,
                      const __nv_bfloat16* B, int ldb, long long sB, float* C, int ldc, long long sC, float beta, int batch) {
  int rc;
  cublasHandle_t h = get_handle(&rc);
  if (!h) return rc;
  cublasStatus_t st = cublasSetStream(h, s);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasSetStream failed (%d)", (int)st);
    return MSTTS_E_CUDA;
  }
  cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
  const float one = 1.f;
  st = cublasGemmStridedBatchedEx(h, transB ? CUBLAS_OP_T : CUBLAS_OP_N, transA ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &one, B, CUDA_R_16BF, ldb,
                                  sB, A, CUDA_R_16BF, lda, sA, &beta, C, CUDA_R_32F, ldc, sC, batch, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
  if (st != CUBLAS_STATUS_SUCCESS) {
    mstts_set_error("gemm: cublasGemmStridedBatchedEx bf16 (M=%d,N=%d,K=%d,batch=%d) failed (%d)", M, N, K, batch, (int)st);
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}
