// Hand-written tcgen05 GEMM over pre-tiled bf16 operands (tc_gemm.cu).
#pragma once
#include "common.cuh"

struct TcGemmParams {
  const uint8_t* A;   // [Mt][Kb][16 KB]
  const uint8_t* B;   // [Nt][Kb][32 KB]
  float* C;           // fp32 output (mode 0), row-major with ldc
  int M, Mt, Nt, Kb, ldc;
  // fused-epilogue operands (WaveGlow modes)
  const float *bias0, *bias1;
  float *out_f32, *skip;
  uint8_t* out_img;
  const float* g_in;
  int T, Tp, dil, first, lastl;
};

template <int MODE>
__device__ __forceinline__ void tc_gemm_epilogue(const TcGemmParams& P, uint32_t ta, int row, int nt);

int tc_gemm_plain(cudaStream_t s, const void* A_tiled, const void* B_tiled, float* C, int M, int N, int K, int ldc);
