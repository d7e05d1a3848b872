// tf.layers.conv1d(padding='same', stride 1) and its gradients on the tensor cores (bf16x3, fp32 accumulation): the encoder
// and postnet convolutions either side of the decoder (Modules.py:25-47,121-143).  fp32 cuDNN runs these
// on the SIMT pipe (22 ms of an 80 ms full-model step at config-2 shapes).
//
// Each of the three contractions is ONE product on the hand-written tcgen05 kernel (tc_gemm.h), with the k taps folded into
// the contraction dimension -- and no im2col buffer: the input is copied once into a zero-padded fp32 buffer xp [B][T + 2p][C],
// in which the im2col row of (b, t) is simply the k * C contiguous floats starting at xp[b][t][0].  The operand pack kernel
// reads a matrix through (pointer, leading dimension), so "rows of k * C values with leading dimension C" (overlapping rows)
// hands it the im2col matrix directly:
//   y  [b]   = im2col(xp[b])  . W[(tap, ci), co]                         batch = B, M = T, K = k Cin, N = Cout (kernel shared)
//   dx [b]   = im2col(dyp[b]) . Wf[(j, co), ci],  Wf = taps reversed^T     batch = B, M = T, K = k Cout, N = Cin
//   dW[(tap, ci), co] = sum over flat padded rows r: xp[r + tap][ci] dyp[r + p][co]     M = k Cin, N = Cout, K = B (T + 2p) - 2p
// (the pad rows of dyp are zero, so the flat contraction never mixes utterances).
#include "common.cuh"
#include "gemm.h"

static inline int cv_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// x [B,T,C] fp32 -> xp [B][T+2p][C] (pad rows are zeroed by the caller's memset)
__global__ void conv_pad_kernel(const float* __restrict__ x, int B, int T, int C, int p, float* __restrict__ dst) {
  const int Tp = T + 2 * p, C4 = C / 4;
  const size_t n = (size_t)B * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t bt = i / C4;
    const size_t row = (bt / T) * Tp + p + (bt % T);
    *reinterpret_cast<float4*>(dst + row * C + c) = *reinterpret_cast<const float4*>(x + bt * C + c);
  }
}

// kernel [k, Cin, Cout] -> wf [Cin][k Cout] with wf[ci][j Cout + co] = W[k-1-j][ci][co]  (the B operand of the input gradient,
// stored N x K)
__global__ void conv_flip_w_kernel(const float* __restrict__ w, int k, int Cin, int Cout, float* __restrict__ wf) {
  const size_t n = (size_t)k * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int j = (int)((i / Cout) % k);
    const int ci = (int)(i / ((size_t)Cout * k));
    wf[i] = w[((size_t)(k - 1 - j) * Cin + ci) * Cout + co];
  }
}

__global__ void conv_bias_fill_kernel(const float* __restrict__ bias, size_t rows, int C, float* __restrict__ y) {
  const size_t n = rows * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = bias[i % C];
}

static bool conv_shape_ok(int B, int T, int Cin, int Cout, int k) {
  return B >= 1 && T >= 1 && k >= 1 && (k & 1) && k <= 15 && Cin % 8 == 0 && Cout % 8 == 0 && Cin >= 8 && Cout >= 8;
}

static size_t conv_xp_bytes(int B, int T, int C, int k) { return align_up((size_t)B * (T + 2 * (k / 2)) * C * sizeof(float), 256); }

extern "C" size_t mstts_conv1d_workspace_bytes(int B, int T, int Cin, int Cout, int k) {
  if (!conv_shape_ok(B, T, Cin, Cout, k)) return 0;
  return conv_xp_bytes(B, T, Cin, k) + conv_xp_bytes(B, T, Cout, k) + align_up((size_t)k * Cin * Cout * sizeof(float), 256) + 1024;
}

extern "C" int mstts_conv1d_fwd(const float* x, const float* kernel, const float* bias, int B, int T, int Cin, int Cout, int k, float* y,
                                void* ws_, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(x && kernel && y && ws_, MSTTS_E_INVALID, "conv1d_fwd: null pointer");
  MSTTS_REQUIRE(conv_shape_ok(B, T, Cin, Cout, k), MSTTS_E_UNSUPPORTED, "conv1d: B=%d T=%d Cin=%d Cout=%d k=%d (odd k <= 15, channels %% 8)", B,
                T, Cin, Cout, k);
  MSTTS_REQUIRE(ws_bytes >= mstts_conv1d_workspace_bytes(B, T, Cin, Cout, k), MSTTS_E_WORKSPACE, "conv1d_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int p = k / 2, Tp = T + 2 * p;
  float* xp = (float*)(((uintptr_t)ws_ + 255) & ~(uintptr_t)255);
  MSTTS_CUDA(cudaMemsetAsync(xp, 0, (size_t)B * Tp * Cin * sizeof(float), s));
  conv_pad_kernel<<<cv_grid((size_t)B * T * Cin / 4), 256, 0, s>>>(x, B, T, Cin, p, xp);
  if (bias) conv_bias_fill_kernel<<<cv_grid((size_t)B * T * Cout), 256, 0, s>>>(bias, (size_t)B * T, Cout, y);
  MSTTS_CUDA(cudaGetLastError());
  return gemm_rowmajor_batched(s, false, false, T, Cout, k * Cin, xp, Cin, (long long)Tp * Cin, kernel, Cout, 0, y, Cout, (long long)T * Cout,
                               bias ? 1.f : 0.f, B);
}

extern "C" int mstts_conv1d_bwd(const float* x, const float* kernel, const float* dy, int B, int T, int Cin, int Cout, int k, float* dx,
                                float* dkernel, void* ws_, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(x && kernel && dy && dkernel && ws_, MSTTS_E_INVALID, "conv1d_bwd: null pointer");
  MSTTS_REQUIRE(conv_shape_ok(B, T, Cin, Cout, k), MSTTS_E_UNSUPPORTED, "conv1d: B=%d T=%d Cin=%d Cout=%d k=%d", B, T, Cin, Cout, k);
  MSTTS_REQUIRE(ws_bytes >= mstts_conv1d_workspace_bytes(B, T, Cin, Cout, k), MSTTS_E_WORKSPACE, "conv1d_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int p = k / 2, Tp = T + 2 * p;
  char* ws = (char*)(((uintptr_t)ws_ + 255) & ~(uintptr_t)255);
  float* xp = (float*)ws;
  float* dyp = (float*)(ws + conv_xp_bytes(B, T, Cin, k));
  float* wf = (float*)(ws + conv_xp_bytes(B, T, Cin, k) + conv_xp_bytes(B, T, Cout, k));
  MSTTS_CUDA(cudaMemsetAsync(xp, 0, (size_t)B * Tp * Cin * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(dyp, 0, (size_t)B * Tp * Cout * sizeof(float), s));
  conv_pad_kernel<<<cv_grid((size_t)B * T * Cin / 4), 256, 0, s>>>(x, B, T, Cin, p, xp);
  conv_pad_kernel<<<cv_grid((size_t)B * T * Cout / 4), 256, 0, s>>>(dy, B, T, Cout, p, dyp);
  int rc;
  if (dx) {
    // dx[b, t, ci] = sum_j sum_co dyp[b, t + j, co] W[k-1-j][ci][co]
    conv_flip_w_kernel<<<cv_grid((size_t)k * Cin * Cout), 256, 0, s>>>(kernel, k, Cin, Cout, wf);
    if ((rc = gemm_rowmajor_batched(s, false, true, T, Cin, k * Cout, dyp, Cout, (long long)Tp * Cout, wf, k * Cout, 0, dx, Cin,
                                    (long long)T * Cin, 0.f, B)))
      return rc;
  }
  MSTTS_CUDA(cudaGetLastError());
  // dW[(tap, ci), co] = sum over flat padded rows r in [0, B Tp - 2p): xp_flat[r + tap][ci] dyp_flat[r + p][co]
  const int R = B * Tp - 2 * p;
  return gemm_rowmajor_ex(s, true, false, k * Cin, Cout, R, xp, Cin, dyp + (size_t)p * Cout, Cout, dkernel, Cout, 0.f);
}
