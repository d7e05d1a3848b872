// tf.layers.conv1d(padding='same', stride 1) and its gradients on the tensor cores (bf16x3, fp32 accumulation): the encoder
// and postnet convolutions either side of the decoder (Modules.py:25-47,121-143).  fp32 cuDNN runs these
// on the SIMT pipe (22 ms of an 80 ms full-model step at config-2 shapes).
//
// The input is packed once as zero-padded stacked rows [B][T + 2p][hi | lo | hi] (bf16); tap j of the convolution is then a
// GEMM over a row-shifted view of that buffer against the stacked tap kernel [W_hi ; W_hi ; W_lo] -- the three bf16x3 partial
// products are folded into K, the k taps accumulate into the dense fp32 output through beta = 1 (strided-batched over the
// utterances so that no pad rows are computed).  Same scheme for the input gradient (d y against the transposed taps) and
// the kernel gradient (contraction over the flat padded rows, three accumulating calls per tap: the fp32 output is tiny).
#include "common.cuh"
#include "gemm.h"

static inline int cv_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// x [B,T,C] fp32 -> [B][T+2p][3C] stacked bf16 (pad rows are zeroed by the caller's memset)
__global__ void conv_pack_x_kernel(const float* __restrict__ x, int B, int T, int C, int p, __nv_bfloat16* __restrict__ dst) {
  const int Tp = T + 2 * p, C4 = C / 4;
  const size_t n = (size_t)B * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t bt = i / C4;
    const size_t row = (bt / T) * Tp + p + (bt % T);
    const float4 v = *reinterpret_cast<const float4*>(x + bt * C + c);
    const float xv[4] = {v.x, v.y, v.z, v.w};
    __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(xv[j]);
      l[j] = __float2bfloat16_rn(xv[j] - __bfloat162float(h[j]));
    }
    __nv_bfloat16* d = dst + row * 3 * C + c;
    *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(d + C) = *reinterpret_cast<const uint2*>(l);
    *reinterpret_cast<uint2*>(d + 2 * C) = *reinterpret_cast<const uint2*>(h);
  }
}

// kernel [k, Cin, Cout] fp32 -> w3 [k][3 Cin][Cout] = (hi ; hi ; lo)   and   w3t [k][3 Cout][Cin] = the same of W_tap^T
__global__ void conv_pack_w_kernel(const float* __restrict__ w, int k, int Cin, int Cout, __nv_bfloat16* __restrict__ w3,
                                   __nv_bfloat16* __restrict__ w3t) {
  const size_t n = (size_t)k * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int tap = (int)(i / ((size_t)Cin * Cout));
    const float x = w[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    if (w3) {
      __nv_bfloat16* d = w3 + ((size_t)tap * 3 * Cin + ci) * Cout + co;
      d[0] = h;
      d[(size_t)Cin * Cout] = h;
      d[(size_t)2 * Cin * Cout] = l;
    }
    if (w3t) {
      __nv_bfloat16* d = w3t + ((size_t)tap * 3 * Cout + co) * Cin + ci;
      d[0] = h;
      d[(size_t)Cout * Cin] = h;
      d[(size_t)2 * Cout * Cin] = l;
    }
  }
}

__global__ void conv_bias_fill_kernel(const float* __restrict__ bias, size_t rows, int C, float* __restrict__ y) {
  const size_t n = rows * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = bias[i % C];
}

static bool conv_shape_ok(int B, int T, int Cin, int Cout, int k) {
  return B >= 1 && T >= 1 && k >= 1 && (k & 1) && k <= 15 && Cin % 8 == 0 && Cout % 8 == 0 && Cin >= 8 && Cout >= 8;
}

extern "C" size_t mstts_conv1d_workspace_bytes(int B, int T, int Cin, int Cout, int k) {
  if (!conv_shape_ok(B, T, Cin, Cout, k)) return 0;
  const size_t Tp = T + 2 * (k / 2);
  const size_t xs = align_up((size_t)B * Tp * 3 * Cin * 2, 256), ys = align_up((size_t)B * Tp * 3 * Cout * 2, 256);
  const size_t w3 = align_up((size_t)k * 3 * Cin * Cout * 2, 256);
  return xs + ys + 2 * w3 + 1024;
}

extern "C" int mstts_conv1d_fwd(const float* x, const float* kernel, const float* bias, int B, int T, int Cin, int Cout, int k, float* y,
                                void* ws_, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(x && kernel && y && ws_, MSTTS_E_INVALID, "conv1d_fwd: null pointer");
  MSTTS_REQUIRE(conv_shape_ok(B, T, Cin, Cout, k), MSTTS_E_UNSUPPORTED, "conv1d: B=%d T=%d Cin=%d Cout=%d k=%d (odd k <= 15, channels %% 8)", B,
                T, Cin, Cout, k);
  MSTTS_REQUIRE(ws_bytes >= mstts_conv1d_workspace_bytes(B, T, Cin, Cout, k), MSTTS_E_WORKSPACE, "conv1d_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int p = k / 2, Tp = T + 2 * p;
  char* ws = (char*)(((uintptr_t)ws_ + 255) & ~(uintptr_t)255);
  __nv_bfloat16* xs = (__nv_bfloat16*)ws;
  __nv_bfloat16* w3 = (__nv_bfloat16*)(ws + align_up((size_t)B * Tp * 3 * Cin * 2, 256) + align_up((size_t)B * Tp * 3 * Cout * 2, 256));
  MSTTS_CUDA(cudaMemsetAsync(xs, 0, (size_t)B * Tp * 3 * Cin * 2, s));
  conv_pack_x_kernel<<<cv_grid((size_t)B * T * Cin / 4), 256, 0, s>>>(x, B, T, Cin, p, xs);
  conv_pack_w_kernel<<<cv_grid((size_t)k * Cin * Cout), 256, 0, s>>>(kernel, k, Cin, Cout, w3, nullptr);
  if (bias)
    conv_bias_fill_kernel<<<cv_grid((size_t)B * T * Cout), 256, 0, s>>>(bias, (size_t)B * T, Cout, y);
  int rc;
  for (int tap = 0; tap < k; ++tap) {
    // y[b, t, :] += xs[b, t + tap, :] . w3[tap]      (xs row t + tap = input time t + tap - p)
    if ((rc = gemm_bf16_batched(s, false, false, T, Cout, 3 * Cin, xs + (size_t)tap * 3 * Cin, 3 * Cin, (long long)Tp * 3 * Cin,
                                w3 + (size_t)tap * 3 * Cin * Cout, Cout, 0, y, Cout, (long long)T * Cout, (tap == 0 && !bias) ? 0.f : 1.f, B)))
      return rc;
  }
  return MSTTS_OK;
}

extern "C" int mstts_conv1d_bwd(const float* x, const float* kernel, const float* dy, int B, int T, int Cin, int Cout, int k, float* dx,
                                float* dkernel, void* ws_, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(x && kernel && dy && dkernel && ws_, MSTTS_E_INVALID, "conv1d_bwd: null pointer");
  MSTTS_REQUIRE(conv_shape_ok(B, T, Cin, Cout, k), MSTTS_E_UNSUPPORTED, "conv1d: B=%d T=%d Cin=%d Cout=%d k=%d", B, T, Cin, Cout, k);
  MSTTS_REQUIRE(ws_bytes >= mstts_conv1d_workspace_bytes(B, T, Cin, Cout, k), MSTTS_E_WORKSPACE, "conv1d_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int p = k / 2, Tp = T + 2 * p;
  char* ws = (char*)(((uintptr_t)ws_ + 255) & ~(uintptr_t)255);
  const size_t xs_b = align_up((size_t)B * Tp * 3 * Cin * 2, 256), ys_b = align_up((size_t)B * Tp * 3 * Cout * 2, 256);
  const size_t w_b = align_up((size_t)k * 3 * Cin * Cout * 2, 256);
  __nv_bfloat16* xs = (__nv_bfloat16*)ws;
  __nv_bfloat16* dys = (__nv_bfloat16*)(ws + xs_b);
  __nv_bfloat16* w3t = (__nv_bfloat16*)(ws + xs_b + ys_b + w_b);
  MSTTS_CUDA(cudaMemsetAsync(xs, 0, (size_t)B * Tp * 3 * Cin * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(dys, 0, (size_t)B * Tp * 3 * Cout * 2, s));
  conv_pack_x_kernel<<<cv_grid((size_t)B * T * Cin / 4), 256, 0, s>>>(x, B, T, Cin, p, xs);
  conv_pack_x_kernel<<<cv_grid((size_t)B * T * Cout / 4), 256, 0, s>>>(dy, B, T, Cout, p, dys);
  int rc;
  if (dx) {
    conv_pack_w_kernel<<<cv_grid((size_t)k * Cin * Cout), 256, 0, s>>>(kernel, k, Cin, Cout, nullptr, w3t);
    for (int tap = 0; tap < k; ++tap) {
      // dx[b, t, :] += dy[b, t - (tap - p), :] . W_tap^T   (dys row of time t' is t' + p  ->  row t + 2p - tap)
      if ((rc = gemm_bf16_batched(s, false, false, T, Cin, 3 * Cout, dys + (size_t)(2 * p - tap) * 3 * Cout, 3 * Cout,
                                  (long long)Tp * 3 * Cout, w3t + (size_t)tap * 3 * Cout * Cin, Cin, 0, dx, Cin, (long long)T * Cin,
                                  tap == 0 ? 0.f : 1.f, B)))
        return rc;
    }
  }
  // dW[tap] = sum over (b, t) x[b, t + tap - p, :]^T dy[b, t, :]: flat padded rows (pads are zero on both sides)
  const int R = B * Tp - 2 * p;  // rows p .. B*Tp - p of dys; x rows shifted by tap - p
  for (int tap = 0; tap < k; ++tap) {
    const __nv_bfloat16* a = xs + (size_t)tap * 3 * Cin;          // row (p + tap - p) = tap
    const __nv_bfloat16* b2 = dys + (size_t)p * 3 * Cout;
    float* dw = dkernel + (size_t)tap * Cin * Cout;
    if ((rc = gemm_bf16_ex(s, true, false, Cin, Cout, R, a, 3 * Cin, b2, 3 * Cout, dw, Cout, 0.f))) return rc;
    if ((rc = gemm_bf16_ex(s, true, false, Cin, Cout, R, a + Cin, 3 * Cin, b2, 3 * Cout, dw, Cout, 1.f))) return rc;
    if ((rc = gemm_bf16_ex(s, true, false, Cin, Cout, R, a, 3 * Cin, b2 + Cout, 3 * Cout, dw, Cout, 1.f))) return rc;
  }
  return MSTTS_OK;
}
