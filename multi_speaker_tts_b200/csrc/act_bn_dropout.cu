// activation -> batch normalisation -> dropout of the encoder / postnet conv stacks, fused (Modules.py:29-45,125-141):
//   a = relu|tanh(x);  y = ((a - mean) * rsqrt(var + eps) * gamma + beta) * mask / keep
// tf.layers.batch_normalization semantics on [B,T,C] (axis -1): in training the biased batch statistics over ALL B*T rows
// (padding included, SURVEY A-4) and moving <- moving * momentum + batch * (1 - momentum); moving statistics at inference,
// where dropout is the identity.  Two passes over the conv output instead of ~15 library elementwise / reduction launches,
// and two for the gradients instead of ~25.  Column reductions are two-stage and fixed-order (deterministic).
#include "common.cuh"

constexpr int kAbdSlices = 64;

__device__ __forceinline__ float abd_act(float x, int act) { return act == 0 ? fmaxf(x, 0.f) : tanhf(x); }

// partial sums over a slice of rows: part[slice][0][c] = sum a, part[slice][1][c] = sum a^2
__global__ void abd_stats_kernel(const float* __restrict__ x, size_t R, int C, int act, float* __restrict__ part) {
  __shared__ float s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t r0 = R * blockIdx.y / kAbdSlices, r1 = R * (blockIdx.y + 1) / kAbdSlices;
  float a1 = 0.f, a2 = 0.f;
  if (c < C)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float a = abd_act(x[r * C + c], act);
      a1 += a;
      a2 = fmaf(a, a, a2);
    }
  s1[threadIdx.y][threadIdx.x] = a1;
  s2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t1 += s1[j][threadIdx.x];
      t2 += s2[j][threadIdx.x];
    }
    part[((size_t)blockIdx.y * 2 + 0) * C + c] = t1;
    part[((size_t)blockIdx.y * 2 + 1) * C + c] = t2;
  }
}

// mean / rstd per channel (+ moving statistics update); stats[0][c] = mean, stats[1][c] = rstd
__global__ void abd_finalize_kernel(const float* __restrict__ part, size_t R, int C, float eps, float momentum, float* __restrict__ moving_mean,
                                    float* __restrict__ moving_var, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int j = 0; j < kAbdSlices; ++j) {
    t1 += (double)part[((size_t)j * 2 + 0) * C + c];
    t2 += (double)part[((size_t)j * 2 + 1) * C + c];
  }
  const double mean = t1 / (double)R;
  const double var = fmax(t2 / (double)R - mean * mean, 0.0);
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
  moving_var[c] = moving_var[c] * momentum + (float)var * (1.f - momentum);
}
__global__ void abd_inference_stats_kernel(const float* __restrict__ moving_mean, const float* __restrict__ moving_var, int C, float eps,
                                           float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  stats[c] = moving_mean[c];
  stats[C + c] = rsqrtf(moving_var[c] + eps);
}

// y = ((a - mean) rstd gamma + beta) * mask / keep ; a saved for the reverse pass (4 channels per thread)
__global__ void abd_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const uint8_t* __restrict__ mask, size_t R, int C, int act, float inv_keep,
                                 float* __restrict__ y, float* __restrict__ a_saved) {
  const int C4 = C / 4;
  const size_t n = R * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = *reinterpret_cast<const float4*>(x + i * 4);
    const float4 mu = *reinterpret_cast<const float4*>(stats + c), rs = *reinterpret_cast<const float4*>(stats + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float a[4] = {abd_act(xv.x, act), abd_act(xv.y, act), abd_act(xv.z, act), abd_act(xv.w, act)};
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w}, gv[4] = {g.x, g.y, g.z, g.w},
                bv[4] = {b.x, b.y, b.z, b.w};
    float o[4];
    uchar4 m = make_uchar4(1, 1, 1, 1);
    if (mask) m = *reinterpret_cast<const uchar4*>(mask + i * 4);
    const float mv[4] = {(float)m.x, (float)m.y, (float)m.z, (float)m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float bn = (a[j] - muv[j]) * rsv[j] * gv[j] + bv[j];
      o[j] = mask ? bn * inv_keep * mv[j] : bn;
    }
    *reinterpret_cast<float4*>(y + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
    if (a_saved) *reinterpret_cast<float4*>(a_saved + i * 4) = make_float4(a[0], a[1], a[2], a[3]);
  }
}

// reverse pass, stage 1: part[slice][0][c] = sum dyd, part[slice][1][c] = sum dyd * xhat   (dyd = dy * mask / keep)
__global__ void abd_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ stats,
                                     const uint8_t* __restrict__ mask, size_t R, int C, float inv_keep, float* __restrict__ part) {
  __shared__ float s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t r0 = R * blockIdx.y / kAbdSlices, r1 = R * (blockIdx.y + 1) / kAbdSlices;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float mu = stats[c], rs = stats[C + c];
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const size_t i = r * C + c;
      const float d = mask ? dy[i] * inv_keep * (float)mask[i] : dy[i];
      a1 += d;
      a2 = fmaf(d, (a[i] - mu) * rs, a2);
    }
  }
  s1[threadIdx.y][threadIdx.x] = a1;
  s2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t1 += s1[j][threadIdx.x];
      t2 += s2[j][threadIdx.x];
    }
    part[((size_t)blockIdx.y * 2 + 0) * C + c] = t1;
    part[((size_t)blockIdx.y * 2 + 1) * C + c] = t2;
  }
}
__global__ void abd_bwd_finalize_kernel(const float* __restrict__ part, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t1 = 0.f, t2 = 0.f;
  for (int j = 0; j < kAbdSlices; ++j) {
    t1 += part[((size_t)j * 2 + 0) * C + c];
    t2 += part[((size_t)j * 2 + 1) * C + c];
  }
  dbeta[c] = t1;
  dgamma[c] = t2;
}
// stage 2: dx = gamma rstd (dyd - dbeta / R - xhat dgamma / R) * act'(a)
__global__ void abd_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                     const uint8_t* __restrict__ mask, size_t R, int C, int act, float inv_keep, float* __restrict__ dx) {
  const int C4 = C / 4;
  const size_t n = R * C4;
  const float invR = 1.f / (float)R;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 dv = *reinterpret_cast<const float4*>(dy + i * 4), av = *reinterpret_cast<const float4*>(a + i * 4);
    uchar4 m = make_uchar4(1, 1, 1, 1);
    if (mask) m = *reinterpret_cast<const uchar4*>(mask + i * 4);
    const float d4[4] = {dv.x, dv.y, dv.z, dv.w}, a4[4] = {av.x, av.y, av.z, av.w}, mv[4] = {(float)m.x, (float)m.y, (float)m.z, (float)m.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mu = stats[c + j], rs = stats[C + c + j];
      const float d = mask ? d4[j] * inv_keep * mv[j] : d4[j];
      const float xhat = (a4[j] - mu) * rs;
      const float da = gamma[c + j] * rs * (d - dbeta[c + j] * invR - xhat * dgamma[c + j] * invR);
      o[j] = act == 0 ? (a4[j] > 0.f ? da : 0.f) : da * (1.f - a4[j] * a4[j]);
    }
    *reinterpret_cast<float4*>(dx + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static inline int abd_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

extern "C" size_t mstts_act_bn_dropout_workspace_bytes(int C) { return C > 0 ? (size_t)kAbdSlices * 2 * C * sizeof(float) + 256 : 0; }

extern "C" int mstts_act_bn_dropout_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                                        const uint8_t* mask, long long R, int C, int act, int training, float keep, float momentum, float eps,
                                        float* y, float* a_saved, float* stats, void* ws, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(x && gamma && beta && moving_mean && moving_var && y && stats && ws, MSTTS_E_INVALID, "act_bn_dropout_fwd: null pointer");
  MSTTS_REQUIRE(R >= 1 && C >= 4 && C % 4 == 0 && (act == 0 || act == 1), MSTTS_E_INVALID, "act_bn_dropout: R=%lld C=%d act=%d", R, C, act);
  MSTTS_REQUIRE(!training || (mask && keep > 0.f && keep <= 1.f), MSTTS_E_INVALID, "act_bn_dropout_fwd: training needs a mask and 0 < keep <= 1");
  MSTTS_REQUIRE(ws_bytes >= mstts_act_bn_dropout_workspace_bytes(C), MSTTS_E_WORKSPACE, "act_bn_dropout_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* part = (float*)ws;
  if (training) {
    abd_stats_kernel<<<dim3((C + 31) / 32, kAbdSlices), dim3(32, 8), 0, s>>>(x, (size_t)R, C, act, part);
    abd_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, (size_t)R, C, eps, momentum, moving_mean, moving_var, stats);
  } else {
    abd_inference_stats_kernel<<<(C + 127) / 128, 128, 0, s>>>(moving_mean, moving_var, C, eps, stats);
  }
  abd_apply_kernel<<<abd_grid((size_t)R * C / 4), 256, 0, s>>>(x, stats, gamma, beta, training ? mask : nullptr, (size_t)R, C, act,
                                                              training ? 1.f / keep : 1.f, y, training ? a_saved : nullptr);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

extern "C" int mstts_act_bn_dropout_bwd(const float* dy, const float* a_saved, const float* stats, const float* gamma, const uint8_t* mask,
                                        long long R, int C, int act, float keep, float* dx, float* dgamma, float* dbeta, void* ws,
                                        size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(dy && a_saved && stats && gamma && dx && dgamma && dbeta && ws, MSTTS_E_INVALID, "act_bn_dropout_bwd: null pointer");
  MSTTS_REQUIRE(R >= 1 && C >= 4 && C % 4 == 0 && (act == 0 || act == 1), MSTTS_E_INVALID, "act_bn_dropout: R=%lld C=%d act=%d", R, C, act);
  MSTTS_REQUIRE(ws_bytes >= mstts_act_bn_dropout_workspace_bytes(C), MSTTS_E_WORKSPACE, "act_bn_dropout_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* part = (float*)ws;
  const float inv_keep = mask ? 1.f / keep : 1.f;
  abd_bwd_stats_kernel<<<dim3((C + 31) / 32, kAbdSlices), dim3(32, 8), 0, s>>>(dy, a_saved, stats, mask, (size_t)R, C, inv_keep, part);
  abd_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, C, dgamma, dbeta);
  abd_bwd_apply_kernel<<<abd_grid((size_t)R * C / 4), 256, 0, s>>>(dy, a_saved, stats, gamma, dgamma, dbeta, mask, (size_t)R, C, act, inv_keep, dx);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
