// WaveGlow affine-coupling flows (WaveGlow/Modules.py:210-371, WaveGlow/Inv1x1.py:9-32), forward (training direction,
// with the log-likelihood sums) and reverse (synthesis direction).
//
// Numerics: bf16x3 as in the decoder kernels (operands split hi + lo, fp32 accumulation), with the three partial products
// folded into the K dimension: every activation row is stored as [hi | lo | hi] and every weight as [W_hi ; W_hi ; W_lo], so a
// product is ONE bf16 GEMM with K tripled (three K=512 calls each read-modify-wrote the 65 MB fp32 pre-activation and were
// HBM-bound: 29 us measured vs 12.5 us of tensor work).
//
// Two paths share the flow prologue / epilogue kernels:
//   * forward / inverse without saved activations (Glow_Train for evaluation, Glow_Inference): the hand-written tcgen05 GEMMs of
//     tc_gemm.cu.  Operands live in the tensor core's tile layout; the gate GEMM (K = 6528: im2col taps of the dilated conv +
//     conditioning) applies bias + tanh * sigmoid in its epilogue and writes the operand of the res/skip GEMM, whose epilogue
//     adds the residual onto the GATED activation (reference quirk B-6), writes the next layer's taps and accumulates the skip.
//     wn_apply_tiled_kernel / mel_tiled_kernel / flow_pre_tiled_kernel produce the tile images.
//   * training forward that keeps every layer's operands + the reverse pass (second half of this file): the same tcgen05 kernel
//     through its row-major front-end (tc_gemm.h: tc_pack_hl / tc_gemm_images / tc_gemm_hl) over row-major hi/lo operands in a
//     time-padded layout [N][T+2P][C] (P = 128 zero rows per utterance side, so a tap of the dilated conv is a row-shifted view),
//     the per-layer products merged so that every operand is packed once; gate_kernel / resskip_kernel and their backward
//     counterparts around them.
// Shared: wn_scale (weight norm g*v/sqrt(max(sum v^2,1e-5))), flow_post_kernel (512->c end conv + affine transform with log_s
// clamped at 8 in forward only + sum(log_s) in double, or the inverse transform followed by the inverse 1x1), mel upsampling.
#include "common.cuh"
#include "gemm.h"
#include "scratch_pool.h"
#include "tc_gemm.h"

constexpr int kWnCh = 512, kWnLayers = 8, kWnMel = 640, kWgPad = 128, kWgFlows = 12;

static inline int ew_grid(size_t n, int per = 256) {
  size_t g = (n + per - 1) / per;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// ---- weight norm + split, batched: one launch handles every weight tensor of a flow ----
// v [k*in, out] (out fastest), g [out] -> effective w = g * v / sqrt(max(sum_r v[r][o]^2, 1e-5)) (WaveGlow/Modules.py:31-33)
// written as the stacked bf16 operand [W_hi ; W_hi ; W_lo] per tap (seg rows per tap), or as fp32 (start conv).
struct WnJob {
  const float* v;
  const float* g;
  float* eff;             // fp32 destination (start conv) or NULL
  __nv_bfloat16* dst;     // stacked bf16 destination or NULL
  int kin, out, seg;      // seg = rows per tap (kin for 1x1 convs, kin/3 for the k=3 conv)
};
constexpr int kWnJobsPerFlow = 1 + 3 * 8;
struct WnJobs {
  WnJob j[kWnJobsPerFlow];
  float* scale;  // [kWnJobsPerFlow][1024]
};

// pass 1: scale[o] = g[o] / sqrt(max(sum_r v[r][o]^2, 1e-5)); block = 32 output channels x 32 row lanes, fixed order
__global__ void wn_scale_kernel(const WnJobs J) {
  __shared__ float part[32][33];
  const WnJob& job = J.j[blockIdx.y];
  const int o = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= job.out) return;
  float ss = 0.f;
  if (o < job.out)
    for (int r = threadIdx.y; r < job.kin; r += 32) {
      const float x = job.v[(size_t)r * job.out + o];
      ss = fmaf(x, x, ss);
    }
  part[threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y == 0 && o < job.out) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) tot += part[j][threadIdx.x];
    J.scale[blockIdx.y * 1024 + o] = job.g[o] * rsqrtf(fmaxf(tot, 1e-5f));
  }
}
// pass 2: w = v * scale -> fp32 (eff) or the stacked bf16 operand
__global__ void wn_apply_kernel(const WnJobs J) {
  const WnJob& job = J.j[blockIdx.y];
  const float* scale = J.scale + blockIdx.y * 1024;
  const size_t n = (size_t)job.kin * job.out;
  const size_t tap_stride = (size_t)3 * job.seg * job.out, seg_stride = (size_t)job.seg * job.out;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % job.out);
    const int r = (int)(i / job.out);
    const float w = job.v[i] * scale[o];
    if (job.eff) job.eff[i] = w;
    if (job.dst) {
      const int tap = r / job.seg, rr = r - tap * job.seg;
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      __nv_bfloat16* d = job.dst + tap * tap_stride + (size_t)rr * job.out + o;
      d[0] = h;
      d[seg_stride] = h;
      d[2 * seg_stride] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
  }
}

// activation row in the K-stacked operand layout: [hi (C) | lo (C) | hi (C)]
__device__ __forceinline__ void store_x3(__nv_bfloat16* row, int C, int c, float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  row[c] = h;
  row[C + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  row[2 * C + c] = h;
}

// ---- ConvTranspose1d 80->80, k=1024, stride 256, VALID (WaveGlow/Modules.py:198-208) ----
// out[n, p, co] = bias[co] + sum_{t: 0 <= p-256t < 1024} sum_ci mel[n,t,ci] * K[p-256t, co, ci]
// Step 1 (the hand-written GEMM, gemm.h): C[n*Tm+t][k*80+co] = sum_ci mel[n,t,ci] K[k,co,ci]   ([N*Tm,80] x [81920,80]^T)
// Step 2 (this kernel): overlap-add of the <= 4 frames that reach output position p, + bias
__global__ void upsample_overlap_add_kernel(const float* __restrict__ C, const float* __restrict__ bias, float* __restrict__ out, int N,
                                            int Tm, int Lkeep) {
  const size_t n_out = (size_t)N * Lkeep * 80;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % 80);
    const size_t np = i / 80;
    const int p = (int)(np % Lkeep), n = (int)(np / Lkeep);
    float s = bias[co];
    const int t_hi = min(Tm - 1, p / 256), t_lo = max(0, (p - 1023 + 255) / 256);
    for (int t = t_lo; t <= t_hi; ++t) s += C[((size_t)n * Tm + t) * (1024 * 80) + (size_t)(p - 256 * t) * 80 + co];
    out[i] = s;
  }
}

// mel [N,T,640] fp32 -> padded stacked operand [N][Tp][3*640]
__global__ void pad_split_kernel(const float* __restrict__ src, int N, int T, int C, __nv_bfloat16* __restrict__ dst) {
  const int Tp = T + 2 * kWgPad;
  const size_t n = (size_t)N * T * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t nt = i / C;
    const int t = (int)(nt % T), b = (int)(nt / T);
    store_x3(dst + ((size_t)b * Tp + kWgPad + t) * 3 * C, C, c, src[i]);
  }
}

// ---- flow prologue: y = x W (forward) or y = x (reverse); start conv h = y[:half] Ws + bs -> hi/lo padded ----
// x [N,T,c] -> y [N,T,c] fp32 (coupling input kept for the epilogue); Wm [c,c] row-major (y_j = sum_i x_i Wm[i][j])
__global__ void flow_pre_kernel(const float* __restrict__ x, const float* __restrict__ Wm, const float* __restrict__ Ws,
                                const float* __restrict__ bs, float* __restrict__ y, __nv_bfloat16* __restrict__ h3, int N, int T,
                                int c, int apply_w) {
  __shared__ float w_s[64], ws_s[4 * kWnCh], bs_s[kWnCh];
  const int half = c / 2, Tp = T + 2 * kWgPad;
  for (int i = threadIdx.x; i < c * c; i += blockDim.x) w_s[i] = Wm[i];
  for (int i = threadIdx.x; i < half * kWnCh; i += blockDim.x) ws_s[i] = Ws[i];
  for (int i = threadIdx.x; i < kWnCh; i += blockDim.x) bs_s[i] = bs[i];
  __syncthreads();
  const size_t rows = (size_t)N * T;
  // one warp per row: lanes 0..c-1 compute y, then all lanes sweep the 512 channels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (size_t r = (size_t)blockIdx.x * wpb + warp; r < rows; r += (size_t)gridDim.x * wpb) {
    float yv = 0.f;
    if (lane < c) {
      if (apply_w) {
        for (int i = 0; i < c; ++i) yv = fmaf(x[r * c + i], w_s[i * c + lane], yv);
      } else {
        yv = x[r * c + lane];
      }
      y[r * c + lane] = yv;
    }
    float x0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x0[j] = __shfl_sync(0xffffffffu, yv, j);
    const int b = (int)(r / T), t = (int)(r % T);
    __nv_bfloat16* hrow = h3 + ((size_t)b * Tp + kWgPad + t) * 3 * kWnCh;
    for (int ch = lane; ch < kWnCh; ch += 32) {
      float s = bs_s[ch];
      for (int j = 0; j < half; ++j) s = fmaf(x0[j], ws_s[j * kWnCh + ch], s);
      store_x3(hrow, kWnCh, ch, s);
    }
  }
}

// four consecutive channels of an activation row in the K-stacked layout [hi | lo | hi]: 8-byte stores
__device__ __forceinline__ void store_x3_vec4(__nv_bfloat16* row, int C, int c4, const float (&x)[4]) {
  __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(x[j]);
    l[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h[j]));
  }
  const uint2 hv = *reinterpret_cast<const uint2*>(h), lv = *reinterpret_cast<const uint2*>(l);
  *reinterpret_cast<uint2*>(row + c4) = hv;
  *reinterpret_cast<uint2*>(row + C + c4) = lv;
  *reinterpret_cast<uint2*>(row + 2 * C + c4) = hv;
}

// ---- gate: g = tanh(a[:, :512] + b1[:512] + b2[:512]) * sigmoid(a[:, 512:] + ...) over the valid rows; 4 channels/thread ----
__global__ void gate_kernel(const float* __restrict__ a, const float* __restrict__ b_in, const float* __restrict__ b_cond,
                            float* __restrict__ g, __nv_bfloat16* __restrict__ g3, int N, int T) {
  const int Tp = T + 2 * kWgPad;
  constexpr int C4 = kWnCh / 4;
  const size_t n = (size_t)N * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C4) * 4;
    const size_t nt = i / C4;
    const size_t row = (nt / T) * Tp + kWgPad + (nt % T);
    const float4 at = *reinterpret_cast<const float4*>(a + row * 2 * kWnCh + ch);
    const float4 as = *reinterpret_cast<const float4*>(a + row * 2 * kWnCh + kWnCh + ch);
    const float4 bt1 = *reinterpret_cast<const float4*>(b_in + ch), bt2 = *reinterpret_cast<const float4*>(b_cond + ch);
    const float4 bs1 = *reinterpret_cast<const float4*>(b_in + kWnCh + ch), bs2 = *reinterpret_cast<const float4*>(b_cond + kWnCh + ch);
    const float t4[4] = {at.x + bt1.x + bt2.x, at.y + bt1.y + bt2.y, at.z + bt1.z + bt2.z, at.w + bt1.w + bt2.w};
    const float s4[4] = {as.x + bs1.x + bs2.x, as.y + bs1.y + bs2.y, as.z + bs1.z + bs2.z, as.w + bs1.w + bs2.w};
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = tanhf(t4[j]) * (1.f / (1.f + expf(-s4[j])));
    *reinterpret_cast<float4*>(g + row * kWnCh + ch) = make_float4(v[0], v[1], v[2], v[3]);
    store_x3_vec4(g3 + row * 3 * kWnCh, kWnCh, ch, v);
  }
}

// ---- residual + skip: h = g + rs[:, :512] + b[:512] (-> stacked bf16), skip (+)= rs[:, 512:] + b[512:]; last layer: skip += rs + b ----
__global__ void resskip_kernel(const float* __restrict__ rs, const float* __restrict__ b_res, const float* __restrict__ g,
                               __nv_bfloat16* __restrict__ h3, float* __restrict__ skip, int N, int T, int first, int lastl) {
  const int Tp = T + 2 * kWgPad;
  const int ldr = lastl ? kWnCh : 2 * kWnCh;
  constexpr int C4 = kWnCh / 4;
  const size_t n = (size_t)N * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C4) * 4;
    const size_t nt = i / C4;
    const size_t row = (nt / T) * Tp + kWgPad + (nt % T);
    float4 sk;
    if (!lastl) {
      const float4 gv = *reinterpret_cast<const float4*>(g + row * kWnCh + ch);
      const float4 r0 = *reinterpret_cast<const float4*>(rs + row * ldr + ch);
      const float4 b0 = *reinterpret_cast<const float4*>(b_res + ch);
      const float hv[4] = {gv.x + (r0.x + b0.x), gv.y + (r0.y + b0.y), gv.z + (r0.z + b0.z), gv.w + (r0.w + b0.w)};
      store_x3_vec4(h3 + row * 3 * kWnCh, kWnCh, ch, hv);
      const float4 r1 = *reinterpret_cast<const float4*>(rs + row * ldr + kWnCh + ch);
      const float4 b1 = *reinterpret_cast<const float4*>(b_res + kWnCh + ch);
      sk = make_float4(r1.x + b1.x, r1.y + b1.y, r1.z + b1.z, r1.w + b1.w);
    } else {
      const float4 r0 = *reinterpret_cast<const float4*>(rs + row * ldr + ch);
      const float4 b0 = *reinterpret_cast<const float4*>(b_res + ch);
      sk = make_float4(r0.x + b0.x, r0.y + b0.y, r0.z + b0.z, r0.w + b0.w);
    }
    float4* sp = reinterpret_cast<float4*>(skip + row * kWnCh + ch);
    if (!first) {
      const float4 o = *sp;
      sk = make_float4(o.x + sk.x, o.y + sk.y, o.z + sk.z, o.w + sk.w);
    }
    *sp = sk;
  }
}

// ---- flow epilogue: o = skip We + be (512 -> c); forward: x1' = exp(min(log_s, 8)) x1 + b, sums += log_s;
//      reverse: x1 = (x1' - b) / exp(log_s), then x = [x0, x1] Winv.   One warp per row. ----
__global__ void flow_post_kernel(const float* __restrict__ skip, const float* __restrict__ We, const float* __restrict__ be,
                                 const float* __restrict__ y, const float* __restrict__ Winv, float* __restrict__ xout,
                                 double* __restrict__ partial, int N, int T, int c, int reverse) {
  __shared__ float we_s[kWnCh * 8], w_s[64];
  __shared__ double red[8];
  const int half = c / 2, Tp = T + 2 * kWgPad;
  for (int i = threadIdx.x; i < kWnCh * c; i += blockDim.x) we_s[i] = We[i];
  if (reverse)
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) w_s[i] = Winv[i];
  __syncthreads();
  const size_t rows = (size_t)N * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double local = 0.0;
  for (size_t r = (size_t)blockIdx.x * wpb + warp; r < rows; r += (size_t)gridDim.x * wpb) {
    const size_t prow = (r / T) * Tp + kWgPad + (r % T);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < kWnCh; k += 32) {
      const float sv = skip[prow * kWnCh + k];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < c) acc[j] = fmaf(sv, we_s[k * c + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    // lane j < half owns coupling channel j: log_s = o[j], b = o[half + j]
    float ls = 0.f, bb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j == lane) ls = acc[j] + be[j];
      if (j == lane + half && j < c) bb = acc[j] + be[j];
    }
    float outv = 0.f;
    if (lane < half) {
      const float x1 = y[r * c + half + lane];
      if (!reverse) {
        ls = fminf(ls, 8.0f);
        outv = expf(ls) * x1 + bb;
        local += (double)ls;
      } else {
        outv = (x1 - bb) / expf(ls);
      }
    }
    if (!reverse) {
      if (lane < half) {
        xout[r * c + lane] = y[r * c + lane];
        xout[r * c + half + lane] = outv;
      }
    } else {
      // z = [x0, x1]; x = z Winv
      float z = 0.f;
      if (lane < half) z = y[r * c + lane];
      const float x1v = __shfl_sync(0xffffffffu, outv, (lane >= half && lane < c) ? lane - half : 0);
      if (lane >= half && lane < c) z = x1v;
      float s = 0.f;
      for (int i = 0; i < c; ++i) {
        const float zi = __shfl_sync(0xffffffffu, z, i);
        if (lane < c) s = fmaf(zi, w_s[i * c + lane], s);
      }
      if (lane < c) xout[r * c + lane] = s;
    }
  }
  if (!reverse) {
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0) red[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w2 = 0; w2 < wpb; ++w2) s += red[w2];
      partial[blockIdx.x] = s;
    }
  }
}

// fixed-order final sums: out[0] += sum(partial[0..n)) ; optionally out[1] = sum z^2 handled by sumsq_kernel
__global__ void finish_sum_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    out[0] += s;
  }
}

__global__ void sumsq_kernel(const float* __restrict__ z, size_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  double local = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) local += (double)z[i] * (double)z[i];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) s += red[w2];
    partial[blockIdx.x] = s;
  }
}

// early-output bookkeeping: dst[n,t, dc0 .. dc0+nc) = src[n,t, sc0 .. sc0+nc)
__global__ void copy_channels_kernel(const float* __restrict__ src, int sc, int sc0, float* __restrict__ dst, int dc, int dc0, int nc,
                                     size_t rows) {
  const size_t n = rows * nc;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / nc;
    const int j = (int)(i % nc);
    dst[r * dc + dc0 + j] = src[r * sc + sc0 + j];
  }
}

// =====================================================================================================================
// tcgen05 path (forward / inverse without saved activations): operands live in the tensor core's tile layout (tc_gemm.h)
// =====================================================================================================================
// effective weight -> tiled B image (hi tile and lo tile per k-block).  Each thread takes 8 consecutive K rows of one output
// channel: one 16-byte chunk per tile.  perm = 1: gate permutation (n-tile j = tanh channels 128j.. | sigmoid channels 128j..)
struct WnTileJob {
  const float* v;
  const float* g_unused;
  __nv_bfloat16* dst;
  int kin, out, seg, Kb, kbase, perm;
};
struct WnTileJobs {
  WnTileJob j[24];
  const float* scale[24];
};
__global__ void wn_apply_tiled_kernel(const WnTileJobs J) {
  const WnTileJob& job = J.j[blockIdx.y];
  const float* scale = J.scale[blockIdx.y];
  const int r8n = job.kin / 8;
  const size_t n = (size_t)r8n * job.out;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % job.out);
    const int r0 = (int)(i / job.out) * 8;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
    const float sc = scale[o];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float wv = job.v[(size_t)(r0 + q) * job.out + o] * sc;
      hi[q] = __float2bfloat16_rn(wv);
      lo[q] = __float2bfloat16_rn(wv - __bfloat162float(hi[q]));
    }
    int jt, c;
    if (job.perm) {
      const int oo = o < 512 ? o : o - 512;
      jt = oo >> 7;
      c = (oo & 127) + (o < 512 ? 0 : 128);
    } else {
      jt = o >> 8;
      c = o & 255;
    }
    const int k0 = job.kbase + r0;  // K index inside the GEMM (taps are consecutive 512-row segments of the conv kernel)
    // bf16 element offset of the 8-element chunk (column c, K index k0) in the hi tile; the lo tile follows 256*64 elements later
    const size_t off = (((size_t)jt * job.Kb + (k0 >> 6)) * 2) * (256 * 64) + (size_t)(c >> 3) * 512 + (size_t)((k0 & 63) >> 3) * 64 +
                       (size_t)(c & 7) * 8;
    *reinterpret_cast<uint4*>(job.dst + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(job.dst + off + 256 * 64) = *reinterpret_cast<const uint4*>(lo);
  }
}

// conditioning [N,T,640] -> the mel K range of the im2col image (written once per call; the layers only rewrite the taps)
__global__ void mel_tiled_kernel(const float* __restrict__ src, int N, int T, uint8_t* __restrict__ img) {
  const size_t n = (size_t)N * T * (kWnMel / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % (kWnMel / 8)) * 8;
    const size_t nt = i / (kWnMel / 8);
    const int m = (int)nt;  // compact rows: m = n * T + t
    const float4 a = *reinterpret_cast<const float4*>(src + nt * kWnMel + ch), b = *reinterpret_cast<const float4*>(src + nt * kWnMel + ch + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    wn_store_hl(img, m, kWnK1 / 64, 3 * kWnCh + ch, x);
  }
}

// flow prologue for the tiled path: y = x W (or x), start conv h0 = y[:half] Ws + bs -> taps of the first layer (dilation 1)
__global__ void flow_pre_tiled_kernel(const float* __restrict__ x, const float* __restrict__ Wm, const float* __restrict__ Ws,
                                      const float* __restrict__ bs, float* __restrict__ y, uint8_t* __restrict__ img, int N, int T, int c,
                                      int apply_w) {
  __shared__ float w_s[64], ws_s[4 * kWnCh], bs_s[kWnCh];
  const int half = c / 2;
  for (int i = threadIdx.x; i < c * c; i += blockDim.x) w_s[i] = Wm[i];
  for (int i = threadIdx.x; i < half * kWnCh; i += blockDim.x) ws_s[i] = Ws[i];
  for (int i = threadIdx.x; i < kWnCh; i += blockDim.x) bs_s[i] = bs[i];
  __syncthreads();
  const size_t n = (size_t)N * T * (kWnCh / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % (kWnCh / 8)) * 8;
    const size_t r = i / (kWnCh / 8);
    const int t = (int)(r % T);
    const int m = (int)r;  // compact rows
    float xin[8], yv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xin[j] = j < c ? x[r * c + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sacc = 0.f;
      if (j < c) {
        if (apply_w) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < c) sacc = fmaf(xin[q], w_s[q * c + j], sacc);
        } else {
          sacc = xin[j];
        }
      }
      yv[j] = sacc;
    }
    if (ch == 0)
      for (int j = 0; j < c; ++j) y[r * c + j] = yv[j];
    float h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float sacc = bs_s[ch + q];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < half) sacc = fmaf(yv[j], ws_s[j * kWnCh + ch + q], sacc);
      h[q] = sacc;
    }
    wn_store_taps(img, m, t, T, 1, ch, h);
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct WgLayout {
  size_t wq;                // stacked bf16 weights, per flow: {in (3 taps x [1536,1024]) | cond [1920,1024] | res [1536,1024 or 512]} x 8
  size_t start_eff;         // [12][4*512] effective start kernels (fp32)
  size_t wscale;            // [12][25][1024] weight-norm scales
  size_t mel_up;            // [N, S, 80] = [N, T, 640]
  size_t mel3;              // padded stacked [N][Tp][3*640]
  size_t h3, g3;            // padded stacked [N][Tp][3*512]
  size_t g, skip;           // padded fp32 [N][Tp][512]
  size_t a;                 // padded fp32 [N][Tp][1024]
  size_t rs;                // padded fp32 [N][Tp][1024]: res/skip output when the pre-activations are being saved
  size_t wqt;               // tcgen05 path: tiled weight images, per flow per layer {gate B [4][34][2 x 32 KB] | res B [4|2][8][2 x 32 KB]}
  size_t a1, a2;            // tcgen05 path: im2col image [Mt][34][2 x 16 KB], gated-activation image [Mt][8][2 x 16 KB]
  size_t tcscr;             // tcgen05 path: split-K tail scratch (tc_gemm.h)
  size_t y, xa, xb;         // [N,T,8]
  size_t partial;           // doubles
  size_t total;
};
constexpr size_t kWgFlowW = (size_t)kWnLayers * (3 * kWnCh * 2 * kWnCh + kWnMel * 2 * kWnCh) + (size_t)(kWnLayers - 1) * kWnCh * 2 * kWnCh +
                            (size_t)kWnCh * kWnCh;

// bytes of tiled weight images per flow: 8 gate images (4 x 34 x 64 KB) + 7 res images (4 x 8 x 64 KB) + 1 (2 x 8 x 64 KB)
constexpr size_t kWgGateImg = (size_t)4 * (kWnK1 / 64) * 65536, kWgResImg = (size_t)4 * (kWnK2 / 64) * 65536;
constexpr size_t kWgFlowWt = 8 * kWgGateImg + 7 * kWgResImg + kWgResImg / 2;

static WgLayout wg_layout(int N, int T) {
  WgLayout l;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  const size_t rows_p = (size_t)N * (T + 2 * kWgPad);
  l.wq = take(kWgFlowW * kWgFlows * 3 * 2);
  l.start_eff = take((size_t)kWgFlows * 4 * kWnCh * 4);
  l.wscale = take((size_t)kWgFlows * kWnJobsPerFlow * 1024 * 4);
  l.mel_up = take((size_t)N * T * kWnMel * 4);
  l.mel3 = take(rows_p * 3 * kWnMel * 2);
  l.h3 = take(rows_p * 3 * kWnCh * 2);
  l.g3 = take(rows_p * 3 * kWnCh * 2);
  l.g = take(rows_p * kWnCh * 4);
  l.skip = take(rows_p * kWnCh * 4);
  l.a = take(rows_p * 2 * kWnCh * 4);
  l.rs = take(rows_p * 2 * kWnCh * 4);
  {
    const size_t Mt = ((size_t)N * T + 127) / 128;  // the tiled path has no pad rows: the im2col taps carry the zero borders
    l.wqt = take(kWgFlowWt * kWgFlows + 1024);
    l.a1 = take(Mt * (kWnK1 / 64) * 32768 + 1024);
    l.a2 = take(Mt * (kWnK2 / 64) * 32768 + 1024);
    l.tcscr = take(kTcGemmScratchBytes);
  }
  l.y = take((size_t)N * T * 8 * 4);
  l.xa = take((size_t)N * T * 8 * 4);
  l.xb = take((size_t)N * T * 8 * 4);
  l.partial = take(4096 * 8);
  l.total = off;
  return l;
}

extern "C" size_t mstts_waveglow_workspace_bytes(int N, int T) {
  if (N <= 0 || T <= 0) return 0;
  return wg_layout(N, T).total;
}

// Saved activations of a training-direction forward (what the reverse pass reads).  180 GB of HBM makes keeping every
// layer's operands (163 MB per layer, 15.7 GB at config 3) cheaper than recomputing the WN stack through the inverse flow.
struct WgSave {
  size_t xin, y;       // per flow [rows, 8] fp32: flow input (after the early split) / after the invertible 1x1
  size_t skip;         // per flow [rows_p, 512] fp32 (padded layout)
  size_t h3, g3;       // per (flow, layer) stacked bf16 [rows_p, 1536]: layer input / gated activation
  size_t a;            // per (flow, layer) fp32 [rows_p, 1024]: pre-activations without the biases
  size_t total;
};
static WgSave wg_save_layout(int N, int T, size_t base) {
  WgSave v;
  size_t off = base;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  const size_t rows = (size_t)N * T, rows_p = (size_t)N * (T + 2 * kWgPad);
  v.xin = take(kWgFlows * rows * 8 * 4);
  v.y = take(kWgFlows * rows * 8 * 4);
  v.skip = take(kWgFlows * rows_p * kWnCh * 4);
  v.h3 = take((size_t)kWgFlows * kWnLayers * rows_p * 3 * kWnCh * 2);
  v.g3 = take((size_t)kWgFlows * kWnLayers * rows_p * 3 * kWnCh * 2);
  v.a = take((size_t)kWgFlows * kWnLayers * rows_p * 2 * kWnCh * 4);
  v.total = off;
  return v;
}

static inline int flow_c(int f) { return 8 - 2 * (f / 4); }

// Products over the K-stacked operands of the saved-activation path, on the hand-written kernel (the pack step reads the hi
// and lo blocks of the stacked rows; the third block is a leftover of the folded-K library form and is ignored):
//   activation rows  [hi (K) | lo (K) | hi (K)]            leading dimension 3K
//   row-stacked W    [W_hi (K rows) ; W_hi ; W_lo] x N     (forward products)
//   column-stacked W rows [hi (N) | hi (N) | lo (N)]       (transposed products, stored N_out x 3K)
static int gemm_stacked(cudaStream_t s, int M, int N, int K, const __nv_bfloat16* A3, const __nv_bfloat16* B3, float* C, int ldc, float beta) {
  return tc_gemm_hl(s, false, false, M, N, K, A3, A3 + K, 3 * K, 0, B3, B3 + (size_t)2 * K * N, N, 0, C, ldc, 0, beta, 1);
}
// C[M, N] (+)= A . W^T with A activation-stacked (width K) and W column-stacked [N][3K]
static int gemm_stacked_nt(cudaStream_t s, int M, int N, int K, const __nv_bfloat16* A3, const __nv_bfloat16* Wc, float* C, int ldc, float beta) {
  return tc_gemm_hl(s, false, true, M, N, K, A3, A3 + K, 3 * K, 0, Wc, Wc + 2 * K, 3 * K, 0, C, ldc, 0, beta, 1);
}

// Runs all 12 flows.  direction 0: training direction x -> z with sums[0] = sum(log_s), sums[1] = sum(z^2);
// direction 1: synthesis z -> x (early_noise[2]: the two [N,T,2] noise tensors injected before flows 7 and 3 are run,
// i.e. after undoing flows 8 and 4; inv_w then holds the INVERSE 1x1 kernels).
// zero the pad rows (kWgPad leading / trailing rows per utterance) of `nslots` consecutive padded buffers [N][T+2P][row16 x 16 B]
__global__ void zero_pad_rows_kernel(uint4* __restrict__ base, int nslots, size_t row16, int N, int T) {
  const int Tp = T + 2 * kWgPad;
  const size_t per_utt = (size_t)2 * kWgPad * row16, per_slot = (size_t)N * per_utt, n = (size_t)nslots * per_slot;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t slot = i / per_slot, r = i % per_slot;
    const size_t u = r / per_utt, q = r % per_utt;
    const size_t prow = q / row16, col = q % row16;
    const size_t row = prow < (size_t)kWgPad ? prow : (size_t)T + prow;  // second half: rows T+P .. T+2P-1
    base[((slot * N + u) * Tp + row) * row16 + col] = make_uint4(0u, 0u, 0u, 0u);
  }
}

static bool g_waveglow_force_library = false;  // A/B switch for tools/bench_secondary.py (mstts_waveglow_set_path)
extern "C" int mstts_waveglow_set_path(int library_gemm) {
  g_waveglow_force_library = library_gemm != 0;
  return MSTTS_OK;
}

static int waveglow_flows_impl(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T,
                               int direction, const float* const* early_noise, float* out, double* sums, void* ws_, size_t ws_bytes,
                               void* stream_, const WgSave* save) {
  MSTTS_REQUIRE(w && audio_in && mel_nt640 && out && ws_, MSTTS_E_INVALID, "waveglow: null argument");
  MSTTS_REQUIRE(N >= 1 && T >= 1, MSTTS_E_INVALID, "waveglow: N=%d T=%d", N, T);
  MSTTS_REQUIRE(direction == 0 || (early_noise && early_noise[0] && early_noise[1]), MSTTS_E_INVALID, "waveglow: reverse needs early noise");
  MSTTS_REQUIRE(direction == 1 || sums, MSTTS_E_INVALID, "waveglow: forward needs the sums output");
  const WgLayout l = wg_layout(N, T);
  MSTTS_REQUIRE(ws_bytes >= l.total, MSTTS_E_WORKSPACE, "waveglow: workspace %zu < %zu", ws_bytes, l.total);
  cudaStream_t s = (cudaStream_t)stream_;
  char* ws = (char*)ws_;
  const int Tp = T + 2 * kWgPad;
  const size_t rows = (size_t)N * T, rows_p = (size_t)N * Tp;
  auto BF = [&](size_t off) { return (__nv_bfloat16*)(ws + off); };
  auto FP = [&](size_t off) { return (float*)(ws + off); };
  int rc;

  // The hand-written tcgen05 GEMMs with fused gate / residual-skip epilogues run the forward and inverse directions; the
  // training forward that keeps every layer's operands for the reverse pass stays on the stacked row-major path.
  const bool use_tc = save == nullptr && !g_waveglow_force_library;
  // ---- effective weights (weight norm is part of the per-step graph in the reference, Modules.py:31-33) ----
  for (int f = 0; f < kWgFlows; ++f) {
    WnJobs J;
    memset(&J, 0, sizeof(J));
    J.scale = FP(l.wscale) + (size_t)f * kWnJobsPerFlow * 1024;
    __nv_bfloat16* wq = BF(l.wq) + (size_t)f * kWgFlowW * 3;
    const int half = flow_c(f) / 2;
    int nj = 0;
    J.j[nj++] = WnJob{w->start_v[f], w->start_g[f], FP(l.start_eff) + (size_t)f * 4 * kWnCh, nullptr, half, kWnCh, half};
    size_t wo = 0;
    for (int i = 0; i < kWnLayers; ++i) {
      J.j[nj++] = WnJob{w->in_v[f][i], w->in_g[f][i], nullptr, wq + wo, 3 * kWnCh, 2 * kWnCh, kWnCh};
      wo += (size_t)9 * kWnCh * 2 * kWnCh;
      J.j[nj++] = WnJob{w->cond_v[f][i], w->cond_g[f][i], nullptr, wq + wo, kWnMel, 2 * kWnCh, kWnMel};
      wo += (size_t)3 * kWnMel * 2 * kWnCh;
      const int rout = i < kWnLayers - 1 ? 2 * kWnCh : kWnCh;
      J.j[nj++] = WnJob{w->res_v[f][i], w->res_g[f][i], nullptr, wq + wo, kWnCh, rout, kWnCh};
      wo += (size_t)3 * kWnCh * rout;
    }
    wn_scale_kernel<<<dim3(32, kWnJobsPerFlow), dim3(32, 32), 0, s>>>(J);
    if (!use_tc) {
      wn_apply_kernel<<<dim3(64, kWnJobsPerFlow), 256, 0, s>>>(J);
    } else {
      wn_apply_kernel<<<dim3(64, 1), 256, 0, s>>>(J);  // job 0: the start conv's fp32 effective kernel
      WnTileJobs TJ;
      memset(&TJ, 0, sizeof(TJ));
      __nv_bfloat16* img = (__nv_bfloat16*)(ws + l.wqt + (size_t)f * kWgFlowWt);
      size_t io = 0;  // byte offset inside this flow's images
      for (int i = 0; i < kWnLayers; ++i) {
        __nv_bfloat16* gate_img = (__nv_bfloat16*)((char*)img + io);
        io += kWgGateImg;
        __nv_bfloat16* res_img = (__nv_bfloat16*)((char*)img + io);
        const int rout = i < kWnLayers - 1 ? 2 * kWnCh : kWnCh;
        io += i < kWnLayers - 1 ? kWgResImg : kWgResImg / 2;
        TJ.j[3 * i + 0] = WnTileJob{w->in_v[f][i], nullptr, gate_img, 3 * kWnCh, 2 * kWnCh, kWnCh, kWnK1 / 64, 0, 1};
        TJ.j[3 * i + 1] = WnTileJob{w->cond_v[f][i], nullptr, gate_img, kWnMel, 2 * kWnCh, kWnMel, kWnK1 / 64, 3 * kWnCh, 1};
        TJ.j[3 * i + 2] = WnTileJob{w->res_v[f][i], nullptr, res_img, kWnCh, rout, kWnCh, kWnK2 / 64, 0, 0};
        for (int q = 0; q < 3; ++q) TJ.scale[3 * i + q] = J.scale + (size_t)(1 + 3 * i + q) * 1024;
      }
      wn_apply_tiled_kernel<<<dim3(48, 24), 256, 0, s>>>(TJ);
    }
  }
  // ---- conditioning operand + zeroed pads ----
  MSTTS_CUDA(cudaMemsetAsync(ws + l.mel3, 0, rows_p * 3 * kWnMel * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.h3, 0, rows_p * 3 * kWnCh * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.g3, 0, rows_p * 3 * kWnCh * 2, s));
  if (save) {
    // saved skip / h3 / g3 slots: the pad rows must read 0 (shifted conv views, flat-row weight-gradient contractions); the
    // valid rows are fully overwritten by the layers, so only the 2 x 128 pad rows per utterance are cleared (11 % of 11 GB)
    zero_pad_rows_kernel<<<148 * 8, 256, 0, s>>>((uint4*)(ws + save->skip), kWgFlows, (size_t)kWnCh * 4 / 16, N, T);
    zero_pad_rows_kernel<<<148 * 8, 256, 0, s>>>((uint4*)(ws + save->h3), 2 * kWgFlows * kWnLayers, (size_t)3 * kWnCh * 2 / 16, N, T);
  }
  const size_t slot3 = rows_p * 3 * kWnCh * 2, slota = rows_p * 2 * kWnCh * 4;
  auto H3 = [&](int f, int i) { return save ? BF(save->h3 + ((size_t)f * kWnLayers + i) * slot3) : BF(l.h3); };
  auto G3 = [&](int f, int i) { return save ? BF(save->g3 + ((size_t)f * kWnLayers + i) * slot3) : BF(l.g3); };
  auto APRE = [&](int f, int i) { return save ? FP(save->a + ((size_t)f * kWnLayers + i) * slota) : FP(l.a); };
  auto SKIP = [&](int f) { return save ? FP(save->skip + (size_t)f * rows_p * kWnCh * 4) : FP(l.skip); };
  auto YBUF = [&](int f) { return save ? FP(save->y + (size_t)f * rows * 8 * 4) : FP(l.y); };
  if (use_tc) {
    const size_t Mt = (rows + 127) / 128;
    MSTTS_CUDA(cudaMemsetAsync(ws + l.a1, 0, Mt * (kWnK1 / 64) * 32768, s));
    MSTTS_CUDA(cudaMemsetAsync(ws + l.a2, 0, Mt * (kWnK2 / 64) * 32768, s));
    mel_tiled_kernel<<<ew_grid((size_t)rows * kWnMel / 8), 256, 0, s>>>(mel_nt640, N, T, (uint8_t*)(ws + l.a1));
  } else {
    pad_split_kernel<<<ew_grid((size_t)rows * kWnMel), 256, 0, s>>>(mel_nt640, N, T, kWnMel, BF(l.mel3));
  }
  if (direction == 0) MSTTS_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s));

  const int M = (int)(rows_p - 2 * kWgPad);  // GEMM rows: everything except the outermost pads
  ScratchScope wg_scope(s);
  void *a1img = nullptr, *b1img = nullptr;
  if (!use_tc) {
    if ((rc = wg_scope.get(&a1img, tc_image_bytes(M, kWnK1, 128)))) return rc;
    if ((rc = wg_scope.get(&b1img, tc_image_bytes(2 * kWnCh, kWnK1, 256)))) return rc;
    const __nv_bfloat16* m3 = BF(l.mel3) + (size_t)kWgPad * 3 * kWnMel;
    if ((rc = tc_pack_hl(s, m3, m3 + kWnMel, 3 * kWnMel, false, M, kWnMel, 128, kWnK1 / 64, a1img, 0, 24))) return rc;
  }
  const float* xcur = audio_in;              // [N,T,c] of the current flow
  float* xbuf[2] = {FP(l.xa), FP(l.xb)};
  int xsel = 0;
  int zc0 = 0;  // forward: next output channel of z
  const int nblk_post = 148 * 2;
  for (int step = 0; step < kWgFlows; ++step) {
    const int f = direction == 0 ? step : kWgFlows - 1 - step;
    const int c = flow_c(f);
    if (direction == 0 && f % 4 == 0 && f > 0) {
      // early output: first 2 channels leave the chain (Modules.py:334-336)
      copy_channels_kernel<<<ew_grid(rows * 2), 256, 0, s>>>(xcur, c + 2, 0, out, 8, zc0, 2, rows);
      zc0 += 2;
      float* nx = xbuf[xsel ^ 1];
      copy_channels_kernel<<<ew_grid(rows * c), 256, 0, s>>>(xcur, c + 2, 2, nx, c, 0, c, rows);
      xcur = nx;
      xsel ^= 1;
    }
    if (save) MSTTS_CUDA(cudaMemcpyAsync(ws + save->xin + (size_t)f * rows * 8 * 4, xcur, rows * c * 4, cudaMemcpyDeviceToDevice, s));
    if (use_tc)
      flow_pre_tiled_kernel<<<148 * 4, 256, 0, s>>>(xcur, w->inv_w[f], FP(l.start_eff) + (size_t)f * 4 * kWnCh, w->start_b[f], YBUF(f),
                                                    (uint8_t*)(ws + l.a1), N, T, c, direction == 0 ? 1 : 0);
    else
      flow_pre_kernel<<<148 * 4, 256, 0, s>>>(xcur, w->inv_w[f], FP(l.start_eff) + (size_t)f * 4 * kWnCh, w->start_b[f], YBUF(f), H3(f, 0), N,
                                              T, c, direction == 0 ? 1 : 0);
    const __nv_bfloat16* wq = BF(l.wq) + (size_t)f * kWgFlowW * 3;
    size_t wo = 0;
    if (use_tc) {
      const char* img = ws + l.wqt + (size_t)f * kWgFlowWt;
      size_t io = 0;
      for (int i = 0; i < kWnLayers; ++i) {
        const char* gate_img = img + io;
        io += kWgGateImg;
        const char* res_img = img + io;
        const bool lastl = i == kWnLayers - 1;
        io += lastl ? kWgResImg / 2 : kWgResImg;
        if ((rc = tc_gemm_wn_gate(s, ws + l.a1, gate_img, (int)rows, w->in_b[f][i], w->cond_b[f][i], FP(l.g), ws + l.a2, T, Tp, ws + l.tcscr))) return rc;
        if ((rc = tc_gemm_wn_res(s, ws + l.a2, res_img, (int)rows, w->res_b[f][i], FP(l.g), SKIP(f), ws + l.a1, T, Tp, 2 << i, i == 0, lastl ? 1 : 0, ws + l.tcscr)))
          return rc;
      }
    }
    for (int i = 0; i < (use_tc ? 0 : kWnLayers); ++i) {
      const int d = 1 << i;
      const int K3 = 3 * kWnCh;
      float* a_out = APRE(f, i) + (size_t)kWgPad * 2 * kWnCh;
      const __nv_bfloat16* Wtap = wq + wo;
      const __nv_bfloat16* Wc = wq + wo + (size_t)3 * K3 * 2 * kWnCh;
      // ONE product per gate pre-activation: the operand image [3 taps x 512 | mel 640] (K = 2176) is assembled from the three
      // row-shifted views of the layer input; its conditioning k-blocks were packed once before the flow loop and stay in place
      for (int k = 0; k < 3; ++k) {
        const long long shift = (long long)(kWgPad + (k - 1) * d) * K3;
        const __nv_bfloat16* hk = H3(f, i) + shift;
        if ((rc = tc_pack_hl(s, hk, hk + kWnCh, K3, false, M, kWnCh, 128, kWnK1 / 64, a1img, 0, 8 * k))) return rc;
        const __nv_bfloat16* wk = Wtap + (size_t)k * K3 * 2 * kWnCh;
        if ((rc = tc_pack_hl(s, wk, wk + (size_t)2 * kWnCh * 2 * kWnCh, 2 * kWnCh, true, 2 * kWnCh, kWnCh, 256, kWnK1 / 64, b1img, 0, 8 * k))) return rc;
      }
      if ((rc = tc_pack_hl(s, Wc, Wc + (size_t)2 * kWnMel * 2 * kWnCh, 2 * kWnCh, true, 2 * kWnCh, kWnMel, 256, kWnK1 / 64, b1img, 0, 24))) return rc;
      if ((rc = tc_gemm_images(s, a1img, b1img, M, 2 * kWnCh, kWnK1, a_out, 2 * kWnCh, 0.f))) return rc;
      wo += (size_t)3 * K3 * 2 * kWnCh + (size_t)3 * kWnMel * 2 * kWnCh;
      gate_kernel<<<ew_grid(rows * kWnCh / 4), 256, 0, s>>>(APRE(f, i), w->in_b[f][i], w->cond_b[f][i], FP(l.g), G3(f, i), N, T);
      const int rout = i < kWnLayers - 1 ? 2 * kWnCh : kWnCh;
      // res/skip output: the scratch pre-activation buffer when the pre-activations are saved elsewhere, else in place
      float* rsbuf = save ? FP(l.rs) : FP(l.a);
      if ((rc = gemm_stacked(s, M, rout, kWnCh, G3(f, i) + (size_t)kWgPad * K3, wq + wo, rsbuf + (size_t)kWgPad * rout, rout, 0.f))) return rc;
      wo += (size_t)K3 * rout;
      resskip_kernel<<<ew_grid(rows * kWnCh / 4), 256, 0, s>>>(rsbuf, w->res_b[f][i], FP(l.g), H3(f, i < kWnLayers - 1 ? i + 1 : i), SKIP(f), N,
                                                           T, i == 0, i == kWnLayers - 1);
    }
    float* xnext = xbuf[xsel ^ 1];
    flow_post_kernel<<<nblk_post, 256, 0, s>>>(SKIP(f), w->end_w[f], w->end_b[f], YBUF(f), w->inv_w[f], xnext, (double*)(ws + l.partial),
                                               N, T, c, direction);
    if (direction == 0) finish_sum_kernel<<<1, 32, 0, s>>>((double*)(ws + l.partial), nblk_post, sums);
    xcur = xnext;
    xsel ^= 1;
    if (direction == 1 && f % 4 == 0 && f > 0) {
      // prepend 2 fresh noise channels (Modules.py:363-369): early_noise[0] after flow 8, [1] after flow 4
      float* nx = xbuf[xsel ^ 1];
      copy_channels_kernel<<<ew_grid(rows * 2), 256, 0, s>>>(early_noise[f == 8 ? 0 : 1], 2, 0, nx, c + 2, 0, 2, rows);
      copy_channels_kernel<<<ew_grid(rows * c), 256, 0, s>>>(xcur, c, 0, nx, c + 2, 2, c, rows);
      xcur = nx;
      xsel ^= 1;
    }
  }
  if (direction == 0) {
    copy_channels_kernel<<<ew_grid(rows * 4), 256, 0, s>>>(xcur, 4, 0, out, 8, zc0, 4, rows);
    sumsq_kernel<<<256, 256, 0, s>>>(out, rows * 8, (double*)(ws + l.partial));
    finish_sum_kernel<<<1, 32, 0, s>>>((double*)(ws + l.partial), 256, sums + 1);
  } else {
    MSTTS_CUDA(cudaMemcpyAsync(out, xcur, rows * 8 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

extern "C" int mstts_waveglow_flows(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T,
                                    int direction, const float* const* early_noise, float* out, double* sums, void* ws, size_t ws_bytes,
                                    void* stream) {
  return waveglow_flows_impl(w, audio_in, mel_nt640, N, T, direction, early_noise, out, sums, ws, ws_bytes, stream, nullptr);
}

extern "C" size_t mstts_upsample_mel_workspace_bytes(int N, int Tm) {
  if (N <= 0 || Tm <= 0) return 0;
  return (size_t)N * Tm * 1024 * 80 * sizeof(float);
}

extern "C" int mstts_upsample_mel(const float* mel, const float* kernel, const float* bias, int N, int Tm, int keep, float* out, void* ws,
                                  size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(mel && kernel && bias && out && ws, MSTTS_E_INVALID, "upsample_mel: null pointer");
  const int Lout = (Tm - 1) * 256 + 1024;
  MSTTS_REQUIRE(keep >= 1 && keep <= Lout, MSTTS_E_INVALID, "upsample_mel: keep=%d outside [1,%d] (the reference's tf.slice would fail)", keep,
                Lout);
  MSTTS_REQUIRE(ws_bytes >= mstts_upsample_mel_workspace_bytes(N, Tm), MSTTS_E_WORKSPACE, "upsample_mel: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = gemm_rowmajor_ex(s, false, true, N * Tm, 1024 * 80, 80, mel, 80, kernel, 80, (float*)ws, 1024 * 80, 0.f);
  if (rc) return rc;
  upsample_overlap_add_kernel<<<ew_grid((size_t)N * keep * 80), 256, 0, s>>>((const float*)ws, bias, out, N, Tm, keep);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

// =====================================================================================================================
// Training: forward with saved activations + reverse pass (what tf.gradients builds for WaveGlow/WaveGlow.py:48-70).
//   L = -sum(log_s)/n - sum(logdet_W)/n + sum(z^2)/(2 sigma^2 n),  n = N*T*8  (WaveGlow/Modules.py:373-384)
// The logdet term only touches the twelve c x c kernels and stays host math, like in the forward direction.
// Every dense product is again a bf16 GEMM with the bf16x3 terms folded into K (activation gradients [hi|lo|hi] against
// column-stacked weights [W_hi | W_hi | W_lo]); weight gradients contract over the 16 000 positions with a small fp32
// output, so their three partial products are three accumulating GEMM calls.
// =====================================================================================================================
struct WgBwdLayout {
  size_t wqc;                  // column-stacked bf16 weights, per flow per layer: in taps 3 x [512, 3072] | cond [640, 3072] | res [512, 3*rout]
  size_t dh[2];                // fp32 [rows_p, 512] gradient w.r.t. a layer input (ping-pong)
  size_t dskip;                // fp32 [rows_p, 512]
  size_t drs3;                 // stacked bf16 [rows_p, 3*1024]
  size_t drs3_last;            // stacked bf16 [rows_p, 3*512] (last layer: res conv has 512 outputs)
  size_t dg;                   // fp32 [rows_p, 512]
  size_t da;                   // fp32 [rows_p, 1024]
  size_t da3;                  // stacked bf16 [rows_p, 3*1024]
  size_t dmel;                 // fp32 [rows_p, 640] accumulated over all layers and flows
  size_t dopad;                // fp32 [rows_p, 8] gradient w.r.t. the end-conv output (padded layout)
  size_t dy, dx, dx2;          // [rows, 8] fp32
  size_t dw_eff;               // fp32 [3*512*1024] gradient w.r.t. one effective weight tensor
  size_t part;                 // fp32 partial sums (colsum / small reductions)
  size_t total;
};
constexpr size_t kWgFlowWc = kWgFlowW;  // same element count per flow as the row-stacked image (x3 copies)
constexpr int kWgPartSlices = 64;

static WgBwdLayout wg_bwd_layout(int N, int T, size_t base) {
  WgBwdLayout b;
  size_t off = base;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  const size_t rows = (size_t)N * T, rows_p = (size_t)N * (T + 2 * kWgPad);
  b.wqc = take(kWgFlowWc * kWgFlows * 3 * 2);
  b.dh[0] = take(rows_p * kWnCh * 4);
  b.dh[1] = take(rows_p * kWnCh * 4);
  b.dskip = take(rows_p * kWnCh * 4);
  b.drs3 = take(rows_p * 3 * 2 * kWnCh * 2);
  b.drs3_last = take(rows_p * 3 * kWnCh * 2);
  b.dg = take(rows_p * kWnCh * 4);
  b.da = take(rows_p * 2 * kWnCh * 4);
  b.da3 = take(rows_p * 3 * 2 * kWnCh * 2);
  b.dmel = take(rows_p * kWnMel * 4);
  b.dopad = take(rows_p * 8 * 4);
  b.dy = take(rows * 8 * 4);
  b.dx = take(rows * 8 * 4);
  b.dx2 = take(rows * 8 * 4);
  b.dw_eff = take((size_t)3 * kWnCh * 2 * kWnCh * 4);
  b.part = take((size_t)kWgPartSlices * 4 * 2 * kWnCh * 4 + 148 * 8 * 4 * kWnCh * 4);
  b.total = off;
  return b;
}

extern "C" size_t mstts_waveglow_train_workspace_bytes(int N, int T) {
  if (N <= 0 || T <= 0) return 0;
  const WgSave v = wg_save_layout(N, T, wg_layout(N, T).total);
  return wg_bwd_layout(N, T, v.total).total;
}

extern "C" int mstts_waveglow_train_fwd(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T, float* z,
                                        double* sums, void* ws, size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(N >= 1 && T >= 1, MSTTS_E_INVALID, "waveglow: N=%d T=%d", N, T);
  MSTTS_REQUIRE(ws_bytes >= mstts_waveglow_train_workspace_bytes(N, T), MSTTS_E_WORKSPACE, "waveglow train: workspace %zu < %zu", ws_bytes,
                mstts_waveglow_train_workspace_bytes(N, T));
  const WgSave v = wg_save_layout(N, T, wg_layout(N, T).total);
  return waveglow_flows_impl(w, audio_in, mel_nt640, N, T, 0, nullptr, z, sums, ws, ws_bytes, stream, &v);
}

// ---- column-stacked weights for the transposed products: dst[r][s*out + o], s = 0,1 -> hi, s = 2 -> lo (per tap) ----
__global__ void wn_apply_cols_kernel(const WnJobs J) {
  const WnJob& job = J.j[blockIdx.y];
  if (!job.dst) return;
  const float* scale = J.scale + blockIdx.y * 1024;
  const size_t n = (size_t)job.kin * job.out;
  const size_t tap_stride = (size_t)3 * job.seg * job.out;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % job.out);
    const int r = (int)(i / job.out);
    const float wv = job.v[i] * scale[o];
    const int tap = r / job.seg, rr = r - tap * job.seg;
    const __nv_bfloat16 h = __float2bfloat16_rn(wv);
    __nv_bfloat16* d = job.dst + tap * tap_stride + (size_t)rr * 3 * job.out + o;
    d[0] = h;
    d[job.out] = h;
    d[2 * job.out] = __float2bfloat16_rn(wv - __bfloat162float(h));
  }
}

// weight-norm backward for one tensor: w = g v / sqrt(max(ss, 1e-5)), ss = sum_r v^2 per output channel
//   dg = sum_r dw v / sqrt(.) ;  dv = g dw / sqrt(.) - [ss > 1e-5] g v (sum_r dw v) / sqrt(.)^3
__global__ void wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ dw, int kin, int out,
                              float* __restrict__ dg, float* __restrict__ dv) {
  __shared__ float p_ss[32][33], p_dot[32][33];
  const int o = blockIdx.x * 32 + threadIdx.x;
  float ss = 0.f, dot = 0.f;
  if (o < out)
    for (int r = threadIdx.y; r < kin; r += 32) {
      const float x = v[(size_t)r * out + o];
      ss = fmaf(x, x, ss);
      dot = fmaf(dw[(size_t)r * out + o], x, dot);
    }
  p_ss[threadIdx.y][threadIdx.x] = ss;
  p_dot[threadIdx.y][threadIdx.x] = dot;
  __syncthreads();
  ss = 0.f;
  dot = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    ss += p_ss[j][threadIdx.x];
    dot += p_dot[j][threadIdx.x];
  }
  if (o >= out) return;
  const float inv = rsqrtf(fmaxf(ss, 1e-5f));
  const float gv = g[o];
  if (threadIdx.y == 0) dg[o] = dot * inv;
  const float c2 = ss > 1e-5f ? gv * dot * inv * inv * inv : 0.f;
  for (int r = threadIdx.y; r < kin; r += 32) {
    const size_t i = (size_t)r * out + o;
    dv[i] = gv * inv * dw[i] - c2 * v[i];
  }
}

// column sums over the VALID rows of a padded-layout fp32 matrix [N][Tp][C] (ld = C): deterministic two passes
__global__ void colsum_valid_partial_kernel(const float* __restrict__ in, int ld, float* __restrict__ part, int N, int T, int C) {
  __shared__ float sh[8][33];
  const int Tp = T + 2 * kWgPad;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t R = (size_t)N * T;
  const size_t r0 = R * blockIdx.y / kWgPartSlices, r1 = R * (blockIdx.y + 1) / kWgPartSlices;
  float s = 0.f;
  if (c < C)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) s += in[((r / T) * Tp + kWgPad + (r % T)) * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += sh[j][threadIdx.x];
    part[(size_t)blockIdx.y * C + c] = tot;
  }
}
__global__ void sum_partials_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float tot = 0.f;
  for (int j = 0; j < nparts; ++j) tot += part[(size_t)j * C + c];
  out[c] = accumulate ? out[c] + tot : tot;
}

// dz = z / (sigma^2 n): rows [rows, 8] -> the three channel groups of the flow chain
__global__ void dz_kernel(const float* __restrict__ z, float scale, size_t n, float* __restrict__ dz) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dz[i] = z[i] * scale;
}

// coupling + end conv backward, one warp per row.
//   recompute o = skip We + be -> (log_s_raw, b); ls = min(log_s_raw, 8); x1' = exp(ls) x1 + b
//   in : dxo [rows, c] gradient w.r.t. the flow output [x0, x1'], y = [x0, x1] (after the 1x1)
//   out: dy [rows, c] = [dxo_x0 (the WN path is added later), dx1' exp(ls)], do_pad [rows_p, 8] = [d log_s, d b, 0...],
//        dskip [rows_p, 512] = do We^T
__global__ void coupling_bwd_kernel(const float* __restrict__ skip, const float* __restrict__ We, const float* __restrict__ be,
                                    const float* __restrict__ y, const float* __restrict__ dxo, float coef_ls, float* __restrict__ dy,
                                    float* __restrict__ do_pad, float* __restrict__ dskip, int N, int T, int c) {
  __shared__ float we_s[kWnCh * 8];
  const int half = c / 2, Tp = T + 2 * kWgPad;
  for (int i = threadIdx.x; i < kWnCh * c; i += blockDim.x) we_s[i] = We[i];
  __syncthreads();
  const size_t rows = (size_t)N * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (size_t r = (size_t)blockIdx.x * wpb + warp; r < rows; r += (size_t)gridDim.x * wpb) {
    const size_t prow = (r / T) * Tp + kWgPad + (r % T);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < kWnCh; k += 32) {
      const float sv = skip[prow * kWnCh + k];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < c) acc[j] = fmaf(sv, we_s[k * c + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    // every lane now holds o (without bias) for all c outputs; lane j < half handles coupling channel j
    float dov[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dov[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < half) {
        const float ls_raw = acc[j] + be[j];
        const float ls = fminf(ls_raw, 8.0f);
        const float e = expf(ls);
        const float x1 = y[r * c + half + j];
        const float dx1p = dxo[r * c + half + j];
        dov[j] = ls_raw <= 8.0f ? dx1p * e * x1 + coef_ls : 0.f;  // d log_s (the clamp passes no gradient)
        dov[half + j] = dx1p;                                     // d b
        if (lane == 0) {
          dy[r * c + half + j] = dx1p * e;
          dy[r * c + j] = dxo[r * c + j];
        }
      }
    }
    if (lane < 8) {
      float vsel = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j == lane) vsel = dov[j];
      do_pad[prow * 8 + lane] = lane < c ? vsel : 0.f;
    }
    for (int k = lane; k < kWnCh; k += 32) {
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < c) sacc = fmaf(dov[j], we_s[k * c + j], sacc);
      dskip[prow * kWnCh + k] = sacc;
    }
  }
}

// d rs of a layer in the stacked operand layout: [d_h | d_skip] (R = 1024) or d_skip alone (last layer, R = 512); valid rows
__global__ void stack_drs_kernel(const float* __restrict__ dh, const float* __restrict__ dskip, __nv_bfloat16* __restrict__ drs3, int N, int T,
                                 int lastl) {
  const int Tp = T + 2 * kWgPad;
  const int R = lastl ? kWnCh : 2 * kWnCh;
  const int R4 = R / 4;
  const size_t n = (size_t)N * T * R4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % R4) * 4;
    const size_t nt = i / R4;
    const size_t row = (nt / T) * Tp + kWgPad + (nt % T);
    float4 v;
    if (!lastl && ch < kWnCh)
      v = *reinterpret_cast<const float4*>(dh + row * kWnCh + ch);
    else
      v = *reinterpret_cast<const float4*>(dskip + row * kWnCh + (lastl ? ch : ch - kWnCh));
    const float x[4] = {v.x, v.y, v.z, v.w};
    store_x3_vec4(drs3 + row * 3 * R, R, ch, x);
  }
}

// gate backward over the valid rows: g = tanh(at) sigmoid(as), at/as = a + b_in + b_cond;
// dg_tot = dg (through the res/skip conv) + dh (direct residual path, layers < 7)  ->  da fp32 + stacked bf16
__global__ void gate_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b_in, const float* __restrict__ b_cond,
                                const float* __restrict__ dg, const float* __restrict__ dh, float* __restrict__ da,
                                __nv_bfloat16* __restrict__ da3, int N, int T) {
  const int Tp = T + 2 * kWgPad;
  constexpr int C4 = kWnCh / 4;
  const size_t n = (size_t)N * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C4) * 4;
    const size_t nt = i / C4;
    const size_t row = (nt / T) * Tp + kWgPad + (nt % T);
    const float4 at = *reinterpret_cast<const float4*>(a + row * 2 * kWnCh + ch);
    const float4 as = *reinterpret_cast<const float4*>(a + row * 2 * kWnCh + kWnCh + ch);
    const float4 bt1 = *reinterpret_cast<const float4*>(b_in + ch), bt2 = *reinterpret_cast<const float4*>(b_cond + ch);
    const float4 bs1 = *reinterpret_cast<const float4*>(b_in + kWnCh + ch), bs2 = *reinterpret_cast<const float4*>(b_cond + kWnCh + ch);
    float4 d = *reinterpret_cast<const float4*>(dg + row * kWnCh + ch);
    if (dh) {
      const float4 e = *reinterpret_cast<const float4*>(dh + row * kWnCh + ch);
      d = make_float4(d.x + e.x, d.y + e.y, d.z + e.z, d.w + e.w);
    }
    const float t4[4] = {at.x + bt1.x + bt2.x, at.y + bt1.y + bt2.y, at.z + bt1.z + bt2.z, at.w + bt1.w + bt2.w};
    const float s4[4] = {as.x + bs1.x + bs2.x, as.y + bs1.y + bs2.y, as.z + bs1.z + bs2.z, as.w + bs1.w + bs2.w};
    const float d4[4] = {d.x, d.y, d.z, d.w};
    float dt[4], ds[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float th = tanhf(t4[j]);
      const float sg = 1.f / (1.f + expf(-s4[j]));
      dt[j] = d4[j] * sg * (1.f - th * th);
      ds[j] = d4[j] * th * sg * (1.f - sg);
    }
    *reinterpret_cast<float4*>(da + row * 2 * kWnCh + ch) = make_float4(dt[0], dt[1], dt[2], dt[3]);
    *reinterpret_cast<float4*>(da + row * 2 * kWnCh + kWnCh + ch) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    store_x3_vec4(da3 + row * 3 * 2 * kWnCh, 2 * kWnCh, ch, dt);
    store_x3_vec4(da3 + row * 3 * 2 * kWnCh, 2 * kWnCh, kWnCh + ch, ds);
  }
}

// start conv backward, one warp per row: dx0[j] += sum_ch dh0[ch] Ws[j][ch]; per-block partials of dWs[j][ch] = sum x0[j] dh0[ch]
__global__ void start_bwd_kernel(const float* __restrict__ dh0, const float* __restrict__ Ws, const float* __restrict__ y, float* __restrict__ dy,
                                 float* __restrict__ part, int N, int T, int c) {
  __shared__ float ws_s[4 * kWnCh];
  __shared__ float red[8][4 * kWnCh / 8];  // reused in 8 chunks
  const int half = c / 2, Tp = T + 2 * kWgPad;
  for (int i = threadIdx.x; i < half * kWnCh; i += blockDim.x) ws_s[i] = Ws[i];
  __syncthreads();
  const size_t rows = (size_t)N * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float accw[4][16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 16; ++k) accw[j][k] = 0.f;
  for (size_t r = (size_t)blockIdx.x * wpb + warp; r < rows; r += (size_t)gridDim.x * wpb) {
    const size_t prow = (r / T) * Tp + kWgPad + (r % T);
    float x0[4], dx0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x0[j] = j < half ? y[r * c + j] : 0.f;
      dx0[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int ch = lane + 32 * k;
      const float d = dh0[prow * kWnCh + ch];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < half) {
          dx0[j] = fmaf(d, ws_s[j * kWnCh + ch], dx0[j]);
          accw[j][k] = fmaf(x0[j], d, accw[j][k]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tot = warp_sum(dx0[j]);
      if (lane == 0 && j < half) dy[r * c + j] += tot;
    }
  }
  // block reduction of the dWs partials, fixed order over the 8 warps
  float* sh = &red[0][0];  // [8 warps][32 lanes] per (j, k)
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 16; ++k) {
      __syncthreads();
      sh[warp * 32 + lane] = accw[j][k];
      __syncthreads();
      if (warp == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) tot += sh[w2 * 32 + lane];
        part[(size_t)blockIdx.x * 4 * kWnCh + j * kWnCh + lane + 32 * k] = tot;
      }
    }
}

// invertible 1x1 backward: dx = dy W^T; per-block partials of dW[i][j] = sum_rows x[row][i] dy[row][j]
__global__ void inv1x1_bwd_kernel(const float* __restrict__ xin, const float* __restrict__ dy, const float* __restrict__ Wm,
                                  float* __restrict__ dx, float* __restrict__ part, size_t rows, int c) {
  __shared__ float w_s[64];
  __shared__ float red[256];
  for (int i = threadIdx.x; i < c * c; i += blockDim.x) w_s[i] = Wm[i];
  __syncthreads();
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float xv[8], dv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xv[i] = i < c ? xin[r * c + i] : 0.f;
      dv[i] = i < c ? dy[r * c + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < c) {
        float sacc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < c) sacc = fmaf(dv[j], w_s[i * c + j], sacc);
        dx[r * c + i] = sacc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(xv[i], dv[j], acc[i * 8 + j]);
      }
    }
  }
  for (int e = 0; e < 64; ++e) {
    __syncthreads();
    red[threadIdx.x] = acc[e];
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int t2 = 0; t2 < (int)blockDim.x; ++t2) tot += red[t2];
      part[(size_t)blockIdx.x * 64 + e] = tot;
    }
  }
}
// dW[i][j] (c x c) from the [nblk][64] partials (8-wide rows)
__global__ void inv1x1_dw_final_kernel(const float* __restrict__ part, int nblk, int c, float* __restrict__ dW) {
  const int e = threadIdx.x;
  if (e >= 64) return;
  const int i = e / 8, j = e % 8;
  if (i >= c || j >= c) return;
  float tot = 0.f;
  for (int b = 0; b < nblk; ++b) tot += part[(size_t)b * 64 + e];
  dW[i * c + j] = tot;
}

// padded [rows_p, C] valid rows -> dense [rows, C]
__global__ void unpad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int T, int C) {
  const int Tp = T + 2 * kWgPad;
  const size_t n = (size_t)N * T * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % C);
    const size_t nt = i / C;
    dst[i] = src[((nt / T) * Tp + kWgPad + (nt % T)) * C + cc];
  }
}

static void colsum_valid(cudaStream_t s, const float* in, int ld, int C, float* out, float* out2, float* part, int N, int T) {
  colsum_valid_partial_kernel<<<dim3((C + 31) / 32, kWgPartSlices), dim3(32, 8), 0, s>>>(in, ld, part, N, T, C);
  sum_partials_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, kWgPartSlices, C, out, 0);
  if (out2) sum_partials_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, kWgPartSlices, C, out2, 0);
}

// P [Mr rows = padded rows 128 .. rows_p - 129][3 x 512 | 640]: d h(u) = sum_k P_k(u - (k-1) d) (rows outside the utterance
// contribute nothing: their d a is zero), d mel(u) += P_cond(u).  One thread = 4 channels of one valid position.
__global__ void shift_add_kernel(const float* __restrict__ P, float* __restrict__ dh, float* __restrict__ dmel, int N, int T, int d) {
  const int Tp = T + 2 * kWgPad;
  constexpr int W4 = kWnK1 / 4, H4 = kWnCh / 4;  // float4 columns of a P row / of the d h part
  const size_t n = (size_t)N * T * (H4 + kWnMel / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (H4 + kWnMel / 4));
    const size_t nt = i / (H4 + kWnMel / 4);
    const int t = (int)(nt % T), u = (int)(nt / T);
    const size_t prow = (size_t)u * Tp + kWgPad + t;   // padded row; P row = prow - kWgPad
    const float4* Pr = reinterpret_cast<const float4*>(P);
    if (c4 < H4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ts = t - (k - 1) * d;
        if (ts >= 0 && ts < T) {
          const float4 v = Pr[((size_t)u * Tp + ts) * W4 + k * H4 + c4];  // P row of padded row (u Tp + 128 + ts)
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      reinterpret_cast<float4*>(dh)[prow * H4 + c4] = acc;
    } else if (dmel) {
      const int m4 = c4 - H4;
      const float4 v = Pr[(prow - kWgPad) * W4 + 3 * H4 + m4];
      float4* o = reinterpret_cast<float4*>(dmel) + prow * (kWnMel / 4) + m4;
      float4 c = *o;
      c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
      *o = c;
    }
  }
}

// dW = A_hi^T B_hi + A_lo^T B_hi + A_hi^T B_lo over `rows` rows of two stacked operands (block widths ka / nb)
static int wgrad_x3(cudaStream_t s, int ka, int nb, int rows, const __nv_bfloat16* A3, int lda, const __nv_bfloat16* B3, int ldb, float* C) {
  return tc_gemm_hl(s, true, false, ka, nb, rows, A3, A3 + ka, lda, 0, B3, B3 + nb, ldb, 0, C, nb, 0, 0.f, 1);
}

extern "C" int mstts_waveglow_train_bwd(const MsttsWaveGlowWeights* w, const MsttsWaveGlowGrads* dwt, const float* z, int N, int T,
                                        float sigma, float* d_mel_nt640, void* ws_, size_t ws_bytes, void* stream_) {
  MSTTS_REQUIRE(w && dwt && z && ws_, MSTTS_E_INVALID, "waveglow bwd: null argument");
  MSTTS_REQUIRE(N >= 1 && T >= 1 && sigma > 0.f, MSTTS_E_INVALID, "waveglow bwd: N=%d T=%d sigma=%g", N, T, sigma);
  const WgLayout l = wg_layout(N, T);
  const WgSave v = wg_save_layout(N, T, l.total);
  const WgBwdLayout b = wg_bwd_layout(N, T, v.total);
  MSTTS_REQUIRE(ws_bytes >= b.total, MSTTS_E_WORKSPACE, "waveglow bwd: workspace %zu < %zu", ws_bytes, b.total);
  cudaStream_t s = (cudaStream_t)stream_;
  char* ws = (char*)ws_;
  const int Tp = T + 2 * kWgPad;
  const size_t rows = (size_t)N * T, rows_p = (size_t)N * Tp;
  auto BF = [&](size_t off) { return (__nv_bfloat16*)(ws + off); };
  auto FP = [&](size_t off) { return (float*)(ws + off); };
  const size_t slot3 = rows_p * 3 * kWnCh * 2, slota = rows_p * 2 * kWnCh * 4;
  auto H3 = [&](int f, int i) { return BF(v.h3 + ((size_t)f * kWnLayers + i) * slot3); };
  auto G3 = [&](int f, int i) { return BF(v.g3 + ((size_t)f * kWnLayers + i) * slot3); };
  auto APRE = [&](int f, int i) { return FP(v.a + ((size_t)f * kWnLayers + i) * slota); };
  auto SKIP = [&](int f) { return FP(v.skip + (size_t)f * rows_p * kWnCh * 4); };
  auto YBUF = [&](int f) { return FP(v.y + (size_t)f * rows * 8 * 4); };
  auto XIN = [&](int f) { return FP(v.xin + (size_t)f * rows * 8 * 4); };
  int rc;
  const int Mr = (int)(rows_p - 2 * kWgPad);
  const int K3 = 3 * kWnCh, A3 = 3 * 2 * kWnCh;  // stacked widths of 512- and 1024-wide rows
  const float n_el = (float)rows * 8.f;

  // ---- column-stacked effective weights (the scales are still in the workspace from the forward pass) ----
  for (int f = 0; f < kWgFlows; ++f) {
    WnJobs J;
    memset(&J, 0, sizeof(J));
    J.scale = FP(l.wscale) + (size_t)f * kWnJobsPerFlow * 1024;
    __nv_bfloat16* wq = BF(b.wqc) + (size_t)f * kWgFlowW * 3;
    int nj = 0;
    J.j[nj++] = WnJob{w->start_v[f], w->start_g[f], nullptr, nullptr, flow_c(f) / 2, kWnCh, flow_c(f) / 2};
    size_t wo = 0;
    for (int i = 0; i < kWnLayers; ++i) {
      J.j[nj++] = WnJob{w->in_v[f][i], w->in_g[f][i], nullptr, wq + wo, 3 * kWnCh, 2 * kWnCh, kWnCh};
      wo += (size_t)9 * kWnCh * 2 * kWnCh;
      J.j[nj++] = WnJob{w->cond_v[f][i], w->cond_g[f][i], nullptr, wq + wo, kWnMel, 2 * kWnCh, kWnMel};
      wo += (size_t)3 * kWnMel * 2 * kWnCh;
      const int rout = i < kWnLayers - 1 ? 2 * kWnCh : kWnCh;
      J.j[nj++] = WnJob{w->res_v[f][i], w->res_g[f][i], nullptr, wq + wo, kWnCh, rout, kWnCh};
      wo += (size_t)3 * kWnCh * rout;
    }
    wn_apply_cols_kernel<<<dim3(64, kWnJobsPerFlow), 256, 0, s>>>(J);
  }
  MSTTS_CUDA(cudaMemsetAsync(ws + b.drs3, 0, rows_p * A3 * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + b.drs3_last, 0, rows_p * K3 * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + b.da3, 0, rows_p * A3 * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + b.dmel, 0, rows_p * kWnMel * 4, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + b.dopad, 0, rows_p * 8 * 4, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + b.dskip, 0, rows_p * kWnCh * 4, s));

  // operand images and product buffers of the merged per-layer products (stream-ordered scratch, one set per call)
  ScratchScope wg_scope(s);
  const int KbR = (Mr + 63) / 64;
  void *wimgA = nullptr, *daT = nullptr, *daA = nullptr, *wimgB = nullptr;
  float *dw_all = nullptr, *Pbuf = nullptr;
  if ((rc = wg_scope.get(&wimgA, tc_image_bytes(kWnK1, Mr, 128)))) return rc;     // [3 taps x 512 | mel 640] rows, K = positions
  if ((rc = wg_scope.get(&daT, tc_image_bytes(2 * kWnCh, Mr, 256)))) return rc;   // d a^T
  if ((rc = wg_scope.get(&daA, tc_image_bytes(Mr, 2 * kWnCh, 128)))) return rc;   // d a
  if ((rc = wg_scope.get(&wimgB, tc_image_bytes(9 * 256, 2 * kWnCh, 256)))) return rc;  // 3 x 512 tap rows + 640 cond rows (padded to 768)
  if ((rc = wg_scope.get((void**)&dw_all, (size_t)kWnK1 * 2 * kWnCh * sizeof(float)))) return rc;
  if ((rc = wg_scope.get((void**)&Pbuf, (size_t)Mr * kWnK1 * sizeof(float)))) return rc;
  MSTTS_CUDA(cudaMemsetAsync(wimgB, 0, tc_image_bytes(9 * 256, 2 * kWnCh, 256), s));  // the 128 pad rows of the last n-tile stay zero
  {
    const __nv_bfloat16* m3 = BF(l.mel3) + (size_t)kWgPad * 3 * kWnMel;
    if ((rc = tc_pack_hl(s, m3, m3 + kWnMel, 3 * kWnMel, true, kWnMel, Mr, 128, KbR, wimgA, 12, 0))) return rc;
  }

  float* dz = FP(l.xa);  // forward scratch, free now
  dz_kernel<<<ew_grid(rows * 8), 256, 0, s>>>(z, 1.f / (sigma * sigma * n_el), rows * 8, dz);
  float* dxcur = FP(b.dx2);
  copy_channels_kernel<<<ew_grid(rows * 4), 256, 0, s>>>(dz, 8, 4, dxcur, 4, 0, 4, rows);
  const int nblk_small = 148 * 2;

  for (int f = kWgFlows - 1; f >= 0; --f) {
    const int c = flow_c(f), half = c / 2;
    const __nv_bfloat16* wqc = BF(b.wqc) + (size_t)f * kWgFlowW * 3;
    coupling_bwd_kernel<<<148 * 2, 256, 0, s>>>(SKIP(f), w->end_w[f], w->end_b[f], YBUF(f), dxcur, -1.f / n_el, FP(b.dy), FP(b.dopad),
                                               FP(b.dskip), N, T, c);
    // end conv: dWe = skip^T do, dbe = colsum(do)   (512 x c output over all positions: the 3-way-split precision level)
    if ((rc = gemm_rowmajor_p(s, TC_PRECISE, true, false, kWnCh, c, (int)rows_p, SKIP(f), kWnCh, FP(b.dopad), 8, dwt->end_w[f], c, 0.f))) return rc;
    colsum_valid(s, FP(b.dopad), 8, c, dwt->end_b[f], nullptr, FP(b.part), N, T);
    int cur = 0;
    // per-layer offsets inside the flow's stacked weight image
    size_t wo_layer[kWnLayers];
    {
      size_t wo = 0;
      for (int i = 0; i < kWnLayers; ++i) {
        wo_layer[i] = wo;
        wo += (size_t)9 * kWnCh * 2 * kWnCh + (size_t)3 * kWnMel * 2 * kWnCh + (size_t)3 * kWnCh * (i < kWnLayers - 1 ? 2 * kWnCh : kWnCh);
      }
    }
    for (int i = kWnLayers - 1; i >= 0; --i) {
      const int d = 1 << i;
      const bool lastl = i == kWnLayers - 1;
      const int R = lastl ? kWnCh : 2 * kWnCh;
      const float* dh_in = lastl ? nullptr : FP(b.dh[cur]);
      const __nv_bfloat16* Wtap_c = wqc + wo_layer[i];
      const __nv_bfloat16* Wc_c = Wtap_c + (size_t)9 * kWnCh * 2 * kWnCh;
      const __nv_bfloat16* Wr_c = Wc_c + (size_t)3 * kWnMel * 2 * kWnCh;
      __nv_bfloat16* drs3 = lastl ? BF(b.drs3_last) : BF(b.drs3);
      stack_drs_kernel<<<ew_grid(rows * R / 4), 256, 0, s>>>(dh_in, FP(b.dskip), drs3, N, T, lastl ? 1 : 0);
      // d res bias = colsum of d rs = [colsum(dh) | colsum(dskip)]; d skip is the same for all 8 layers of the flow: its
      // column sums are computed once (last layer) and copied
      if (!lastl) {
        colsum_valid(s, dh_in, kWnCh, kWnCh, dwt->res_b[f][i], nullptr, FP(b.part), N, T);
        MSTTS_CUDA(cudaMemcpyAsync(dwt->res_b[f][i] + kWnCh, dwt->res_b[f][kWnLayers - 1], kWnCh * sizeof(float), cudaMemcpyDeviceToDevice, s));
      } else {
        colsum_valid(s, FP(b.dskip), kWnCh, kWnCh, dwt->res_b[f][i], nullptr, FP(b.part), N, T);
      }
      // d W_res (effective) = g^T d rs ; then through the weight norm
      if ((rc = wgrad_x3(s, kWnCh, R, Mr, G3(f, i) + (size_t)kWgPad * K3, K3, drs3 + (size_t)kWgPad * 3 * R, 3 * R, FP(b.dw_eff)))) return rc;
      wn_bwd_kernel<<<(R + 31) / 32, dim3(32, 32), 0, s>>>(w->res_v[f][i], w->res_g[f][i], FP(b.dw_eff), kWnCh, R, dwt->res_g[f][i],
                                                          dwt->res_v[f][i]);
      // d g = d rs W_res^T
      if ((rc = gemm_stacked_nt(s, Mr, kWnCh, R, drs3 + (size_t)kWgPad * 3 * R, Wr_c, FP(b.dg) + (size_t)kWgPad * kWnCh, kWnCh, 0.f))) return rc;
      gate_bwd_kernel<<<ew_grid(rows * kWnCh / 4), 256, 0, s>>>(APRE(f, i), w->in_b[f][i], w->cond_b[f][i], FP(b.dg), dh_in, FP(b.da), BF(b.da3),
                                                            N, T);
      colsum_valid(s, FP(b.da), 2 * kWnCh, 2 * kWnCh, dwt->in_b[f][i], dwt->cond_b[f][i], FP(b.part), N, T);
      const __nv_bfloat16* da3p = BF(b.da3) + (size_t)kWgPad * A3;
      // ---- weight gradients of the dilated conv and the conditioning conv: ONE product ----
      //   dW[(tap k, ci) | mel m][co] = sum over rows  [h(t + (k-1) d) | mel(t)]^T  d a(t)        (M = 2176, N = 1024, K = rows)
      // operand rows: three transposed row-shifted views of the layer input + the transposed conditioning (packed once per
      // call); the d a operand is packed once and shared
      for (int k = 0; k < 3; ++k) {
        const __nv_bfloat16* hk = H3(f, i) + (long long)(kWgPad + (k - 1) * d) * K3;
        if ((rc = tc_pack_hl(s, hk, hk + kWnCh, K3, true, kWnCh, Mr, 128, KbR, wimgA, 4 * k, 0))) return rc;
      }
      if ((rc = tc_pack_hl(s, da3p, da3p + 2 * kWnCh, A3, true, 2 * kWnCh, Mr, 256, KbR, daT, 0, 0))) return rc;
      if ((rc = tc_gemm_images(s, wimgA, daT, kWnK1, 2 * kWnCh, Mr, dw_all, 2 * kWnCh, 0.f))) return rc;
      wn_bwd_kernel<<<(2 * kWnCh + 31) / 32, dim3(32, 32), 0, s>>>(w->cond_v[f][i], w->cond_g[f][i], dw_all + (size_t)3 * kWnCh * 2 * kWnCh, kWnMel,
                                                                  2 * kWnCh, dwt->cond_g[f][i], dwt->cond_v[f][i]);
      wn_bwd_kernel<<<(2 * kWnCh + 31) / 32, dim3(32, 32), 0, s>>>(w->in_v[f][i], w->in_g[f][i], dw_all, 3 * kWnCh, 2 * kWnCh,
                                                                  dwt->in_g[f][i], dwt->in_v[f][i]);
      // ---- activation-side products: ONE product P = d a . [W_in[0]^T | W_in[1]^T | W_in[2]^T | W_cond^T]  (N = 2176, K = 1024),
      //      then d h(u) = sum_k P_k(u - (k-1) d)  and  d mel += P_cond  in one pass ----
      if ((rc = tc_pack_hl(s, da3p, da3p + 2 * kWnCh, A3, false, Mr, 2 * kWnCh, 128, 2 * kWnCh / 64, daA, 0, 0))) return rc;
      for (int k = 0; k < 3; ++k) {
        const __nv_bfloat16* wk = Wtap_c + (size_t)k * kWnCh * A3;
        if ((rc = tc_pack_hl(s, wk, wk + 2 * 2 * kWnCh, A3, false, kWnCh, 2 * kWnCh, 256, 2 * kWnCh / 64, wimgB, 2 * k, 0))) return rc;
      }
      if ((rc = tc_pack_hl(s, Wc_c, Wc_c + 2 * 2 * kWnCh, A3, false, kWnMel, 2 * kWnCh, 256, 2 * kWnCh / 64, wimgB, 6, 0))) return rc;
      if ((rc = tc_gemm_images(s, daA, wimgB, Mr, kWnK1, 2 * kWnCh, Pbuf, kWnK1, 0.f))) return rc;
      const int nxt = cur ^ 1;
      shift_add_kernel<<<ew_grid(rows * (kWnK1 / 4)), 256, 0, s>>>(Pbuf, FP(b.dh[nxt]), d_mel_nt640 ? FP(b.dmel) : nullptr, N, T, d);
      cur = nxt;
    }
    // ---- start conv (weight-normed 1x1, c/2 -> 512) ----
    const float* dh0 = FP(b.dh[cur]);
    colsum_valid(s, dh0, kWnCh, kWnCh, dwt->start_b[f], nullptr, FP(b.part), N, T);
    float* part2 = FP(b.part) + (size_t)kWgPartSlices * 4 * 2 * kWnCh;
    start_bwd_kernel<<<nblk_small, 256, 0, s>>>(dh0, FP(l.start_eff) + (size_t)f * 4 * kWnCh, YBUF(f), FP(b.dy), part2, N, T, c);
    sum_partials_kernel<<<(4 * kWnCh + 127) / 128, 128, 0, s>>>(part2, nblk_small, 4 * kWnCh, FP(b.dw_eff), 0);
    wn_bwd_kernel<<<(kWnCh + 31) / 32, dim3(32, 32), 0, s>>>(w->start_v[f], w->start_g[f], FP(b.dw_eff), half, kWnCh, dwt->start_g[f],
                                                            dwt->start_v[f]);
    // ---- invertible 1x1 ----
    inv1x1_bwd_kernel<<<nblk_small, 256, 0, s>>>(XIN(f), FP(b.dy), w->inv_w[f], FP(b.dx), part2, rows, c);
    inv1x1_dw_final_kernel<<<1, 64, 0, s>>>(part2, nblk_small, c, dwt->inv_w[f]);
    dxcur = FP(b.dx);
    if (f % 4 == 0 && f > 0) {  // the two early-output channels rejoin the chain (Modules.py:334-336 in reverse)
      float* nx = FP(b.dx2);
      copy_channels_kernel<<<ew_grid(rows * 2), 256, 0, s>>>(dz, 8, f == 8 ? 2 : 0, nx, c + 2, 0, 2, rows);
      copy_channels_kernel<<<ew_grid(rows * c), 256, 0, s>>>(dxcur, c, 0, nx, c + 2, 2, c, rows);
      dxcur = nx;
    }
  }
  if (d_mel_nt640) unpad_rows_kernel<<<ew_grid(rows * kWnMel), 256, 0, s>>>(FP(b.dmel), d_mel_nt640, N, T, kWnMel);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

// Upsample_Mel backward: d kernel [1024, out, in], d bias [80] from d up [N, keep, 80] (zero beyond keep)
extern "C" size_t mstts_upsample_mel_bwd_workspace_bytes(int N, int Tm) {
  if (N <= 0 || Tm <= 0) return 0;
  return (size_t)N * ((size_t)(Tm - 1) * 256 + 1024) * 80 * sizeof(float) + 64 * 80 * sizeof(float) + 256;
}
__global__ void pad_dup_kernel(const float* __restrict__ dup, float* __restrict__ dst, int N, int keep, int Lfull) {
  const size_t n = (size_t)N * Lfull * 80;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % 80);
    const size_t np = i / 80;
    const int p = (int)(np % Lfull), nn = (int)(np / Lfull);
    dst[i] = p < keep ? dup[((size_t)nn * keep + p) * 80 + co] : 0.f;
  }
}
__global__ void colsum_dense_partial_kernel(const float* __restrict__ in, float* __restrict__ part, size_t R, int C) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t r0 = R * blockIdx.y / 64, r1 = R * (blockIdx.y + 1) / 64;
  float sacc = 0.f;
  if (c < C)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) sacc += in[r * C + c];
  sh[threadIdx.y][threadIdx.x] = sacc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += sh[j][threadIdx.x];
    part[(size_t)blockIdx.y * C + c] = tot;
  }
}
extern "C" int mstts_upsample_mel_bwd(const float* mel, const float* d_up, int N, int Tm, int keep, float* d_kernel, float* d_bias, void* ws_,
                                      size_t ws_bytes, void* stream) {
  MSTTS_REQUIRE(mel && d_up && d_kernel && d_bias && ws_, MSTTS_E_INVALID, "upsample_mel_bwd: null pointer");
  const int Lfull = (Tm - 1) * 256 + 1024;
  MSTTS_REQUIRE(keep >= 1 && keep <= Lfull, MSTTS_E_INVALID, "upsample_mel_bwd: keep=%d outside [1,%d]", keep, Lfull);
  MSTTS_REQUIRE(ws_bytes >= mstts_upsample_mel_bwd_workspace_bytes(N, Tm), MSTTS_E_WORKSPACE, "upsample_mel_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* dpad = (float*)ws_;
  float* part = dpad + (size_t)N * Lfull * 80;
  pad_dup_kernel<<<ew_grid((size_t)N * Lfull * 80), 256, 0, s>>>(d_up, dpad, N, keep, Lfull);
  colsum_dense_partial_kernel<<<dim3(3, 64), dim3(32, 8), 0, s>>>(dpad, part, (size_t)N * Lfull, 80);
  sum_partials_kernel<<<1, 128, 0, s>>>(part, 64, 80, d_bias, 0);
  // dK[(k,co), ci] = sum_{n,t} dpad[n, 256 t + k, co] mel[n, t, ci].  With k = 256 q + r the padded gradient of one utterance
  // is a [(Tm+3), 256*80] matrix and quarter q of the kernel contracts its rows q .. q+Tm-1 with mel: 4 GEMMs per utterance.
  int rc;
  for (int n = 0; n < N; ++n)
    for (int q = 0; q < 4; ++q)
      if ((rc = gemm_rowmajor_ex(s, true, false, 256 * 80, 80, Tm, dpad + ((size_t)n * Lfull + 256 * q) * 80, 256 * 80,
                                 mel + (size_t)n * Tm * 80, 80, d_kernel + (size_t)q * 256 * 80 * 80, 80, n == 0 ? 0.f : 1.f)))
        return rc;
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
