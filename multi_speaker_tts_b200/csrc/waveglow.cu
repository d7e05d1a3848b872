// WaveGlow affine-coupling flows (WaveGlow/Modules.py:210-371, WaveGlow/Inv1x1.py:9-32), forward (training direction,
// with the log-likelihood sums) and reverse (synthesis direction).
//
// Structure: the dense contractions of the WN stack (dilated k=3 conv 512->1024, mel conditioning 640->1024, res/skip
// 512->1024) run as bf16x3 tensor-core GEMMs through cuBLAS with fp32 accumulation (same numerics contract as the
// decoder kernels).  The three partial products of bf16x3 are folded into the K dimension: every activation row is stored
// as [hi | lo | hi] and every weight as [W_hi ; W_hi ; W_lo], so a product is ONE bf16 GEMM with K tripled instead of three
// GEMMs that each read-modify-write the 65 MB fp32 pre-activation (which made the K=512 calls HBM-bound: 29 us measured vs
// 12.5 us of tensor work).  Everything around the GEMMs is hand-written and fused:
//   wn_scale / wn_apply weight norm g*v/sqrt(max(sum v^2,1e-5)) -> bf16 hi/lo weight operands
//   flow_pre_kernel     invertible 1x1 (or plain split in reverse) + the K<=4 start conv -> bf16 hi/lo activations
//   gate_kernel         bias + tanh * sigmoid -> fp32 + bf16 hi/lo
//   resskip_kernel      residual onto the GATED activation (reference quirk) + skip accumulation
//   flow_post_kernel    512->c end conv + affine transform (clamp log_s at 8 in forward only) + sum(log_s), or the
//                       inverse transform followed by the inverse 1x1
// Activations live in a time-padded layout [N][T+2P][C], P = 128 zero rows on both sides of every utterance, so the
// dilated conv is three GEMMs over row-shifted views of one buffer (no im2col, no per-utterance launches).
// A single fused tcgen05 kernel per flow is the planned replacement ("what comes next").
#include "common.cuh"
#include "gemm.h"

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
// Step 1 (library SGEMM, fp32): C[n*Tm+t][k*80+co] = sum_ci mel[n,t,ci] K[k,co,ci]   ([N*Tm,80] x [81920,80]^T)
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
  size_t y, xa, xb;         // [N,T,8]
  size_t partial;           // doubles
  size_t total;
};
constexpr size_t kWgFlowW = (size_t)kWnLayers * (3 * kWnCh * 2 * kWnCh + kWnMel * 2 * kWnCh) + (size_t)(kWnLayers - 1) * kWnCh * 2 * kWnCh +
                            (size_t)kWnCh * kWnCh;

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

static inline int flow_c(int f) { return 8 - 2 * (f / 4); }

// Runs all 12 flows.  direction 0: training direction x -> z with sums[0] = sum(log_s), sums[1] = sum(z^2);
// direction 1: synthesis z -> x (early_noise[2]: the two [N,T,2] noise tensors injected before flows 7 and 3 are run,
// i.e. after undoing flows 8 and 4; inv_w then holds the INVERSE 1x1 kernels).
extern "C" int mstts_waveglow_flows(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T,
                                    int direction, const float* const* early_noise, float* out, double* sums, void* ws_, size_t ws_bytes,
                                    void* stream_) {
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
    wn_apply_kernel<<<dim3(64, kWnJobsPerFlow), 256, 0, s>>>(J);
  }
  // ---- conditioning operand + zeroed pads ----
  MSTTS_CUDA(cudaMemsetAsync(ws + l.mel3, 0, rows_p * 3 * kWnMel * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.h3, 0, rows_p * 3 * kWnCh * 2, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.g3, 0, rows_p * 3 * kWnCh * 2, s));
  pad_split_kernel<<<ew_grid((size_t)rows * kWnMel), 256, 0, s>>>(mel_nt640, N, T, kWnMel, BF(l.mel3));
  if (direction == 0) MSTTS_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s));

  const int M = (int)(rows_p - 2 * kWgPad);  // GEMM rows: everything except the outermost pads
  const float* xcur = audio_in;              // [N,T,c] of the current flow
  float* xbuf[2] = {FP(l.xa), FP(l.xb)};
  int xsel = 0;
  int zc0 = 0;  // forward: next output channel of z
  const int nblk_post = 148 * 2;
  for (int step = 0; step < kWgFlows; ++step) {
    const int f = direction == 0 ? step : kWgFlows - 1 - step;
    const int c = flow_c(f), half = c / 2;
    if (direction == 0 && f % 4 == 0 && f > 0) {
      // early output: first 2 channels leave the chain (Modules.py:334-336)
      copy_channels_kernel<<<ew_grid(rows * 2), 256, 0, s>>>(xcur, c + 2, 0, out, 8, zc0, 2, rows);
      zc0 += 2;
      float* nx = xbuf[xsel ^ 1];
      copy_channels_kernel<<<ew_grid(rows * c), 256, 0, s>>>(xcur, c + 2, 2, nx, c, 0, c, rows);
      xcur = nx;
      xsel ^= 1;
    }
    flow_pre_kernel<<<148 * 4, 256, 0, s>>>(xcur, w->inv_w[f], FP(l.start_eff) + (size_t)f * 4 * kWnCh, w->start_b[f], FP(l.y), BF(l.h3), N,
                                            T, c, direction == 0 ? 1 : 0);
    const __nv_bfloat16* wq = BF(l.wq) + (size_t)f * kWgFlowW * 3;
    size_t wo = 0;
    for (int i = 0; i < kWnLayers; ++i) {
      const int d = 1 << i;
      const int K3 = 3 * kWnCh;
      float* a_out = FP(l.a) + (size_t)kWgPad * 2 * kWnCh;
      const __nv_bfloat16* Wtap = wq + wo;
      const __nv_bfloat16* Wc = wq + wo + (size_t)3 * K3 * 2 * kWnCh;
      // conditioning first (beta = 0), then the three taps accumulate: 4 GEMMs with K = 1920 / 1536
      if ((rc = gemm_rowmajor_bf16(s, M, 2 * kWnCh, 3 * kWnMel, BF(l.mel3) + (size_t)kWgPad * 3 * kWnMel, 3 * kWnMel, Wc, 2 * kWnCh, a_out,
                                   2 * kWnCh, 0.f)))
        return rc;
      for (int k = 0; k < 3; ++k) {
        const long long shift = (long long)(kWgPad + (k - 1) * d) * K3;
        if ((rc = gemm_rowmajor_bf16(s, M, 2 * kWnCh, K3, BF(l.h3) + shift, K3, Wtap + (size_t)k * K3 * 2 * kWnCh, 2 * kWnCh, a_out,
                                     2 * kWnCh, 1.f)))
          return rc;
      }
      wo += (size_t)3 * K3 * 2 * kWnCh + (size_t)3 * kWnMel * 2 * kWnCh;
      gate_kernel<<<ew_grid(rows * kWnCh / 4), 256, 0, s>>>(FP(l.a), w->in_b[f][i], w->cond_b[f][i], FP(l.g), BF(l.g3), N, T);
      const int rout = i < kWnLayers - 1 ? 2 * kWnCh : kWnCh;
      // res/skip output reuses the pre-activation buffer (row stride = rout)
      if ((rc = gemm_rowmajor_bf16(s, M, rout, K3, BF(l.g3) + (size_t)kWgPad * K3, K3, wq + wo, rout, FP(l.a) + (size_t)kWgPad * rout, rout,
                                   0.f)))
        return rc;
      wo += (size_t)K3 * rout;
      resskip_kernel<<<ew_grid(rows * kWnCh / 4), 256, 0, s>>>(FP(l.a), w->res_b[f][i], FP(l.g), BF(l.h3), FP(l.skip), N, T, i == 0,
                                                           i == kWnLayers - 1);
    }
    float* xnext = xbuf[xsel ^ 1];
    flow_post_kernel<<<nblk_post, 256, 0, s>>>(FP(l.skip), w->end_w[f], w->end_b[f], FP(l.y), w->inv_w[f], xnext, (double*)(ws + l.partial),
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
