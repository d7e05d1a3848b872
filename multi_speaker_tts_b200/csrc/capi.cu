// Library-level entry points: version, error string, device check, and the small HBM-bound utility
// kernels (mask generator, TF-style Adam, decoder loss).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mstts_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional kernel timing (bench / roofline): events around the persistent kernels ----
// With profiling on, every launch of a persistent kernel is bracketed by a fresh event pair on the caller's
// stream (no synchronisation); mstts_kernel_ms(which) later waits for them and returns sum / count.
// The timers are process-wide (guarded by a mutex), not thread-local: torch.autograd runs the reverse pass on its own
// engine thread, and the caller reads the totals from the main thread.
#include <mutex>
#include <vector>
static int g_profiling = 0;
static std::mutex g_timer_mu;
struct KernelTimer {
  std::vector<cudaEvent_t> ev;  // start0, stop0, start1, stop1, ...
  size_t used = 0;
};
static KernelTimer g_timers[2];  // 0 = decoder forward loop, 1 = decoder reverse loop

void mstts_timer_start(int which, cudaStream_t s) {
  if (!g_profiling) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  KernelTimer& t = g_timers[which];
  if (t.used + 2 > t.ev.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    t.ev.push_back(a);
    t.ev.push_back(b);
  }
  cudaEventRecord(t.ev[t.used], s);
}
void mstts_timer_stop(int which, cudaStream_t s) {
  if (!g_profiling) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  KernelTimer& t = g_timers[which];
  cudaEventRecord(t.ev[t.used + 1], s);
  t.used += 2;
}

extern "C" int mstts_set_profiling(int on) {
  std::lock_guard<std::mutex> lk(g_timer_mu);
  g_profiling = on;
  g_timers[0].used = 0;
  g_timers[1].used = 0;
  return MSTTS_OK;
}
extern "C" int mstts_kernel_ms(int which, float* sum_ms, int* count) {
  MSTTS_REQUIRE(which >= 0 && which < 2 && sum_ms && count, MSTTS_E_INVALID, "kernel_ms: bad argument");
  std::lock_guard<std::mutex> lk(g_timer_mu);
  KernelTimer& t = g_timers[which];
  float total = 0.f;
  for (size_t i = 0; i + 1 < t.used; i += 2) {
    float ms = 0.f;
    MSTTS_CUDA(cudaEventSynchronize(t.ev[i + 1]));
    MSTTS_CUDA(cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]));
    total += ms;
  }
  *sum_ms = total;
  *count = (int)(t.used / 2);
  return MSTTS_OK;
}

// ---- stream-ordered scratch (scratch_pool.h) ----
#include "scratch_pool.h"
static int scratch_pool_ready(int dev) {
  static std::once_flag once[64];
  static cudaError_t err[64];
  std::call_once(once[dev & 63], [dev] {
    cudaMemPool_t pool;
    err[dev & 63] = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (err[dev & 63] == cudaSuccess) {
      uint64_t keep = UINT64_MAX;  // never hand the memory back between steps
      err[dev & 63] = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  });
  if (err[dev & 63] != cudaSuccess) {
    mstts_set_error("scratch pool: %s", cudaGetErrorString(err[dev & 63]));
    return MSTTS_E_CUDA;
  }
  return MSTTS_OK;
}
int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  int rc = scratch_pool_ready(dev);
  if (rc) return rc;
  MSTTS_CUDA(cudaMallocAsync(p, bytes < 256 ? 256 : bytes, s));
  return MSTTS_OK;
}
void scratch_free(void* p, cudaStream_t s) {
  if (p) cudaFreeAsync(p, s);
}
extern "C" int mstts_release_scratch(void) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  MSTTS_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  MSTTS_CUDA(cudaDeviceSynchronize());
  MSTTS_CUDA(cudaMemPoolTrimTo(pool, 0));
  return MSTTS_OK;
}

extern "C" int mstts_version(void) { return MSTTS_VERSION; }
extern "C" const char* mstts_last_error(void) { return g_err; }

extern "C" int mstts_device_check(int device) {
  cudaDeviceProp p;
  MSTTS_CUDA(cudaGetDeviceProperties(&p, device));
  MSTTS_REQUIRE(p.major == 10, MSTTS_E_DEVICE, "device %d is sm_%d%d, need sm_100", device, p.major, p.minor);
  return p.multiProcessorCount;
}

// ---- counter-based mask generator -----------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// 8 mask bytes per thread-iteration from one 64-bit hash (8 x 8-bit uniforms would be too coarse for
// keep=0.9, so each byte uses its own 32-bit draw from two hashes of the 4-element group).
__global__ void fill_mask_kernel(uint8_t* __restrict__ out, size_t n, uint32_t thresh, uint64_t seed) {
  const size_t ngroups = (n + 3) / 4;
  const uint64_t key = splitmix64(splitmix64(seed) ^ 0xD1B54A32D192ED03ull);
  for (size_t gidx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gidx < ngroups; gidx += (size_t)gridDim.x * blockDim.x) {
    // the seed is hashed into a key BEFORE the counter is mixed in: callers pass consecutive small seeds (step numbers), and
    // seed ^ counter would make the masks of different steps XOR-permutations of each other
    const uint64_t h0 = splitmix64(key + gidx * 2 + 0);
    const uint64_t h1 = splitmix64(key + gidx * 2 + 1);
    const uint32_t r[4] = {(uint32_t)h0, (uint32_t)(h0 >> 32), (uint32_t)h1, (uint32_t)(h1 >> 32)};
    const size_t base = gidx * 4;
    if (base + 3 < n && (((uintptr_t)(out + base)) & 3) == 0) {
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) packed |= (uint32_t)(r[j] < thresh) << (8 * j);
      *reinterpret_cast<uint32_t*>(out + base) = packed;
    } else {
      for (int j = 0; j < 4 && base + j < n; ++j) out[base + j] = r[j] < thresh;
    }
  }
}

extern "C" int mstts_fill_mask(uint8_t* out, size_t n, float keep_prob, uint64_t seed, void* stream) {
  MSTTS_REQUIRE(out || n == 0, MSTTS_E_INVALID, "fill_mask: null output");
  MSTTS_REQUIRE(keep_prob >= 0.f && keep_prob <= 1.f, MSTTS_E_INVALID, "fill_mask: keep_prob %f", keep_prob);
  if (n == 0) return MSTTS_OK;
  const double t = (double)keep_prob * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  size_t g = ((n + 3) / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  fill_mask_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(out, n, thresh, seed);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

// ---- TF-style Adam (epsilon outside the bias-corrected step) -----------------------------------
__global__ void adam_tf_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                               const float* __restrict__ g, size_t n, float lr_t, float b1, float b2, float eps,
                               float gs, float l2) {
  const size_t n4 = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = tid0; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pa = &pp.x;
    float* ma = &mm.x;
    float* va = &vv.x;
    const float* ga = &gg.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = ga[j] * gs + l2 * pa[j];
      ma[j] = b1 * ma[j] + (1.f - b1) * gj;
      va[j] = b2 * va[j] + (1.f - b2) * gj * gj;
      pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (size_t i = n4 * 4 + tid0; i < n; i += stride) {
    const float gj = g[i] * gs + l2 * p[i];
    const float mj = b1 * m[i] + (1.f - b1) * gj;
    const float vj = b2 * v[i] + (1.f - b2) * gj * gj;
    m[i] = mj;
    v[i] = vj;
    p[i] -= lr_t * mj / (sqrtf(vj) + eps);
  }
}

extern "C" int mstts_adam_tf(float* p, float* m, float* v, const float* g, size_t n, float lr_t, float b1, float b2,
                             float eps, float grad_scale, float l2, void* stream) {
  MSTTS_REQUIRE(p && m && v && g, MSTTS_E_INVALID, "adam: null pointer");
  MSTTS_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) == 0, MSTTS_E_INVALID,
                "adam: buffers must be 16-byte aligned");
  if (n == 0) return MSTTS_OK;
  size_t gsz = (n / 4 + 255) / 256;
  if (gsz > 148 * 8) gsz = 148 * 8;
  if (gsz == 0) gsz = 1;
  adam_tf_kernel<<<(int)gsz, 256, 0, (cudaStream_t)stream>>>(p, m, v, g, n, lr_t, b1, b2, eps, grad_scale, l2);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

// ---- decoder loss + gradient (MSTTS_SV.py:127-144) ---------------------------------------------
// loss2[0] = mean((lin-mel)^2) (+ mean|lin-mel|) over [B,L,80] using linear[:, :L];  loss2[1] = mean BCE over [B,T]
__global__ void decoder_loss_kernel(const float* __restrict__ linear, const float* __restrict__ stop,
                                    const float* __restrict__ mel, const int* __restrict__ mel_len, int B, int L, int T,
                                    int use_l1, float* __restrict__ part, float* __restrict__ d_linear,
                                    float* __restrict__ d_stop) {
  const size_t n_lin = (size_t)B * T * kMel;
  const size_t n_stop = (size_t)B * T;
  const float inv_lin = 1.f / ((float)B * (float)L * (float)kMel);
  const float inv_stop = 1.f / ((float)B * (float)T);
  float s_lin = 0.f, s_stop = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_lin + n_stop; i += stride) {
    if (i < n_lin) {
      const int c = (int)(i % kMel);
      const size_t bt = i / kMel;
      const int t = (int)(bt % T), b = (int)(bt / T);
      float gval = 0.f;
      if (t < L && t < T - 1) {
        const float d = linear[i] - mel[((size_t)b * L + t) * kMel + c];
        s_lin += d * d * inv_lin;
        gval = 2.f * d * inv_lin;
        if (use_l1) {
          s_lin += fabsf(d) * inv_lin;
          gval += (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_lin;
        }
      }
      d_linear[i] = gval;
    } else {
      const size_t j = i - n_lin;
      const int t = (int)(j % T), b = (int)(j / T);
      const float x = stop[j];
      const float z = (t >= mel_len[b]) ? 1.f : 0.f;
      s_stop += (fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)))) * inv_stop;
      d_stop[j] = (1.f / (1.f + expf(-x)) - z) * inv_stop;
    }
  }
  s_lin = warp_sum(s_lin);
  s_stop = warp_sum(s_stop);
  __shared__ float sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = s_lin;
    sh[1][warp] = s_stop;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += sh[0][w];
      b2 += sh[1][w];
    }
    part[blockIdx.x] = a;  // per-block partials, added in block order by decoder_loss_final_kernel (atomics would make the
    part[gridDim.x + blockIdx.x] = b2;  // last bits of the reported loss depend on arrival order)
  }
}

__global__ void decoder_loss_final_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ loss2) {
  const int lane = threadIdx.x;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < nblocks; i += 32) {
    a += part[i];
    b += part[nblocks + i];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    loss2[0] = a;
    loss2[1] = b;
  }
}

extern "C" int mstts_decoder_loss(const float* linear, const float* stop, const float* mel, const int32_t* mel_len, int B,
                                  int L, int n_steps, int use_l1, float* loss2, float* d_linear, float* d_stop,
                                  void* stream) {
  MSTTS_REQUIRE(linear && stop && mel && mel_len && loss2 && d_linear && d_stop, MSTTS_E_INVALID, "loss: null pointer");
  MSTTS_REQUIRE(n_steps == L + 1, MSTTS_E_INVALID,
                "loss: linear[:, :-1] must match mel: n_steps=%d, L=%d (MSTTS_SV.py:138)", n_steps, L);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)B * n_steps * (kMel + 1);
  size_t g = (n + 255) / 256;
  if (g > 148 * 4) g = 148 * 4;
  ScratchScope scope(s);
  void* part = nullptr;
  int rc = scope.get(&part, 2 * g * sizeof(float));
  if (rc) return rc;
  decoder_loss_kernel<<<(int)g, 256, 0, s>>>(linear, stop, mel, mel_len, B, L, n_steps, use_l1, (float*)part, d_linear, d_stop);
  decoder_loss_final_kernel<<<1, 32, 0, s>>>((const float*)part, (int)g, loss2);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
