// Shared host/device helpers for libmstts_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mstts_b200.h"

// ---- error plumbing -------------------------------------------------------------------------
void mstts_set_error(const char* fmt, ...);
void mstts_timer_start(int which, cudaStream_t s);  // no-ops unless mstts_set_profiling(1)
void mstts_timer_stop(int which, cudaStream_t s);

#define MSTTS_CUDA(call)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      mstts_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));      \
      return MSTTS_E_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define MSTTS_REQUIRE(cond, code, ...)                                                           \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      mstts_set_error(__VA_ARGS__);                                                              \
      return (code);                                                                             \
    }                                                                                            \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- model constants (Hyper_Parameters.py:7,32-54) ------------------------------------------
constexpr int kMel = 80;
constexpr int kPrenet = 256;
constexpr int kCell = 1024;
constexpr int kGates = 4 * kCell;
constexpr int kAtt = 128;
constexpr int kConvK = 31;
constexpr int kConvC = 32;
constexpr float kZoneKeep = 0.9f;  // 1 - Zoneout_Rate, applied in train AND inference (ZoneoutLSTMCell.py:259)
constexpr float kForgetBias = 1.0f;

// persistent decoder grid: 32 clusters x 4 CTAs.  Each CTA owns 8 LSTM units (32 gate columns) of both
// cells; each cluster owns one batch row of the attention phase.
constexpr int kDecGrid = 128;
constexpr int kDecCluster = 4;
constexpr int kDecThreads = 256;
constexpr int kUnitsPerCta = kCell / kDecGrid;  // 8

// launch attribute shared by the four persistent decoder launchers (grid barrier inside): cudaLaunchAttributeCooperative,
// valid together with the compile-time cluster dimensions.  MSTTS_NO_COOP=1 in the environment drops it (A/B measurements).
static inline void dec_cooperative_attr(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr) {
  static const bool off = [] {
    const char* e = getenv("MSTTS_NO_COOP");
    return e && e[0] == '1';
  }();
  if (off) return;
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float ld_nc_na(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier over a monotonically increasing counter (zeroed by the host before launch).
// All CTAs of the grid are co-resident (grid <= SM count, 1 CTA/SM, checked on the host).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    red_release_gpu_add(counter, 1u);
    while (ld_acquire_gpu(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif
