// Standalone micro-benchmarks that size the persistent decoder kernels (run on the GPU box):
//   1. how many clusters of 4 / 8 CTAs (1 CTA/SM, ~200 KB smem) the device co-schedules
//   2. L2 -> shared streaming rate of a ring of cp.async.bulk slots (weights re-read every decoder step)
//   3. latency of the global-memory grid barrier
// Build: make probes.  Usage: stream_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../multi_speaker_tts_b200/csrc/sm100_ptx.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int kThreads = 512;

struct StreamParams {
  const uint8_t* w;      // [grid][slice_bytes]
  const uint8_t* x;      // shared activation image, x_bytes per tile (all CTAs read the same bytes)
  int slice_bytes, slot_w_bytes, slot_x_bytes, nslots, passes, nprod, split, poll;
  unsigned long long* cycles;  // [grid]
};

__global__ void __launch_bounds__(kThreads, 1) stream_kernel(const StreamParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int slot_bytes = P.slot_w_bytes + P.slot_x_bytes;
  if (tid == 0) {
    for (int i = 0; i < P.nslots; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int tiles_per_pass = P.slice_bytes / P.slot_w_bytes;
  const int ntiles = tiles_per_pass * P.passes;
  const uint8_t* wbase = P.w + (size_t)blockIdx.x * P.slice_bytes;
  long long t0 = clock64();
  if (warp >= 8 && warp < 8 + P.nprod && (tid & 31) == 0) {  // producer(s): thread j takes tiles j, j+nprod, ...
    for (int i = warp - 8; i < ntiles; i += P.nprod) {
      const int s = i % P.nslots, round = i / P.nslots;
      if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
      ptx::mbar_arrive_expect_tx(&full[s], slot_bytes);
      uint8_t* dst = smem + (size_t)s * slot_bytes;
      const int part = P.slot_w_bytes / P.split;
      for (int q = 0; q < P.split; ++q)
        ptx::bulk_g2s(dst + q * part, wbase + (size_t)(i % tiles_per_pass) * P.slot_w_bytes + q * part, part, &full[s]);
      if (P.slot_x_bytes)
        ptx::bulk_g2s(dst + P.slot_w_bytes, P.x + (size_t)((i % tiles_per_pass) % 16) * P.slot_x_bytes, P.slot_x_bytes, &full[s]);
    }
  } else if (warp == 7 && (tid & 31) == 0) {  // consumer
    unsigned acc = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % P.nslots, round = i / P.nslots;
      ptx::mbar_wait(&full[s], round & 1);
      acc += *reinterpret_cast<volatile unsigned*>(smem + (size_t)s * slot_bytes);
      ptx::mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) printf("x");
  }
  __syncthreads();
  if (tid == 0) P.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256, 1) barrier_kernel(unsigned* counter, int iters, float* sink, int variant) {
  unsigned target = 0;
  for (int i = 0; i < iters; ++i) {
    if (variant == 1) sink[blockIdx.x * 256 + threadIdx.x] = (float)i;  // a store to flush before the release
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      if (variant == 1) __threadfence();
      red_release_gpu_add(counter, 1u);
      while (ld_acquire_gpu(counter) < target) {
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, 1) ldg_stream_kernel(const uint4* __restrict__ w, int slice_bytes, int passes, unsigned* out) {
  const uint4* base = w + (size_t)blockIdx.x * (slice_bytes / 16);
  const int n = slice_bytes / 16;
  unsigned acc = 0;
  for (int p = 0; p < passes; ++p) {
#pragma unroll 8
    for (int i = threadIdx.x; i < n; i += 256) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345u) out[0] = acc;
}

__global__ void dummy_cluster_kernel(int* p) {
  extern __shared__ uint8_t sm[];
  if (p) p[0] = sm[0];
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB, smem optin %zu\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20,
         prop.sharedMemPerBlockOptin);
  // ---- 1. cluster co-scheduling ----
  for (int cs : {1, 2, 4, 8}) {
    for (int smem : {100 * 1024, 200 * 1024, 225 * 1024}) {
      CK(cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(cs * 16);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_cluster_kernel, &cfg);
      printf("cluster %d, smem %3d KB, %d threads: max active clusters %d (%d CTAs) %s\n", cs, smem >> 10, kThreads, n, n * cs,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  // ---- 2. streaming ----
  const int maxgrid = 148;
  const int slice = 480 * 1024;
  uint8_t *w, *x;
  CK(cudaMalloc(&w, (size_t)maxgrid * slice));
  CK(cudaMemset(w, 1, (size_t)maxgrid * slice));
  CK(cudaMalloc(&x, 16 * 16 * 1024));
  CK(cudaMemset(x, 1, 16 * 16 * 1024));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, maxgrid * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Cfg { int grid, slot_w, slot_x, nslots, nprod, split; };
  const Cfg cfgs[] = {{128, 32768, 0, 4, 1, 1},  {128, 32768, 0, 4, 2, 1}, {128, 32768, 0, 4, 4, 1}, {128, 32768, 0, 4, 1, 4},
                      {128, 32768, 0, 4, 1, 8}, {128, 65536, 0, 3, 1, 1}, {128, 65536, 0, 2, 1, 1}, {128, 65536, 0, 3, 3, 4},
                      {128, 16384, 0, 8, 1, 1}, {128, 16384, 0, 8, 4, 1}, {128, 8192, 0, 16, 4, 1}, {128, 8192, 0, 8, 8, 1},
                      {128, 32768, 8192, 4, 2, 2}, {16, 32768, 0, 4, 1, 1}, {1, 32768, 0, 4, 1, 1}, {1, 32768, 0, 4, 4, 1}};
  for (const Cfg& c : cfgs) {
    StreamParams P;
    P.w = w; P.x = x; P.slice_bytes = slice; P.slot_w_bytes = c.slot_w; P.slot_x_bytes = c.slot_x; P.nslots = c.nslots;
    P.passes = 50; P.cycles = cyc; P.nprod = c.nprod; P.split = c.split; P.poll = 0;
    const int smem = (c.slot_w + c.slot_x) * c.nslots;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<c.grid, kThreads, smem>>>(P);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)c.grid * P.passes * (double)(slice / c.slot_w) * (c.slot_w + c.slot_x);
    printf("stream grid %3d nprod %d split %d slot %2d+%2d KB x%d: %.3f ms for %d passes -> %.2f us/pass, %.1f GB/s aggregate, %.1f GB/s per SM\n",
           c.grid, c.nprod, c.split, c.slot_w >> 10, c.slot_x >> 10, c.nslots, ms, P.passes, ms * 1e3 / P.passes, bytes / (ms * 1e-3) / 1e9,
           bytes / (ms * 1e-3) / 1e9 / c.grid);
  }
  for (int grid : {128, 148, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      ldg_stream_kernel<<<grid, 256>>>((const uint4*)w, slice, 50, (unsigned*)cyc);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("LDG.cg stream grid %3d: %.2f us/pass, %.1f GB/s per SM\n", grid, ms * 1e3 / 50, (double)slice * 50 / (ms * 1e-3) / 1e9);
  }
  // ---- 3. grid barrier ----
  unsigned* counter;
  float* sink;
  CK(cudaMalloc(&counter, 256));
  CK(cudaMalloc(&sink, 148 * 256 * sizeof(float)));
  for (int variant = 0; variant < 2; ++variant)
    for (int grid : {128, 148, 32}) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(counter, 0, 256));
        CK(cudaEventRecord(e0));
        barrier_kernel<<<grid, 256>>>(counter, iters, sink, variant);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("grid barrier (%s) grid %3d: %.3f us per barrier\n", variant ? "store+threadfence" : "bare", grid, ms * 1e3 / iters);
    }
  return 0;
}
