// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32) for the skinny shapes the decoder uses.
// Operands are static garbage in shared memory / TMEM: only the time per MMA matters.
//   SS  : A and B from shared memory (no-swizzle K-major canonical layout)
//   TS  : A from TMEM, B from shared memory
//   CP  : tcgen05.cp 128x256b shared -> TMEM (one K=16 slice of a 128-row A tile)
// Build: make probes.  Run: build/mma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../multi_speaker_tts_b200/csrc/sm100_ptx.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

struct Params {
  int M, N, mode, elect, nw;  // mode 0 = SS same accumulator, 1 = SS rotating 3 accumulators, 2 = TS, 3 = CP only, 4 = CP + 3 TS MMAs
  int iters;
  unsigned long long* cycles;
};

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void utccp_128x256b(uint32_t tmem_dst, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];  // A: 64 KB, B: 64 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(&tmem_slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (P.nw > 0) {
    if (warp < P.nw && (tid & 31) == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(P.M, P.N);
      const uint32_t abase = ptx::smem_u32(smem), bbase = abase + 64 * 1024;
      long long t0 = clock64();
      for (int i = 0; i < P.iters; ++i) {
        const uint64_t da = ptx::umma_desc(abase + (uint32_t)(i & 15) * 256, 128, 4096);
        const uint64_t db = ptx::umma_desc(bbase + (uint32_t)(i & 7) * 256, 128, 2048);
        ptx::umma_bf16(tmem + warp * P.N, da, db, idesc, 1u);
      }
      __shared__ __align__(8) uint64_t bars[4];
      ptx::mbar_init(&bars[warp], 1);
      ptx::fence_mbar_init();
      ptx::umma_commit(&bars[warp]);
      ptx::mbar_wait(&bars[warp], 0);
      if (warp == 0) P.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
  } else if (P.elect ? (warp == 0) : (tid == 0)) {
    const bool leader = P.elect ? ptx::elect_one() : true;
    const uint32_t idesc = ptx::umma_idesc_bf16(P.M, P.N);
    const uint32_t abase = ptx::smem_u32(smem), bbase = abase + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < P.iters; ++i) {
      const uint32_t koff = (uint32_t)(i & 15) * 256;  // walk 16 K-steps of a 64 KB operand region
      const uint64_t da = ptx::umma_desc(abase + koff, 128, 4096);
      const uint64_t db = ptx::umma_desc(bbase + (uint32_t)(i & 7) * 256, 128, 2048);  // 256 rows x K=128 -> 64 KB
      if (!leader) continue;
      if (P.mode == 0) {
        ptx::umma_bf16(tmem, da, db, idesc, 1u);
      } else if (P.mode == 1) {
        ptx::umma_bf16(tmem + (uint32_t)(i % 3) * P.N, da, db, idesc, 1u);
      } else if (P.mode == 2) {
        umma_ts(tmem, tmem + 256 + (uint32_t)(i & 7) * 8, db, idesc, 1u);
      } else if (P.mode == 3) {
        utccp_128x256b(tmem + 256 + (uint32_t)(i & 7) * 8, da);
      } else {
        utccp_128x256b(tmem + 256 + (uint32_t)(i & 7) * 8, da);
        utccp_128x256b(tmem + 384 + (uint32_t)(i & 7) * 8, da);
        umma_ts(tmem, tmem + 256 + (uint32_t)(i & 7) * 8, db, idesc, 1u);
        umma_ts(tmem, tmem + 256 + (uint32_t)(i & 7) * 8, db, idesc, 1u);
        umma_ts(tmem, tmem + 384 + (uint32_t)(i & 7) * 8, db, idesc, 1u);
      }
    }
    if (leader) ptx::umma_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    if (leader) P.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  const bool only_multi = argc > 1;  // any argument: only the multi-issuer part
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  const int iters = 4096;
  const char* names[] = {"SS same acc", "SS 3 accs", "TS (A in TMEM)", "CP 128x256b only", "2 CP + 3 TS (one bf16x3 K-step)"};
  for (int elect : {0, 1}) {
  if (only_multi) break;
  for (int grid : {128}) {
    for (int mode = 0; mode < 5; ++mode) {
      for (int M : {128, 64}) {
        for (int N : {32, 64, 128, 256}) {
          if (M == 64 && N != 32) continue;
          if (mode >= 2 && M == 64) continue;
          if (mode >= 3 && N != 32) continue;
          if (mode == 1 && N > 128) continue;
          Params P{M, N, mode, elect, 0, iters, cyc};
          mma_kernel<<<grid, 128, 128 * 1024>>>(P);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("grid %d mode %d M %d N %d: CUDA error %s\n", grid, mode, M, N, cudaGetErrorString(e));
            return 1;
          }
          unsigned long long h[148];
          CK(cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
          double mean = 0;
          for (int i = 0; i < grid; ++i) mean += (double)h[i];
          mean /= grid;
          printf("elect %d grid %3d  %-32s M=%3d N=%3d : %.1f cycles per iteration\n", elect, grid, names[mode], M, N, mean / iters);
        }
      }
    }
  }
  }
  for (int nw : {1, 2, 4})
    for (int N : {32, 64, 128}) {  // N = 64: the swapped-role shape (A = 128 weight rows, B = 32 batch rows x hi/lo)
      Params P{128, N, 0, 0, nw, iters, cyc};
      mma_kernel<<<128, 128, 128 * 1024>>>(P);
      CK(cudaDeviceSynchronize());
      unsigned long long h[148];
      CK(cudaMemcpy(h, cyc, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      printf("%d issuing warps, N=%d: %.1f cycles per MMA per warp -> %.1f cycles per MMA aggregate\n", nw, N, (double)h[0] / iters, (double)h[0] / iters / nw);
    }
  // the decoder's shape (M=64 stacked hi/lo batch rows, N=256 stacked hi/lo gate rows): does a second issuing warp help?
  for (int nw : {1, 2}) {
    Params P{64, 256, 0, 0, nw, iters, cyc};
    mma_kernel<<<128, 128, 128 * 1024>>>(P);
    CK(cudaDeviceSynchronize());
    unsigned long long h[148];
    CK(cudaMemcpy(h, cyc, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("M=64 N=256, %d issuing warps (own accumulators): %.1f cycles per MMA per warp -> %.1f aggregate\n", nw, (double)h[0] / iters,
           (double)h[0] / iters / nw);
  }
  return 0;
}
