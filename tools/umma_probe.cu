// Standalone probe: one tcgen05.mma tile (M=128, N=32, K=64, bf16 -> fp32 in TMEM) with operands written by
// threads in the no-swizzle K-major canonical layout.  Checks which (LBO,SBO) assignment the hardware uses
// and that the TMEM read-back mapping (lane = row, column = n) holds.  Build: make probe; run on the GPU box.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

constexpr int M = 128, N = 32, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;         // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm,
                                                       float* __restrict__ Dout, int variant) {
  __shared__ __align__(1024) uint8_t sA[M * K * 2];
  __shared__ __align__(1024) uint8_t sB[N * K * 2];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // canonical no-swizzle K-major: core matrix = 8 rows x 16 B, contiguous 128 B.
  // offset(r,k) = (r/8)*MDIR + (k/8)*KDIR + (r%8)*16 + (k%8)*2, with KDIR = 128, MDIR = (K/8)*128
  const uint32_t KDIR = 128, MDIR = (K / 8) * 128;
  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = (r / 8) * MDIR + (k / 8) * KDIR + (r % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = (r / 8) * MDIR + (k / 8) * KDIR + (r % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = Bm[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t lbo = variant == 0 ? KDIR : MDIR;
    const uint32_t sbo = variant == 0 ? MDIR : KDIR;
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = make_desc(smem_u32(sA) + k * 2 * KDIR, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + k * 2 * KDIR, lbo, sbo);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int n = 0; n < N; ++n) Dout[tid * N + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
  __nv_bfloat16 *hA = (__nv_bfloat16*)malloc(M * K * 2), *hB = (__nv_bfloat16*)malloc(N * K * 2);
  float* ref = (float*)calloc(M * N, 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((float)(rand() % 7 - 3));
  for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16((float)(rand() % 5 - 2));
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += __bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
      ref[m * N + n] = s;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, M * K * 2);
  cudaMalloc(&dB, N * K * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
  float* out = (float*)malloc(M * N * 4);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0xff, M * N * 4);
    probe_kernel<<<1, 128>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(out, dD, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int i = 0; i < M * N; ++i) {
      double d = fabs((double)out[i] - ref[i]);
      if (!(d <= 1e-3)) ++bad;
      if (d > maxerr) maxerr = d;
    }
    printf("variant %d (LBO=%s): mismatches %d / %d, max err %g ; D[0][0..3] = %g %g %g %g  ref %g %g %g %g\n", variant,
           variant == 0 ? "K-dir,SBO=MN-dir" : "MN-dir,SBO=K-dir", bad, M * N, maxerr, out[0], out[1], out[2], out[3],
           ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
