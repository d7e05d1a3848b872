// Hand-written tcgen05 GEMM for the WaveGlow WN contractions: bf16x3 (both operands split hi + lo, fp32 accumulation in TMEM)
// with the split done ON CHIP -- each operand tile is staged once and multiplied three times.
//
//   C[M, N] = A[M, K] . B[K, N]  ~  A_hi B_hi + A_lo B_hi + A_hi B_lo
//
// Operand images (bf16, K-major canonical layout without swizzle, the same format the decoder kernels stream), per 64-wide
// k-block a "chunk" holding the hi tile followed by the lo tile:
//   A: [M/128 m-tiles][K/64 k-blocks][hi | lo][128 rows x 64 k]  (2 x 16 KB per chunk)
//   B: [N/256 n-tiles][K/64 k-blocks][hi | lo][256 cols x 64 k]  (2 x 32 KB per chunk; "column" = output channel)
//   element (r, k) of a tile at byte (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2  =>  LBO = 128 (K), SBO = 1024 (M/N)
// so one pipeline stage is two contiguous bulk async copies (96 KB, no tensor map, no swizzle) feeding 12 MMAs, and the
// producers of the activations write their outputs directly in this layout (8 consecutive rows x 16 B = one 128-byte line).
// Compared with folding the three products into K ([hi|lo|hi] x [hi;hi;lo], the form a plain bf16 GEMM
// would need) this moves a third less operand data per FLOP and the producers write two copies of every value instead of three.
//
// Kernel: persistent, one CTA per SM, 320 threads.  warp 0 = copy producer, warp 1 = TMEM owner + MMA issuer (one elected
// thread, M = 128, N = 256, K = 16 per instruction), warps 2-9 = epilogue (two per TMEM lane quarter).  2-stage smem ring
// (96 KB / stage), two 256-column accumulators in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"
#include "scratch_pool.h"

constexpr int kGemmStages = 2;
constexpr uint32_t kGemmATile = 128 * 64 * 2, kGemmBTile = 256 * 64 * 2;                 // one hi or lo tile
constexpr uint32_t kGemmAChunk = 2 * kGemmATile, kGemmBChunk = 2 * kGemmBTile, kGemmStage = kGemmAChunk + kGemmBChunk;
constexpr int kGemmThreads = 320;  // producer, MMA, 8 epilogue warps (two per TMEM lane quarter, half the columns each)

__device__ __forceinline__ TcItem tc_item(const TcGemmParams& P, int i) {
  TcItem it;
  if (P.ksplit > 1) {  // general front-end: (tile, K slice)
    it.tile = i / P.ksplit; it.split = i - it.tile * P.ksplit; it.role = 0;
    it.kb0 = (int)((long long)it.split * P.Kb / P.ksplit); it.kb1 = (int)((long long)(it.split + 1) * P.Kb / P.ksplit);
  } else if (i < P.n_full) {
    it.tile = i; it.kb0 = 0; it.kb1 = P.Kb; it.role = 0; it.split = 0;
  } else {
    const int j = (i - P.n_full) >> 1, h = (i - P.n_full) & 1, kh = P.Kb >> 1;
    it.tile = P.n_full + j; it.split = j;
    it.kb0 = h ? kh : 0; it.kb1 = h ? P.Kb : kh; it.role = 1 + h;
  }
  const int per = P.Mt * P.Nt;
  it.b = it.tile / per;
  const int rem = it.tile - it.b * per;
  it.mt = rem / P.Nt;
  it.nt = (rem % P.Nt + it.mt) % P.Nt;  // rotated: every CTA gets a mix of n-tiles
  return it;
}

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1) tc_gemm_kernel(const TcGemmParams P) {
  extern __shared__ __align__(1024) uint8_t gsm[];
  uint8_t* ring = gsm;
  uint64_t* full = reinterpret_cast<uint64_t*>(gsm + kGemmStages * kGemmStage);
  uint64_t* empty = full + kGemmStages;
  uint64_t* acc_full = empty + kGemmStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kGemmStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nitems = P.ksplit > 1 ? P.n_full * P.ksplit : P.n_full + 2 * P.n_split;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      const uint64_t l2pol = P.stream_l2 ? ptx::l2_policy_evict_first() : 0ull;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const TcItem w = tc_item(P, item);
        const uint8_t* a = P.A + (size_t)w.b * P.a_bstride + (size_t)w.mt * P.Kb * kGemmAChunk;
        const uint8_t* b = P.B + (size_t)w.b * P.b_bstride + (size_t)w.nt * P.Kb * kGemmBChunk;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
          const int s = it % kGemmStages, round = it / kGemmStages;
          if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
          ptx::mbar_arrive_expect_tx(&full[s], kGemmStage);
          uint8_t* dst = ring + (size_t)s * kGemmStage;
          if (P.stream_l2) {
            ptx::bulk_g2s_hint(dst, a + (size_t)kb * kGemmAChunk, kGemmAChunk, &full[s], l2pol);
            ptx::bulk_g2s_hint(dst + kGemmAChunk, b + (size_t)kb * kGemmBChunk, kGemmBChunk, &full[s], l2pol);
          } else {
            ptx::bulk_g2s(dst, a + (size_t)kb * kGemmAChunk, kGemmAChunk, &full[s]);
            ptx::bulk_g2s(dst + kGemmAChunk, b + (size_t)kb * kGemmBChunk, kGemmBChunk, &full[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 256);
      int it = 0, ti = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ti) {
        const TcItem w = tc_item(P, item);
        const int ab = ti & 1, use = ti >> 1;
        if (use > 0) ptx::mbar_wait(&acc_empty[ab], (use - 1) & 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem + (uint32_t)ab * 256;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
          const int s = it % kGemmStages, round = it / kGemmStages;
          ptx::mbar_wait(&full[s], round & 1);
          ptx::tc_fence_after();
          const uint32_t a_hi = ptx::smem_u32(ring + (size_t)s * kGemmStage), a_lo = a_hi + kGemmATile;
          const uint32_t b_hi = a_hi + kGemmAChunk, b_lo = b_hi + kGemmBTile;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ah = ptx::umma_desc(a_hi + k * 256, 128, 1024), al = ptx::umma_desc(a_lo + k * 256, 128, 1024);
            const uint64_t bh = ptx::umma_desc(b_hi + k * 256, 128, 1024), bl = ptx::umma_desc(b_lo + k * 256, 128, 1024);
            ptx::umma_bf16(d, al, bh, idesc, (kb == w.kb0 && k == 0) ? 0u : 1u);  // small terms first
            ptx::umma_bf16(d, ah, bl, idesc, 1u);
            ptx::umma_bf16(d, ah, bh, idesc, 1u);
          }
          ptx::umma_commit(&empty[s]);
        }
        ptx::umma_commit(&acc_full[ab]);
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue warps: TMEM lane quarter q = warp & 3 holds rows 32q .. 32q+31 of the tile ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    int ti = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ti) {
      const TcItem w = tc_item(P, item);
      const int mt = w.mt, nt = w.nt;
      const int ab = ti & 1, use = ti >> 1;
      ptx::mbar_wait(&acc_full[ab], use & 1);
      ptx::tc_fence_after();
      const int row = mt * 128 + q * 32 + lane;
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * 256;
      if (w.role == 1) {
        // writer half: park the fp32 accumulator as [col][row] (coalesced over the lanes), then signal
        float* part = P.partial + (size_t)w.split * 256 * 128 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(ta + c0, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) part[(size_t)(c0 + j) * 128] = __uint_as_float(v[j]);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(P.flags + w.split, 1u);
      } else {
        const float* part = nullptr;
        if (w.role == 2) {
          if (lane == 0)
            while (ld_acquire_gpu(P.flags + w.split) < 8u) {
            }
          __syncwarp();
          part = P.partial + (size_t)w.split * 256 * 128 + q * 32 + lane;
        }
        tc_gemm_epilogue<MODE>(P, ta, row, nt, half, part, w);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[ab]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// ---- epilogues --------------------------------------------------------------------------------------------------
// MODE 0: plain fp32 store C[row, nt*256 + c]
template <>
__device__ __forceinline__ void tc_gemm_epilogue<0>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part, const TcItem&) {
#pragma unroll 1
  for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(ta + c0, v);
    ptx::tmem_wait_ld();
    if (part) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(part + (size_t)(c0 + j) * 128));
    }
    if (row < P.M) {
      float* dst = P.C + (size_t)row * P.ldc + nt * 256 + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
}

// MODE 1 (WN gate): columns [0,128) = tanh pre-activations of channels 128 nt + c, [128,256) = the matching sigmoid ones
template <>
__device__ __forceinline__ void tc_gemm_epilogue<1>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part, const TcItem&) {
  const bool valid = row < P.M;  // rows are compact: m = n * T + t
#pragma unroll 1
  for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 16) {
    uint32_t va[16], vs[16];
    ptx::tmem_ld16(ta + c0, va);
    ptx::tmem_ld16(ta + 128 + c0, vs);
    ptx::tmem_wait_ld();
    if (part) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        va[j] = __float_as_uint(__uint_as_float(va[j]) + __ldcg(part + (size_t)(c0 + j) * 128));
        vs[j] = __float_as_uint(__uint_as_float(vs[j]) + __ldcg(part + (size_t)(128 + c0 + j) * 128));
      }
    }
    if (valid) {
      const int ch0 = nt * 128 + c0;
      float g[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float at = __uint_as_float(va[j]) + P.bias0[ch0 + j] + P.bias1[ch0 + j];
        const float as = __uint_as_float(vs[j]) + P.bias0[512 + ch0 + j] + P.bias1[512 + ch0 + j];
        g[j] = tanhf(at) * (1.f / (1.f + expf(-as)));
      }
      const float g0[8] = {g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7]};
      const float g1[8] = {g[8], g[9], g[10], g[11], g[12], g[13], g[14], g[15]};
      wn_store_hl(P.out_img, row, kWnK2 / 64, ch0, g0);
      wn_store_hl(P.out_img, row, kWnK2 / 64, ch0 + 8, g1);
    }
  }
}

// MODE 2 (WN res/skip): layers < 7: n-tiles 0,1 = residual channels, 2,3 = skip channels; last layer: both tiles are skip.
// 32 columns per iteration; everything the chunk needs from memory (bias, the gated activation's hi/lo chunks or the old skip
// values) is requested before the TMEM load is waited for, so the three latencies overlap instead of chaining.
template <>
__device__ __forceinline__ void tc_gemm_epilogue<2>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part, const TcItem&) {
  const int t = row % P.T;
  const bool valid = row < P.M;
  const size_t prow = (size_t)(row / P.T) * P.Tp + 128 + t;  // row of the padded fp32 layout the flow epilogue reads
  const bool is_res = !P.lastl && nt < 2;
  const int chb = is_res ? nt * 256 : (P.lastl ? nt * 256 : (nt - 2) * 256);  // first channel of this tile
  const int bofs = is_res || P.lastl ? 0 : 512;                               // bias offset of the skip half
#pragma unroll 1
  for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(ta + c0, v);
    const int ch0 = chb + c0;
    float4 bias[8], old[8];
    uint4 ghi[4], glo[4];
    if (valid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[j] = __ldg(reinterpret_cast<const float4*>(P.bias0 + bofs + ch0) + j);
      if (is_res) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          ghi[c8] = *reinterpret_cast<const uint4*>(P.A + tc_a_chunk_offset(row, ch0 + 8 * c8, 0, kWnK2 / 64));
          glo[c8] = *reinterpret_cast<const uint4*>(P.A + tc_a_chunk_offset(row, ch0 + 8 * c8, 1, kWnK2 / 64));
        }
      } else if (!P.first) {
#pragma unroll
        for (int j = 0; j < 8; ++j) old[j] = *(reinterpret_cast<const float4*>(P.skip + prow * 512 + ch0) + j);
      }
    }
    ptx::tmem_wait_ld();
    if (part) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(part + (size_t)(c0 + j) * 128));
    }
    if (!valid) continue;
    float x[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bias[j].x;
      x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bias[j].y;
      x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bias[j].z;
      x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bias[j].w;
    }
    if (is_res) {
      // the gated activation this residual is added to: hi + lo from the A operand image (g to 16 mantissa bits, the same
      // value the tensor core multiplied)
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&ghi[c8]);
        const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&glo[c8]);
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = x[8 * c8 + j] + (__bfloat162float(hp[j]) + __bfloat162float(lp[j]));
        wn_store_taps(P.out_img, row, t, P.T, P.dil, ch0 + 8 * c8, h);
      }
    } else {
      float4* sp = reinterpret_cast<float4*>(P.skip + prow * 512 + ch0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        if (!P.first) o = make_float4(o.x + old[j].x, o.y + old[j].y, o.z + old[j].z, o.w + old[j].w);
        sp[j] = o;
      }
    }
  }
}

// MODE 3 (general front-end): C = acc + beta C over the valid M x N window, any leading dimension; or, for a K slice of a
// split product, a plain store into the slice's slab
template <>
__device__ __forceinline__ void tc_gemm_epilogue<3>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float*,
                                                    const TcItem& w) {
  const bool slab = P.ksplit > 1, valid = row < P.M;
  const int Np = P.Nt * 256;
  float* base = slab ? P.slabs + ((size_t)(w.split * P.batch + w.b) * P.Mt * 128 + row) * Np + nt * 256
                     : P.C + (size_t)w.b * P.c_bstride + (size_t)row * P.ldc + nt * 256;
  const int ncols = slab ? 256 : min(256, P.N - nt * 256);
  const bool vec = slab || ((P.ldc & 3) == 0 && (P.c_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
  const float beta = slab ? 0.f : P.beta;
  const bool stream = P.stream_l2 > 1;
#pragma unroll 1
  for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(ta + c0, v);
    ptx::tmem_wait_ld();
    if (!valid || c0 >= ncols) continue;
    float* dst = base + c0;
    if (vec && c0 + 32 <= ncols) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        if (beta != 0.f) {
          const float4 c = stream ? __ldcs(reinterpret_cast<const float4*>(dst + j)) : *reinterpret_cast<const float4*>(dst + j);
          o = make_float4(o.x + beta * c.x, o.y + beta * c.y, o.z + beta * c.z, o.w + beta * c.w);
        }
        if (stream)
          __stcs(reinterpret_cast<float4*>(dst + j), o);
        else
          *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < ncols) dst[j] = __uint_as_float(v[j]) + (beta != 0.f ? beta * dst[j] : 0.f);
    }
  }
}

static size_t tc_gemm_smem() { return (size_t)kGemmStages * kGemmStage + 256; }

template <int MODE>
static int tc_gemm_launch(const TcGemmParams& P0, cudaStream_t s, void* scratch) {
  TcGemmParams P = P0;
  const size_t smem = tc_gemm_smem();
  MSTTS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntiles = P.Mt * P.Nt, ncta = 148;
  // a last round with r <= 74 tiles is run as 2r half-K items (makespan +0.5 instead of +1 tile time).  Measured effect at
  // M=16000 N=1024 K=6528: 0.167 -> 0.164 ms -- the kernel is limited by the power-capped tensor clock, not by the schedule.
  const int r = ntiles % ncta;
  P.n_full = ntiles;
  P.n_split = 0;
  if (scratch && ntiles > ncta && r > 0 && 2 * r <= ncta && P.Kb >= 4) {
    P.n_full = ntiles - r;
    P.n_split = r;
    P.partial = (float*)scratch;
    P.flags = (unsigned*)((char*)scratch + (size_t)74 * 256 * 128 * 4);
    MSTTS_CUDA(cudaMemsetAsync(P.flags, 0, 74 * sizeof(unsigned), s));
  }
  const int nitems = P.n_full + 2 * P.n_split;
  tc_gemm_kernel<MODE><<<nitems < ncta ? nitems : ncta, kGemmThreads, smem, s>>>(P);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

int tc_gemm_plain(cudaStream_t s, const void* A_tiled, const void* B_tiled, float* C, int M, int N, int K, int ldc, void* scratch) {
  MSTTS_REQUIRE(N % 256 == 0 && K % 64 == 0 && M >= 1, MSTTS_E_INVALID, "tc_gemm: M=%d N=%d K=%d (N %% 256, K %% 64)", M, N, K);
  TcGemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = (const uint8_t*)A_tiled; P.B = (const uint8_t*)B_tiled; P.C = C; P.M = M; P.Mt = (M + 127) / 128; P.Nt = N / 256; P.Kb = K / 64;
  P.ldc = ldc;
  return tc_gemm_launch<0>(P, s, scratch);
}

int tc_gemm_wn_gate(cudaStream_t s, const void* A1, const void* B1, int M, const float* b_in, const float* b_cond, float* g_f32, void* A2,
                    int T, int Tp, void* scratch) {
  TcGemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = (const uint8_t*)A1; P.B = (const uint8_t*)B1; P.M = M; P.Mt = (M + 127) / 128; P.Nt = 4; P.Kb = kWnK1 / 64;
  P.bias0 = b_in; P.bias1 = b_cond; P.out_f32 = g_f32; P.out_img = (uint8_t*)A2; P.T = T; P.Tp = Tp;
  return tc_gemm_launch<1>(P, s, scratch);
}

int tc_gemm_wn_res(cudaStream_t s, const void* A2, const void* B2, int M, const float* b_res, const float* g_f32, float* skip, void* A1_next,
                   int T, int Tp, int dil_next, int first, int lastl, void* scratch) {
  TcGemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = (const uint8_t*)A2; P.B = (const uint8_t*)B2; P.M = M; P.Mt = (M + 127) / 128; P.Nt = lastl ? 2 : 4; P.Kb = kWnK2 / 64;
  P.bias0 = b_res; P.g_in = g_f32; P.skip = skip; P.out_img = (uint8_t*)A1_next; P.T = T; P.Tp = Tp; P.dil = dil_next; P.first = first;
  P.lastl = lastl;
  return tc_gemm_launch<2>(P, s, scratch);
}

// ---- operand tiling (tests, and the one-off conversion of weights): src row-major [R, K] fp32 (ld) -> tiled bf16 image
//      with TR rows per tile (128 for A, 256 for B given as [N, K], i.e. B transposed); rows >= R are zero ----
__global__ void tile_rows_kernel(const float* __restrict__ src, int R, int K, int ld, int TR, __nv_bfloat16* __restrict__ dst) {
  const int Rt = (R + TR - 1) / TR, Kb = K / 64;
  const size_t n = (size_t)Rt * TR * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const size_t r = i / K;
    const int rt = (int)(r / TR), rr = (int)(r % TR);
    const float x = r < (size_t)R ? src[r * ld + k] : 0.f;
    const size_t chunk = (size_t)rt * Kb + k / 64;  // [hi tile | lo tile]
    const int kk = k % 64;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const size_t e = chunk * 2 * TR * 64 + (size_t)(rr / 8) * 512 + (kk / 8) * 64 + (rr % 8) * 8 + (kk % 8);
    dst[e] = h;
    dst[e + (size_t)TR * 64] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

extern "C" int mstts_tc_gemm_test(const float* A, const float* Bt, int M, int N, int K, float* C, void* ws, size_t ws_bytes, void* stream) {
  // A [M,K] fp32, Bt [N,K] fp32 (B transposed) -> C [M,N] = A . Bt^T as bf16x3 (hi/lo split on chip) through the hand-written kernel
  const size_t Mt = (M + 127) / 128, Nt = N / 256;
  const size_t abytes = Mt * 128 * (size_t)K * 4, bbytes = Nt * 256 * (size_t)K * 4;  // hi + lo tiles
  MSTTS_REQUIRE(A && Bt && C && ws, MSTTS_E_INVALID, "tc_gemm_test: null pointer");
  MSTTS_REQUIRE(N % 256 == 0 && K % 64 == 0, MSTTS_E_INVALID, "tc_gemm_test: N %% 256, K %% 64");
  MSTTS_REQUIRE(ws_bytes >= abytes + bbytes + 2048 + kTcGemmScratchBytes, MSTTS_E_WORKSPACE, "tc_gemm_test: workspace %zu < %zu", ws_bytes,
                abytes + bbytes + 2048 + kTcGemmScratchBytes);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* a = (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint8_t* b = a + abytes;
  tile_rows_kernel<<<148 * 8, 256, 0, s>>>(A, M, K, K, 128, (__nv_bfloat16*)a);
  tile_rows_kernel<<<148 * 8, 256, 0, s>>>(Bt, N, K, K, 256, (__nv_bfloat16*)b);
  return tc_gemm_plain(s, a, b, C, M, N, K, N, b + bbytes + 256);
}

extern "C" int mstts_tc_gemm_tiled(const void* A_tiled, const void* B_tiled, int M, int N, int K, float* C, int ldc, void* scratch,
                                   void* stream) {
  MSTTS_REQUIRE(A_tiled && B_tiled && C, MSTTS_E_INVALID, "tc_gemm_tiled: null pointer");
  return tc_gemm_plain((cudaStream_t)stream, A_tiled, B_tiled, C, M, N, K, ldc, scratch);
}

// =====================================================================================================================
// General row-major front-ends: pack -> tc_gemm_kernel<3> -> (split-K reduce).  These carry every dense product outside
// the persistent loops (hoisted decoder products and weight gradients, the WaveGlow reverse pass, the mel up-sampling
// contraction, the encoder / postnet convolutions): the library links no vendor GEMM.
// =====================================================================================================================
struct PackSrc {
  const void* hi;   // fp32 matrix (F32) or bf16 hi part
  const void* lo;   // bf16 lo part (unused for F32)
  int ld;
  long long bstride;  // elements between batches
  int trans;          // 0: element (r, k) at [r * ld + k];  1: at [k * ld + r]
};

// One warp packs 256 elements per pass.  Not transposed: 8 rows x 32 k (a lane reads 8 consecutive k of one row, the four
// lanes of a row 128 contiguous bytes of fp32); transposed: 32 rows x 8 k (a lane reads 8 strided values, the warp 128
// contiguous bytes per k).  Either way a lane owns one 16-byte chunk of the hi tile and one of the lo tile, and 8 lanes
// with consecutive rows write one full 128-byte line.
template <bool F32>
__global__ void __launch_bounds__(256) pack_image_kernel(const PackSrc S, int R, int K, int TR, int Kb, int Kb_total,
                                                         uint8_t* __restrict__ dst, size_t dst_bstride, int batch, int vec,
                                                         int precise, int b_side, int stream_l2) {
  const int Rp = (R + TR - 1) / TR * TR, Kp = Kb * 64;
  const long long per = (long long)Rp * Kp / 256, total = per * batch;
  const int lane = threadIdx.x & 31;
  const size_t tile_bytes = (size_t)TR * 128;
  const int Kb_img = Kb_total * (precise ? 4 : 1);  // k-blocks per tile row of the image
  for (long long wu = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); wu < total;
       wu += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(wu / per);
    const long long u = wu - (long long)b * per;
    int r, k0;
    if (!S.trans) {
      const int nrb = Rp >> 3;
      r = (int)(u % nrb) * 8 + (lane & 7);
      k0 = (int)(u / nrb) * 32 + (lane >> 3) * 8;
    } else {
      const int nrb = Rp >> 5;
      r = (int)(u % nrb) * 32 + lane;
      k0 = (int)(u / nrb) * 8;
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8], lo2[8];
    if (F32) {
      const float* src = reinterpret_cast<const float*>(S.hi) + (long long)b * S.bstride;
      float x[8];
      if (!S.trans) {
        const float* p = src + (size_t)r * S.ld + k0;
        if (vec && r < R && k0 + 8 <= K) {
          float4 a, c;
          if (stream_l2) {
            a = __ldcs(reinterpret_cast<const float4*>(p));
            c = __ldcs(reinterpret_cast<const float4*>(p + 4));
          } else {
            a = *reinterpret_cast<const float4*>(p);
            c = *reinterpret_cast<const float4*>(p + 4);
          }
          x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = (r < R && k0 + j < K) ? p[j] : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float* q = src + (size_t)(k0 + j) * S.ld + r;
          x[j] = (r < R && k0 + j < K) ? (stream_l2 > 1 ? __ldcs(q) : *q) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = __float2bfloat16_rn(x[j]);
        const float r1 = x[j] - __bfloat162float(hi[j]);
        lo[j] = __float2bfloat16_rn(r1);
        lo2[j] = __float2bfloat16_rn(r1 - __bfloat162float(lo[j]));
      }
    } else {
      const __nv_bfloat16* sh = reinterpret_cast<const __nv_bfloat16*>(S.hi) + (long long)b * S.bstride;
      const __nv_bfloat16* sl = reinterpret_cast<const __nv_bfloat16*>(S.lo) + (long long)b * S.bstride;
      const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
      if (!S.trans) {
        const size_t o = (size_t)r * S.ld + k0;
        if (vec && r < R && k0 + 8 <= K) {
          *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(sh + o);
          *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(sl + o);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool ok = r < R && k0 + j < K;
            hi[j] = ok ? sh[o + j] : z;
            lo[j] = ok ? sl[o + j] : z;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = r < R && k0 + j < K;
          const size_t o = (size_t)(k0 + j) * S.ld + r;
          hi[j] = ok ? sh[o] : z;
          lo[j] = ok ? sl[o] : z;
        }
      }
    }
    uint8_t* d = dst + (size_t)b * dst_bstride + ((size_t)(r / TR) * Kb_img + (k0 >> 6)) * 2 * tile_bytes + (size_t)((r % TR) >> 3) * 1024 +
                 (size_t)((k0 & 63) >> 3) * 128 + (size_t)(r & 7) * 16;
    const uint4 h4 = *reinterpret_cast<const uint4*>(hi), m4 = *reinterpret_cast<const uint4*>(lo);
    if (stream_l2) {
      __stcs(reinterpret_cast<uint4*>(d), h4);
      __stcs(reinterpret_cast<uint4*>(d + tile_bytes), m4);
    } else {
      *reinterpret_cast<uint4*>(d) = h4;
      *reinterpret_cast<uint4*>(d + tile_bytes) = m4;
    }
    if (F32 && precise) {
      // segments 1-3 of the 3-way split (x = h + m + l, 24 mantissa bits), Kb_total k-blocks apart:
      //   seg 0: (h, m) x (h, m) -> hh + mh + hm     seg 1: (m, 0) x (m, 0) -> mm
      //   seg 2: A (h, 0) x B (l, 0) -> hl           seg 3: A (l, 0) x B (h, 0) -> lh
      const uint4 l4 = *reinterpret_cast<const uint4*>(lo2), z4 = make_uint4(0u, 0u, 0u, 0u);
      const size_t seg = (size_t)Kb_total * 2 * tile_bytes;
      *reinterpret_cast<uint4*>(d + seg) = m4;
      *reinterpret_cast<uint4*>(d + seg + tile_bytes) = z4;
      *reinterpret_cast<uint4*>(d + 2 * seg) = b_side ? l4 : h4;
      *reinterpret_cast<uint4*>(d + 2 * seg + tile_bytes) = z4;
      *reinterpret_cast<uint4*>(d + 3 * seg) = b_side ? h4 : l4;
      *reinterpret_cast<uint4*>(d + 3 * seg + tile_bytes) = z4;
    }
  }
}

// C_b[m, n] = beta C_b[m, n] + sum_s slab[s][b][m][n]   (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ slabs, int S, int batch, int M, int N, int Mp, int Np,
                                                            float* __restrict__ C, int ldc, size_t c_bstride, float beta, int stream_l2) {
  const size_t total = (size_t)batch * M * N, slab = (size_t)batch * Mp * Np;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const size_t bm = i / N;
    const int m = (int)(bm % M), b = (int)(bm / M);
    const float* p = slabs + ((size_t)b * Mp + m) * Np + n;
    double acc = 0.0;  // the slices are short tensor-core accumulation chains; their sum is exact to fp32
    float* c = C + (size_t)b * c_bstride + (size_t)m * ldc + n;
    if (stream_l2) {
      for (int s = 0; s < S; ++s) acc += (double)__ldcs(p + (size_t)s * slab);
      __stcs(c, beta != 0.f ? (float)(acc + (double)beta * (double)__ldcs(c)) : (float)acc);
    } else {
      for (int s = 0; s < S; ++s) acc += (double)p[(size_t)s * slab];
      *c = beta != 0.f ? (float)(acc + (double)beta * (double)*c) : (float)acc;
    }
  }
}

// K slices per tile: few-tile products with a long K (weight gradients over all decoder steps) would leave most SMs idle,
// and a tile count just above a multiple of the SM count wastes most of the last round.  Cost model in k-block times
// (12 MMAs, ~1 us): rounds x (slice length + pipeline fill) + the reduce pass.
static thread_local int g_grid_cap = 0;  // 0 = the whole device
TcGridCap::TcGridCap(int ctas) : prev(g_grid_cap) { g_grid_cap = ctas; }
TcGridCap::~TcGridCap() { g_grid_cap = prev; }
// launches under a grid cap run beside a persistent loop: stream their operands through L2 (MSTTS_SIDE_L2=0 switches it off)
static inline int side_stream_l2() {
  static const int on = [] {
    const char* e = getenv("MSTTS_SIDE_L2");
    return e ? atoi(e) : 2;
  }();
  return g_grid_cap > 0 ? on : 0;
}
static inline int gemm_ctas() { return g_grid_cap > 0 && g_grid_cap < 148 ? g_grid_cap : 148; }

static int choose_ksplit(int ntiles, int Kb, int kb_max) {
  const int ncta = gemm_ctas();
  // kb_max > 0 bounds the length of one tensor-core accumulation chain (k-blocks per slice): the fp32 accumulator in TMEM
  // truncates, so the error of a product grows linearly with the chain length (measured: 2.3e-5 of max at K = 3000, ten
  // times an fp32 SGEMM); short slices summed in double by the reduce kernel bring it back to fp32-SGEMM level
  const int s_min = kb_max > 0 ? (Kb + kb_max - 1) / kb_max : 1;
  int s_hi = Kb / 2 < ncta ? Kb / 2 : ncta;
  if (s_hi < s_min) s_hi = s_min;
  if (s_hi < 1) s_hi = 1;
  int best = s_min;
  double best_cost = 1e30;
  for (int S = s_min; S <= s_hi; ++S) {
    const long long rounds = ((long long)ntiles * S + ncta - 1) / ncta;
    double cost = (double)rounds * ((double)Kb / S + 2.0);
    if (S > 1) cost += 3.0 + 0.03 * S * ntiles;
    if (cost < best_cost * 0.97) {  // prefer fewer slices unless the gain is real
      best_cost = cost;
      best = S;
    }
  }
  return best;
}

static inline bool pack_vec_ok(const PackSrc& P, bool f32) {
  const size_t al = f32 ? 4 : 8;  // elements per 16 bytes
  return !P.trans && P.ld % al == 0 && P.bstride % (long long)al == 0 && ((uintptr_t)P.hi & 15) == 0 && (f32 || ((uintptr_t)P.lo & 15) == 0);
}

// pack `nb` matrices of R rows x K into images of Kb_total source k-blocks (x 4 segments when precise), starting at tile
// row rt0 / k-block kb0 of each image
template <bool F32>
static int pack_launch(cudaStream_t s, const PackSrc& P, int R, int K, int TR, int Kb_total, uint8_t* img, size_t img_bstride, int nb, int rt0,
                       int kb0, bool precise, bool b_side) {
  if (R <= 0 || K <= 0 || nb <= 0) return MSTTS_OK;
  const int Kb = (K + 63) / 64;
  const long long wu = (long long)((R + TR - 1) / TR * TR) * Kb * 64 / 256 * nb;
  long long g = (wu + 7) / 8;
  static const int pack_mult = [] {
    const char* e = getenv("MSTTS_OVERLAP_PACK_MULT");
    const int v = e && *e ? atoi(e) : 6;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  const long long gmax = g_grid_cap > 0 ? (long long)gemm_ctas() * pack_mult : 148 * 16;
  if (g > gmax) g = gmax;
  // the kernel addresses chunks relative to an image whose tile rows hold Kb_img k-blocks: shift the base to (rt0, kb0)
  const int Kb_img = Kb_total * (precise ? 4 : 1);
  uint8_t* base = img + ((size_t)rt0 * Kb_img + kb0) * 2 * (size_t)TR * 128;
  pack_image_kernel<F32><<<(int)g, 256, 0, s>>>(P, R, K, TR, Kb, Kb_total, base, img_bstride, nb, pack_vec_ok(P, F32) ? 1 : 0, precise ? 1 : 0,
                                                b_side ? 1 : 0, side_stream_l2());
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

int tc_pack_f32(cudaStream_t s, const float* src, int ld, bool trans, int R, int K, int TR, int Kb_total, void* img, int rt0, int kb0,
                bool precise) {
  MSTTS_REQUIRE(src && img && (TR == 128 || TR == 256) && kb0 + (K + 63) / 64 <= Kb_total, MSTTS_E_INVALID,
                "tc_pack_f32: R=%d K=%d TR=%d kb0=%d Kb=%d", R, K, TR, kb0, Kb_total);
  const PackSrc P{src, nullptr, ld, 0, trans ? 1 : 0};
  return pack_launch<true>(s, P, R, K, TR, Kb_total, (uint8_t*)img, 0, 1, rt0, kb0, precise, TR == 256);
}

int tc_pack_hl(cudaStream_t s, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, bool trans, int R, int K, int TR, int Kb_total, void* img,
               int rt0, int kb0) {
  MSTTS_REQUIRE(hi && lo && img && (TR == 128 || TR == 256) && kb0 + (K + 63) / 64 <= Kb_total, MSTTS_E_INVALID,
                "tc_pack_hl: R=%d K=%d TR=%d kb0=%d Kb=%d", R, K, TR, kb0, Kb_total);
  const PackSrc P{hi, lo, ld, 0, trans ? 1 : 0};
  return pack_launch<false>(s, P, R, K, TR, Kb_total, (uint8_t*)img, 0, 1, rt0, kb0, false, TR == 256);
}

// the product over packed images: C_b = A_b . B_b^T (+ beta C_b), Kb image k-blocks
static int launch_images(cudaStream_t s, const uint8_t* ai, size_t a_bstride, const uint8_t* bi, size_t b_bstride, int M, int N, int Kb,
                         float* C, int ldc, long long sC, float beta, int batch, int prec) {
  const int Mt = (M + 127) / 128, Nt = (N + 255) / 256;
  const int ntiles = batch * Mt * Nt;
  const int ksplit = choose_ksplit(ntiles, Kb, prec == TC_PRECISE ? 4 : (prec == TC_CHAINED ? 8 : 0));
  ScratchScope scope(s);
  float* slabs = nullptr;
  int rc;
  if (ksplit > 1 && (rc = scope.get((void**)&slabs, (size_t)ksplit * batch * Mt * 128 * Nt * 256 * sizeof(float)))) return rc;
  TcGemmParams P;
  memset(&P, 0, sizeof(P));
  P.A = ai; P.B = bi; P.C = C; P.M = M; P.N = N; P.Mt = Mt; P.Nt = Nt; P.Kb = Kb; P.ldc = ldc;
  P.batch = batch; P.ksplit = ksplit; P.a_bstride = a_bstride; P.b_bstride = b_bstride; P.c_bstride = (size_t)sC; P.beta = beta;
  P.slabs = slabs; P.n_full = ntiles; P.stream_l2 = side_stream_l2();
  const size_t smem = tc_gemm_smem();
  MSTTS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nitems = (long long)ntiles * ksplit;
  tc_gemm_kernel<3><<<(int)(nitems < gemm_ctas() ? nitems : gemm_ctas()), kGemmThreads, smem, s>>>(P);
  MSTTS_CUDA(cudaGetLastError());
  if (ksplit > 1) {
    size_t g = ((size_t)batch * M * N + 255) / 256;
    const size_t gmax = g_grid_cap > 0 ? (size_t)gemm_ctas() * 6 : 148 * 8;
    if (g > gmax) g = gmax;
    splitk_reduce_kernel<<<(int)g, 256, 0, s>>>(slabs, ksplit, batch, M, N, Mt * 128, Nt * 256, C, ldc, (size_t)sC, beta, side_stream_l2() > 1);
    MSTTS_CUDA(cudaGetLastError());
  }
  return MSTTS_OK;
}

int tc_gemm_images(cudaStream_t s, const void* A_img, const void* B_img, int M, int N, int K, float* C, int ldc, float beta, int prec) {
  if (M <= 0 || N <= 0) return MSTTS_OK;
  MSTTS_REQUIRE(A_img && B_img && C && K >= 1, MSTTS_E_INVALID, "tc_gemm_images: null operand or K=%d", K);
  return launch_images(s, (const uint8_t*)A_img, 0, (const uint8_t*)B_img, 0, M, N, (K + 63) / 64 * (prec == TC_PRECISE ? 4 : 1), C, ldc, 0,
                       beta, 1, prec);
}

template <bool F32>
static int tc_gemm_general(cudaStream_t s, PackSrc A, PackSrc B, int M, int N, int K, float* C, int ldc, long long sC, float beta, int batch,
                           int prec) {
  const bool precise = F32 && prec == TC_PRECISE;
  if (M <= 0 || N <= 0 || batch <= 0) return MSTTS_OK;
  MSTTS_REQUIRE(K >= 1 && A.hi && B.hi && C && (F32 || (A.lo && B.lo)), MSTTS_E_INVALID, "tc_gemm: M=%d N=%d K=%d batch=%d or null operand",
                M, N, K, batch);
  const int Mt = (M + 127) / 128, Nt = (N + 255) / 256, Kb = (K + 63) / 64, segs = precise ? 4 : 1;
  const bool shareB = batch > 1 && B.bstride == 0;
  const size_t a_img = (size_t)Mt * Kb * segs * kGemmAChunk, b_img = (size_t)Nt * Kb * segs * kGemmBChunk;
  ScratchScope scope(s);
  uint8_t *ai = nullptr, *bi = nullptr;
  int rc;
  if ((rc = scope.get((void**)&ai, a_img * batch))) return rc;
  if ((rc = scope.get((void**)&bi, b_img * (shareB ? 1 : batch)))) return rc;
  if ((rc = pack_launch<F32>(s, A, M, K, 128, Kb, ai, a_img, batch, 0, 0, precise, false))) return rc;
  if ((rc = pack_launch<F32>(s, B, N, K, 256, Kb, bi, b_img, shareB ? 1 : batch, 0, 0, precise, true))) return rc;
  return launch_images(s, ai, a_img, bi, shareB ? 0 : b_img, M, N, Kb * segs, C, ldc, sC, beta, batch, prec);
}

// Debug aid (tools/grad_probe.py): MSTTS_GEMM_FORCE=<level> overrides the precision level of every fp32 front-end call, or
// only of the calls whose running index (since process start) is listed in MSTTS_GEMM_FORCE_SITES="3,7,..."; MSTTS_GEMM_TRACE=1
// prints index and shape of each call.  Read per call; unset in production.
static int gemm_site_prec(int prec, bool tA, bool tB, int M, int N, int K, int batch) {
  static std::atomic<int> ctr{0};
  const char* force = getenv("MSTTS_GEMM_FORCE");
  const char* trace = getenv("MSTTS_GEMM_TRACE");
  if (!force && !trace) return prec;
  const int id = ctr++;
  int out = prec;
  if (force) {
    const char* sites = getenv("MSTTS_GEMM_FORCE_SITES");
    bool hit = sites == nullptr;
    for (const char* p = sites; p && *p;) {
      char* e;
      const long v = strtol(p, &e, 10);
      if (e == p) break;
      if (v == id) hit = true;
      p = *e ? e + 1 : e;
    }
    if (hit) out = atoi(force);
  }
  if (trace) fprintf(stderr, "[mstts gemm %d] tA=%d tB=%d M=%d N=%d K=%d batch=%d prec %d -> %d\n", id, (int)tA, (int)tB, M, N, K, batch, prec, out);
  return out;
}

int tc_gemm_f32(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const float* A, int lda, long long sA, const float* B,
                int ldb, long long sB, float* C, int ldc, long long sC, float beta, int batch, int prec) {
  prec = gemm_site_prec(prec, transA, transB, M, N, K, batch);
  // image rows of the B operand are output columns: element (n, k) = B[k][n] unless B is given transposed
  const PackSrc a{A, nullptr, lda, sA, transA ? 1 : 0}, b{B, nullptr, ldb, sB, transB ? 0 : 1};
  return tc_gemm_general<true>(s, a, b, M, N, K, C, ldc, sC, beta, batch, prec);
}

int tc_gemm_hl(cudaStream_t s, bool transA, bool transB, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
               int lda, long long sA, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, long long sB, float* C, int ldc,
               long long sC, float beta, int batch) {
  const PackSrc a{A_hi, A_lo, lda, sA, transA ? 1 : 0}, b{B_hi, B_lo, ldb, sB, transB ? 0 : 1};
  return tc_gemm_general<false>(s, a, b, M, N, K, C, ldc, sC, beta, batch, TC_FAST);
}

extern "C" int mstts_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, long long strideA, const float* B,
                              int ldb, long long strideB, float* C, int ldc, long long strideC, float beta, int batch, int precise,
                              void* stream) {
  MSTTS_REQUIRE(precise >= 0 && precise <= 2, MSTTS_E_INVALID, "gemm_f32: precision level %d (0 fast, 1 chained, 2 precise)", precise);
  return tc_gemm_f32((cudaStream_t)stream, transA != 0, transB != 0, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, beta, batch,
                     precise);
}
