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
// Compared with folding the three products into K ([hi|lo|hi] x [hi;hi;lo], the library-GEMM form used elsewhere in this
// repository) this moves a third less operand data per FLOP and the producers write two copies of every value instead of three.
//
// Kernel: persistent, one CTA per SM, 320 threads.  warp 0 = copy producer, warp 1 = TMEM owner + MMA issuer (one elected
// thread, M = 128, N = 256, K = 16 per instruction), warps 2-9 = epilogue (two per TMEM lane quarter).  2-stage smem ring
// (96 KB / stage), two 256-column accumulators in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

constexpr int kGemmStages = 2;
constexpr uint32_t kGemmATile = 128 * 64 * 2, kGemmBTile = 256 * 64 * 2;                 // one hi or lo tile
constexpr uint32_t kGemmAChunk = 2 * kGemmATile, kGemmBChunk = 2 * kGemmBTile, kGemmStage = kGemmAChunk + kGemmBChunk;
constexpr int kGemmThreads = 320;  // producer, MMA, 8 epilogue warps (two per TMEM lane quarter, half the columns each)

// work item i -> (tile, K-block range, role): role 0 = whole tile, 1 = writer (first K half), 2 = finisher (second half)
struct TcItem {
  int tile, kb0, kb1, role, split;
};
__device__ __forceinline__ TcItem tc_item(const TcGemmParams& P, int i) {
  TcItem it;
  if (i < P.n_full) {
    it.tile = i; it.kb0 = 0; it.kb1 = P.Kb; it.role = 0; it.split = 0;
  } else {
    const int j = (i - P.n_full) >> 1, h = (i - P.n_full) & 1, kh = P.Kb >> 1;
    it.tile = P.n_full + j; it.split = j;
    it.kb0 = h ? kh : 0; it.kb1 = h ? P.Kb : kh; it.role = 1 + h;
  }
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
  const int nitems = P.n_full + 2 * P.n_split;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const TcItem w = tc_item(P, item);
        const int mt = w.tile / P.Nt, nt = (w.tile % P.Nt + mt) % P.Nt;  // rotated: every CTA gets a mix of n-tiles
        const uint8_t* a = P.A + (size_t)mt * P.Kb * kGemmAChunk;
        const uint8_t* b = P.B + (size_t)nt * P.Kb * kGemmBChunk;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
          const int s = it % kGemmStages, round = it / kGemmStages;
          if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
          ptx::mbar_arrive_expect_tx(&full[s], kGemmStage);
          uint8_t* dst = ring + (size_t)s * kGemmStage;
          ptx::bulk_g2s(dst, a + (size_t)kb * kGemmAChunk, kGemmAChunk, &full[s]);
          ptx::bulk_g2s(dst + kGemmAChunk, b + (size_t)kb * kGemmBChunk, kGemmBChunk, &full[s]);
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
      const int mt = w.tile / P.Nt, nt = (w.tile % P.Nt + mt) % P.Nt;
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
        tc_gemm_epilogue<MODE>(P, ta, row, nt, half, part);
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
__device__ __forceinline__ void tc_gemm_epilogue<0>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part) {
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
__device__ __forceinline__ void tc_gemm_epilogue<1>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part) {
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
__device__ __forceinline__ void tc_gemm_epilogue<2>(const TcGemmParams& P, uint32_t ta, int row, int nt, int half, const float* part) {
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
