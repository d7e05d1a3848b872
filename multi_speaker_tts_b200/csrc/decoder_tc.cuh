// Shared pieces of the tcgen05 ("bf16x3") persistent decoder kernels.
//
// Numerics: every recurrent skinny GEMM  Y[B,rows] = X[B,K] . W[rows,K]^T  runs on the 5th-gen tensor cores
// with both operands split into bf16 hi + bf16 lo (x = hi + lo to ~16 mantissa bits), fp32 accumulation in
// TMEM: measured L_inf vs the fp32 oracle ~2e-6 on the mel outputs (plain bf16 would be ~1e-3, i.e. at the
// parity gate).  A skinny tcgen05.mma is bound by its issue interval (tools/mma_probe.cu: 78-128 cycles), so the
// hi/lo halves are STACKED into one instruction per K-step: A = [X_hi ; X_lo] (M = 64 rows), B = [W_hi ; W_lo]
// (N = 256 rows) gives all four partial products in one 64 x 256 accumulator; the epilogue adds the quadrants.
//
// Data movement: weights do not fit on chip (hi+lo = 4 B/weight, 63 MB), so each CTA re-streams its slice from
// L2 every step through a ring of shared-memory slots filled by 1-D bulk async copies (TMA engine); the
// activations of the step are gathered the same way from small global "images".  Both images are stored in
// global memory exactly as the tensor core wants them in shared memory (K-major canonical layout, no
// swizzle), so a slot is one contiguous 32 KB (weights) + 8 KB (activations) copy.
//
// Tile image formats (bf16):
//   W tile  [128 rows x 64 k] : hi at +0 (16 KB), lo at +16 KB
//   X tile  [ 64 rows x 64 k] : 8-row groups alternate hi / lo: rows 16g..16g+7 = x_hi of batch 8g..8g+7, rows
//                               16g+8..16g+15 = x_lo of the same batches (zero beyond B).  With M = 64 the tensor core
//                               puts accumulator rows 16g..16g+15 on TMEM lanes 32g..32g+15, so the hi and lo partial
//                               sums of a batch row sit in the same warp's lanes (l, l+8) and combine with one shuffle.
//   element (row, k) at byte (row/8)*1024 + (k/8)*128 + (row%8)*16 + (k%8)*2
//   => UMMA descriptor: LBO (K direction) = 128, SBO (M/N direction) = 1024; one K=16 MMA step = +256 B;
//      hi and lo are adjacent, so a W tile doubles as a 256-row B operand and an X tile as a 64-row A operand.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

constexpr int kTcCompute = 256;           // warps 0..7: epilogues + attention
constexpr int kTcThreads = 384;           // + warps 8,9 (weight producers), 10 (activation producer), 11 (TMEM alloc + MMA issuer)
constexpr int kTcMmaWarp = 11;
constexpr int kTcKT = 64;                 // K per ring slot
constexpr int kTcN = 32;                  // MMA N = padded batch
constexpr uint32_t kWTileBytes = 128 * kTcKT * 2 * 2;  // 32768
constexpr uint32_t kXTileBytes = kTcN * kTcKT * 2 * 2;  // 8192
constexpr uint32_t kSlotBytes = kWTileBytes + kXTileBytes;
constexpr uint32_t kTcLBO = 128, kTcSBO = 1024;
constexpr int kRecvStride = 40;           // floats per (src, batch) row of the K-split reduction buffer [src][batch][gate*8+unit]

// Saved activations and gate gradients pass through L2 exactly once per loop (written by one loop, read by the other or by the
// weight-gradient packs, ~1 MB per step each); marked evict-first they do not displace the weight image every step re-reads.
__device__ __forceinline__ void st_stream(float* p, float v, int on) {
  if (on)
    __stcs(p, v);
  else
    *p = v;
}
__device__ __forceinline__ float ld_stream(const float* p, int on) { return on ? __ldcs(p) : *p; }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// byte offset of batch row b's hi (hl = 0) / lo (hl = 1) 16-byte row inside an X tile (add (k/8)*128 for the k chunk)
__device__ __forceinline__ size_t ximg_row_offset(int b, int hl) { return (size_t)(2 * (b >> 3) + hl) * 1024 + (size_t)(b & 7) * 16; }

__device__ __forceinline__ unsigned ld_volatile_shared(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(unsigned* p, unsigned v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ptx::smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// Barrier among the compute warps of all CTAs of a 4-CTA cluster, built on one mbarrier per CTA (count =
// 4 CTAs x 8 warps).  Orders the callers' earlier st.shared::cluster pushes before the peers' later reads.
// (barrier.cluster cannot be used here: it would also wait for the producer / MMA warps.)
__device__ __forceinline__ void cluster_compute_sync(uint64_t* bar, uint32_t& parity) {
  fence_cluster();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const uint32_t a = ptx::smem_u32(bar);
#pragma unroll
    for (uint32_t dst = 0; dst < (uint32_t)kDecCluster; ++dst) ptx::mbar_arrive_cluster(ptx::mapa(a, dst));
  }
  while (!ptx::mbar_try_wait_cluster(bar, parity)) {
  }
  parity ^= 1u;
}

// gate backward of one (batch, unit): ZoneoutLSTMCell.py:237-260 differentiated (same math as decoder_bwd.cu)
struct CellGradTc {
  float di, dj, df, dop, dc_prev, dh_prev;
};
__device__ __forceinline__ CellGradTc cell_backward_tc(float dm_direct, float dhz, float dcz, float ig, float jg, float fg,
                                                       float og, float cn, float cp, float mc, float mh) {
  CellGradTc r;
  const float dm = dm_direct + kZoneKeep * mh * dhz;
  r.dh_prev = dhz * (1.f - kZoneKeep * mh);
  float dc = kZoneKeep * mc * dcz;
  r.dc_prev = dcz * (1.f - kZoneKeep * mc);
  const float tc = tanhf(cn);
  const float d_o = dm * tc;
  dc += dm * og * (1.f - tc * tc);
  const float d_f = dc * cp;
  r.dc_prev += dc * fg;
  const float d_i = dc * jg, d_j = dc * ig;
  r.di = d_i * ig * (1.f - ig);
  r.dj = d_j * (1.f - jg * jg);
  r.df = d_f * fg * (1.f - fg);
  r.dop = d_o * og * (1.f - og);
  return r;
}

// sum over the 32 lanes of 16 independent per-lane values with a transposing butterfly (31 shuffles instead of 80);
// on return lane L (even lanes only) holds the total of value index ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float send = (lane & 16) ? v[p] : v[p + 8];
    const float keep = (lane & 16) ? v[p + 8] : v[p];
    v[p] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float send = (lane & 8) ? v[p] : v[p + 4];
    const float keep = (lane & 8) ? v[p + 4] : v[p];
    v[p] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float send = (lane & 4) ? v[p] : v[p + 2];
    const float keep = (lane & 4) ? v[p + 2] : v[p];
    v[p] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v[0] : v[1];
    const float keep = (lane & 2) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int warp_sum16_index(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// One lane per warp polls the mbarrier, the rest of the warp parks in __syncwarp: 256 threads spinning on
// try_wait would compete with the producer / MMA threads for the shared-memory pipeline.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) ptx::mbar_wait(bar, parity);
  __syncwarp();
}

// Grid-wide barrier executed by the compute warps only (named barrier 1), split into arrive and wait so that work
// nobody else depends on (stores of saved activations, the next step's location features) fills the ~1.3 us
// round trip.  After the wait thread 0 publishes the event number to the activation producer (`ready_seq`).
__device__ __forceinline__ void grid_arrive_compute(unsigned* counter, unsigned& target, unsigned nblocks) {
  asm volatile("fence.proxy.async.global;" ::: "memory");  // image writes (generic proxy) -> later bulk-copy reads
  ptx::bar_sync(1, kTcCompute);
  target += nblocks;  // every thread tracks the target: the pollers of grid_wait_compute sit in four different warps
  if (threadIdx.x == 0) red_release_gpu_add(counter, 1u);  // release is cumulative over the writes ordered by the bar.sync above
}
// Four lanes (one per warp 0..3) poll: their loops drift apart, so the counter update is seen about a fifth of an L2 round trip
// after it lands instead of half of one; the first poller to see it publishes the (monotonically increasing) event number,
// the others leave through the shared-memory word.
__device__ __forceinline__ void grid_wait_compute(unsigned* counter, unsigned target, unsigned* ready_seq, unsigned event) {
  if ((threadIdx.x & 31) == 0 && threadIdx.x < 128) {
    while (ld_volatile_shared(ready_seq) < event) {
      if (ld_acquire_gpu(counter) >= target) {
        st_volatile_shared(ready_seq, event);
        break;
      }
    }
  }
  ptx::bar_sync(1, kTcCompute);
}
