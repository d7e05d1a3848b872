// Reverse pass of the Tacotron2 decoder loop on tcgen05 (bf16x3, same numerics contract as decoder_fwd_tc.cu): ONE
// persistent reverse-time kernel for the recurrent gradients.  Weight-gradient GEMMs stay hoisted (decoder_bwd.cu).
// Gradient contract: SURVEY.md A-11 (derived from MSTTS_SV.py:58-98,180-191).
//
// Per reverse step t the four transposed skinny GEMMs run as MMA jobs (A = [dG_hi ; dG_lo] 64 rows, B = 128 weight
// rows hi|lo, K = gate columns; the reference kernels are stored [input row][gate column], i.e. already K-major):
//   JB1  d m0_t      = dG1_t . K1[m0 rows]^T      critical   (accumulator 0)
//   JB2  d h1_{t-1}  = dG1_t . K1[h1 rows]^T      deferred   (accumulator 1, consumed one step later)
//   JA1  d ctx_{t-1} = dG0_t . W0[ctx rows]^T     critical   (accumulator 0)
//   JA2  d h0_{t-1}  = dG0_t . W0[h0 rows]^T      deferred   (accumulator 1)
// Partition: cluster c owns output tile c & 7 (128 rows) and K-quarter c >> 3; its CTA r owns gate columns
// 256*(4*(c>>3)+r).. (K split 16 ways).  The 4 partials of a cluster are reduced through distributed shared memory,
// the 4 K-quarters are summed by whoever consumes the result (global "partials" buffers).
// Phases per step (one grid barrier after each):
//   C'   attention backward, cluster b = batch row b (values / keys / d keys resident in TMEM)
//   B'e  cell-1 gate backward  -> dG1_t (fp32 for the weight-gradient GEMMs + bf16 hi/lo operand image)
//   B'g  JB1 epilogue (+ drains JA2 of step t+1)
//   A'e  cell-0 gate backward  -> dG0_t
//   A'g  JA1 epilogue (+ drains JB2 of step t)
#include <cooperative_groups.h>

#include "decoder_layout.h"
#include "decoder_tc.cuh"

namespace cg = cooperative_groups;

struct DecBwdTcParams {
  int B, Te, T, D, nct;  // nct = D/128 context tiles
  const uint8_t* wimg;   // [128][16][32 KB]: JB1 | JB2 | JA1 | JA2 tiles of every CTA
  uint8_t *ximg_g1, *ximg_g0;      // [64 k-tiles][8 KB] operand images of dG1_t / dG0_t
  float *pm0, *ph1, *ph0, *pctx;   // K-quarter partials [4][B][1024] (pctx: [4][B][D])
  const float *Wq, *F, *sw;
  const float *keys, *values;
  const int* text_len;
  const uint8_t* zone_mask;
  const float *act0, *act1, *c0n, *c1n, *cz0, *cz1, *qf, *cum, *align_tm;
  const float* dm1_proj;
  float *dctx, *dG0, *dG1, *dq, *dkeys, *dF, *dsw;
  float* dF_part;  // [32 clusters][32][128]
  unsigned* barrier;
  long long* dbg;
  // texts of 129 .. 256 positions (TE2): two clusters per batch row, each owning 128 positions
  float* dq2;        // [T,B,128] d q partial of the upper position half (summed into dq after the loop)
  float* ldot_part;  // [2][16]   per-half sum_x a[x] (d a[x] + d cum[x]) of the running step
  float* halo_part;  // [16 rows][2 halves][4 CTAs][16] conv-transpose sums reaching the OTHER half's 15 border positions
  int l2_stream;         // saved activations are loaded, gate gradients stored with the evict-first hint (MSTTS_LOOP_STREAM=0: off)
  int w_evict_last;      // weight tiles are loaded with the L2 evict_last hint (experiment: MSTTS_LOOP_L2=1)
  const float* dctx_in;  // projection part of d ctx.  One cluster per row: the same array as dctx (read, then overwritten with the
                         // total by the same CTA).  TE2: a copy -- the other cluster of the row may still be reading it
};

struct TcBwdSmem {
  uint32_t ring, xbuf, recv, scratch, s_s, wq, xs, cum_s, a_s, da_s, de_s, dcum_s, e_loc, g_loc, e_parts1, e_parts2, dctx_s, ehalf, qred,
      bred, bars, total;
};

__host__ __device__ inline TcBwdSmem tc_bwd_smem(int NS, int Te, int D) {
  const int TeP = (Te + 31) & ~31;
  TcBwdSmem s;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t o = off;
    off += (bytes + 127) & ~127u;
    return o;
  };
  s.ring = take(NS * kWTileBytes);
  s.xbuf = take(4 * kXTileBytes);  // ONE dG slice: dG1_t (JB jobs) and dG0_t (JA jobs) take turns
  s.recv = take(2 * kDecCluster * kTcN * kRecvStride * 4);  // two accumulators
  const uint32_t dps_bytes = (TeP + 32) * 32 * 4, dq_bytes = kTcN * (kAtt + 4) * 4;
  s.scratch = take(dps_bytes > dq_bytes ? dps_bytes : dq_bytes);  // d pre-activations (phase C') / dq rows (phase B'e)
  s.s_s = take(16 * kTcCompute * 4);  // tanh(keys + q + loc) of the coming attention' step, [16 positions][256 threads]
  s.wq = take(kUnitsPerCta * (kAtt + 4) * 4);
  s.xs = take(4 * 2 * kTcN * 8 * 2);
  s.cum_s = take((TeP + 32) * 4);
  s.a_s = take(TeP * 4);
  s.da_s = take(TeP * 4);
  s.de_s = take(TeP * 4);
  s.dcum_s = take(TeP * 4);
  s.e_loc = take(TeP * 4);
  s.g_loc = take(TeP * 4);
  s.e_parts1 = take(kDecCluster * TeP * 4);
  s.e_parts2 = take(kDecCluster * TeP * 4);
  s.dctx_s = take((D / kDecCluster) * 4);
  s.ehalf = take(2 * TeP * 4);
  s.qred = take(8 * 32 * 4);
  s.bred = take(32 * 4);
  s.bars = take(256);
  s.total = off;
  return s;
}

// TE2 = 1: texts of 129 .. 256 positions (B <= 16).  The values / keys / d-keys of a batch row no longer fit one cluster's
// TMEM, so TWO clusters serve a row, each owning 128 text positions (cluster c: row c & 15, positions 128 (c >> 4) ..).
// Three quantities cross the position halves: the softmax' inner product sum_x a (d a + d cum) (one float per row and step:
// exchanged through global memory around ONE extra grid barrier per step), d q (two partial buffers, summed by the
// consumer) and the 15 border positions of the location-conv transpose (exchanged through global memory one phase later).
template <int NS, int TE2>
__global__ void __cluster_dims__(kDecCluster, 1, 1) __launch_bounds__(kTcThreads, 1)
    decoder_bwd_tc_kernel(const DecBwdTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = P.B, Teg = P.Te, D = P.D, Dq = D / kDecCluster, T = P.T;
  const int crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kDecCluster;
  const int arow = TE2 ? (cid & 15) : cid;       // batch row of this cluster's attention' phase
  const int ph = TE2 ? (cid >> 4) : 0;            // TE2: position half
  const int x0 = 128 * ph;                        // first text position of this cluster
  const int Te = TE2 ? max(0, min(128, Teg - x0)) : Teg;  // positions of this cluster inside the padded text
  const int TeP = TE2 ? 128 : ((Teg + 31) & ~31);
  const int tile = cid & 7, kq = cid >> 3;         // output tile / K-quarter of this cluster
  const int kslice = kq * 4 + crank;               // gate columns 256*kslice .. +255
  const bool isctx = tile < P.nct;                 // JA1 exists for this cluster
  const TcBwdSmem L = tc_bwd_smem(NS, TE2 ? 128 : Teg, D);

  uint8_t* ring = smem + L.ring;
  uint8_t* xbuf = smem + L.xbuf;  // dG1_t slice from barrier 2 until JB2(t) has read it, then dG0_t slice (JA jobs)
  float* recv = reinterpret_cast<float*>(smem + L.recv);  // [acc 2][src 4][batch 32][40]
  float* scratch = reinterpret_cast<float*>(smem + L.scratch);
  float* s_s = reinterpret_cast<float*>(smem + L.s_s);
  float* wq_s = reinterpret_cast<float*>(smem + L.wq);    // [8][132]
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem + L.xs);  // [gate 4][hi/lo][32][8]
  float* cum_s = reinterpret_cast<float*>(smem + L.cum_s);
  float* a_s = reinterpret_cast<float*>(smem + L.a_s);
  float* dcum_s = reinterpret_cast<float*>(smem + L.dcum_s);
  float* e_parts1 = reinterpret_cast<float*>(smem + L.e_parts1);
  float* e_parts2 = reinterpret_cast<float*>(smem + L.e_parts2);
  float* dctx_s = reinterpret_cast<float*>(smem + L.dctx_s);
  float* ehalf = reinterpret_cast<float*>(smem + L.ehalf);
  float* qred = reinterpret_cast<float*>(smem + L.qred);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + L.bars);  // [NS]
  uint64_t* empty = wfull + NS;        // [NS]
  uint64_t* xfull = empty + NS;        // [2]
  uint64_t* job_done = xfull + 2;      // [4]
  uint64_t* acc1_free = job_done + 4;  // accumulator 1 drained by the epilogue
  uint64_t* rs_bar = acc1_free + 1;
  uint64_t* e_bar = rs_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_bar + 1);
  unsigned* ready_seq = tmem_slot + 1;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&wfull[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&xfull[0], 1);
    ptx::mbar_init(&xfull[1], 1);
    for (int j = 0; j < 4; ++j) ptx::mbar_init(&job_done[j], 1);
    ptx::mbar_init(acc1_free, 1);
    ptx::mbar_init(rs_bar, 1);
    ptx::mbar_init(e_bar, 1);
    *ready_seq = 0;
    ptx::fence_mbar_init();
  }
  if (warp == kTcMmaWarp) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: 0..255 accumulators (acc 0 on lanes 0-15 of each quarter, acc 1 on lanes 16-31);
  // 256.. values slice (lane = text position, column = context dim); 448.. keys; 480.. d keys
  // (keys / d keys: 16-position block w on lane quarter w & 3, columns (w >> 2)*16 + p, lane = attention unit)
  const uint32_t tmem_val = tmem + 256, tmem_keys = tmem + 448, tmem_dkeys = tmem + 480;

  // ---- one-time staging by the compute warps ----
  float F_reg[kConvK], dF_reg[kConvK];
  float sw_l = 0.f, dsw_acc = 0.f;
#pragma unroll
  for (int k = 0; k < kConvK; ++k) {
    F_reg[k] = 0.f;
    dF_reg[k] = 0.f;
  }
  if (warp < 8) {
    const int unit0 = blockIdx.x * kUnitsPerCta;
    for (int i = tid; i < kUnitsPerCta * kAtt; i += kTcCompute)
      wq_s[(i / kAtt) * (kAtt + 4) + (i % kAtt)] = P.Wq[(size_t)(unit0 + i / kAtt) * kAtt + (i % kAtt)];
#pragma unroll
    for (int k = 0; k < kConvK; ++k) F_reg[k] = P.F[k * kAtt + crank * 32 + lane];
    sw_l = P.sw[crank * 32 + lane];
    for (int i = tid; i < TeP; i += kTcCompute) dcum_s[i] = 0.f;
    if (arow < B) {
      const int q = warp & 3, hf = warp >> 2;
      {  // values: lane = text position x = 32q + lane, columns = this CTA's context dims; warps 0-3 / 4-7 take half each
        const int x = q * 32 + lane;
        const float* vg = P.values + ((size_t)arow * Teg + x0 + x) * D + crank * Dq;
        for (int d0 = hf * (Dq / 2); d0 < (hf + 1) * (Dq / 2); d0 += 32) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (x < Te) ? __float_as_uint(vg[d0 + j]) : 0u;
          ptx::tmem_st32(tmem_val + ((uint32_t)(q * 32) << 16) + d0, v);
        }
      }
      {  // keys / d keys of position block `warp`
        uint32_t v[16], z[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int x = warp * 16 + p;
          v[p] = (x < Te) ? __float_as_uint(P.keys[((size_t)arow * Teg + x0 + x) * kAtt + crank * 32 + lane]) : 0u;
          z[p] = 0u;
        }
        ptx::tmem_st16(tmem_keys + ((uint32_t)(q * 32) << 16) + hf * 16, v);
        ptx::tmem_st16(tmem_dkeys + ((uint32_t)(q * 32) << 16) + hf * 16, z);
      }
      ptx::tmem_wait_st();
    }
  }
  ptx::tc_fence_before();
  cluster.sync();
  ptx::tc_fence_after();

  if (warp == 8 || warp == 9) {
    // =========================== weight-tile producers (even / odd tiles) ===========================
    if (lane == 0) {
      const uint8_t* wsrc = P.wimg + (size_t)blockIdx.x * 16 * kWTileBytes;
      const uint64_t wpol = P.w_evict_last ? ptx::l2_policy_evict_last() : 0ull;
      int i = 0;
      for (int t = T - 1; t >= 0; --t) {
        for (int job = 0; job < 4; ++job) {
          if (job > 0 && (t == 0 || (job == 2 && !isctx))) continue;
          for (int kt = 0; kt < 4; ++kt, ++i) {
            if ((i & 1) != (warp - 8)) continue;
            const int s = i % NS, round = i / NS;
            if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
            ptx::mbar_arrive_expect_tx(&wfull[s], kWTileBytes);
            if (P.w_evict_last)
              ptx::bulk_g2s_hint(ring + (size_t)s * kWTileBytes, wsrc + (size_t)(job * 4 + kt) * kWTileBytes, kWTileBytes, &wfull[s], wpol);
            else
              ptx::bulk_g2s(ring + (size_t)s * kWTileBytes, wsrc + (size_t)(job * 4 + kt) * kWTileBytes, kWTileBytes, &wfull[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // =========================== activation producer: one bulk copy per dG slice ===========================
    if (lane == 0) {
      for (int t = T - 1; t >= 0; --t) {
        const unsigned sidx = (unsigned)(T - 1 - t);
        // dG1_t slice (the buffer's previous readers: JA1 / JA2 of step t+1, long finished)
        if (sidx > 0) ptx::mbar_wait(&job_done[3], (sidx - 1) & 1);
        while (ld_volatile_shared(ready_seq) < (5u + TE2) * sidx + 2 + TE2) {  // TE2: one more barrier per step, inside phase C'
        }
        ptx::mbar_arrive_expect_tx(&xfull[1], 4 * kXTileBytes);
        ptx::bulk_g2s(xbuf, P.ximg_g1 + (size_t)kslice * 4 * kXTileBytes, 4 * kXTileBytes, &xfull[1]);
        if (t == 0) break;
        // dG0_t slice (previous readers: JB1 / JB2 of this step; JB2 ends ~3 us before barrier 4 passes)
        ptx::mbar_wait(&job_done[1], sidx & 1);
        while (ld_volatile_shared(ready_seq) < (5u + TE2) * sidx + 4 + TE2) {
        }
        ptx::mbar_arrive_expect_tx(&xfull[0], 4 * kXTileBytes);
        ptx::bulk_g2s(xbuf, P.ximg_g0 + (size_t)kslice * 4 * kXTileBytes, 4 * kXTileBytes, &xfull[0]);
      }
    }
    __syncwarp();
  } else if (warp == kTcMmaWarp) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(64, 256);
      int i = 0;
      uint32_t free_parity = 0;
      for (int t = T - 1; t >= 0; --t) {
        const uint32_t sp = (uint32_t)(T - 1 - t) & 1u;
        for (int job = 0; job < 4; ++job) {
          if (job > 0 && (t == 0 || (job == 2 && !isctx))) continue;
          if (job == 0) ptx::mbar_wait(&xfull[1], sp);
          if (job == 2 || (job == 3 && !isctx)) ptx::mbar_wait(&xfull[0], sp);
          if (job & 1) {  // deferred job: the epilogue must have drained accumulator 1 first
            ptx::mbar_wait(acc1_free, free_parity);
            free_parity ^= 1u;
          }
          const uint32_t d = tmem + ((job & 1) ? (16u << 16) : 0u);
          const uint32_t xbase = ptx::smem_u32(xbuf);
          for (int kt = 0; kt < 4; ++kt, ++i) {
            const int s = i % NS, round = i / NS;
            ptx::mbar_wait(&wfull[s], round & 1);
            ptx::tc_fence_after();
            const uint32_t wbase = ptx::smem_u32(ring + (size_t)s * kWTileBytes);
#pragma unroll
            for (int k = 0; k < kTcKT / 16; ++k) {
              const uint64_t a = ptx::umma_desc(xbase + kt * kXTileBytes + k * 256, kTcLBO, kTcSBO);
              const uint64_t bd = ptx::umma_desc(wbase + k * 256, kTcLBO, kTcSBO);
              ptx::umma_bf16(d, a, bd, idesc, (kt == 0 && k == 0) ? 0u : 1u);
            }
            ptx::umma_commit(&empty[s]);
          }
          ptx::umma_commit(&job_done[job]);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== compute warps ===========================
    const int u8 = tid & 7, b = tid >> 3;  // gate-backward element of this thread: (unit, batch row) in both cells
    const int unit0 = blockIdx.x * kUnitsPerCta, unit = unit0 + u8;
    const bool brow = b < B;
    const size_t BC = (size_t)B * kCell, BG = (size_t)B * kGates;
    const int q4 = warp & 3, half = warp >> 2;
    uint32_t rs_parity = 0, e_parity = 0;
    unsigned bar_target = 0, ev = 0;
    float dc0 = 0.f, dh0d = 0.f, dc1 = 0.f, dh1d = 0.f;  // carries: d c (zoned) and the direct zoneout path of d h
    const int tlg = (arow < B) ? min(P.text_len[arow], Teg) : 0;   // valid positions of the row
    const int tl = TE2 ? max(0, min(128, tlg - x0)) : tlg;          // ... of this cluster's positions
    const bool halo = TE2 && tlg > 128;                              // both halves hold valid positions
    const uint32_t recv_addr = ptx::smem_u32(recv);
    float* dps = scratch;   // [(TeP+32)][32] d pre-activations of attention'(t): live until the prologue of step t-1
    float* dq_s = recv;     // [32][132] dq rows of phase B'e (recv is idle between the drains of A'g(t+1) and B'g(t))
    long long* dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
    // profiling aid: CTA 0 stamps every step with its SM clock; at the middle step EVERY CTA also stamps the global timer
    // (rows 0..127 of the same buffer, one row per CTA) so that tools/phase_times.py can show the arrival spread at each barrier
    long long* dbg_all = (P.dbg && tid == 0) ? P.dbg + (size_t)blockIdx.x * 32 : nullptr;
#define STAMP(k)                                                                 \
  do {                                                                           \
    if (dbg) dbg[(size_t)(T - 1 - t) * 32 + (k)] = clock64();                    \
    if (dbg_all && t == T / 2) {                                                 \
      unsigned long long gt;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));                     \
      dbg_all[(k)] = (long long)gt;                                              \
    }                                                                            \
  } while (0)

    // Drains both accumulators (acc 0 = lanes 0-15, acc 1 = lanes 16-31 of every TMEM lane quarter), adds the hi/lo
    // quadrants, reduces the cluster's 4 K-slices through DSMEM and writes this CTA's 32 rows of each to the partial
    // buffers out0 / out1 ([kq][b][row]); a null pointer skips that accumulator.
    auto drain = [&](uint64_t* done_bar, uint32_t done_parity, float* out0, int ld0, float* out1, int ld1) {
      if (tid == 0) ptx::mbar_arrive_expect_tx(rs_bar, 2 * kDecCluster * 32 * kTcN * 4);
      mbar_wait_warp(done_bar, done_parity);
      ptx::tc_fence_after();
      const int acc = lane >> 4, l16 = lane & 15;
      const bool pusher = l16 < 8;
      const int bq = 8 * q4 + (l16 & 7);
      const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16) + half * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t vh[32], vl[32];
        ptx::tmem_ld32(ta + ch * 32, vh);
        ptx::tmem_ld32(ta + 128 + ch * 32, vl);
        ptx::tmem_wait_ld();
        const int rr = half * 2 + ch;
        const uint32_t dst = ptx::mapa(recv_addr, (uint32_t)rr) + (uint32_t)((((acc * kDecCluster + crank) * kTcN + bq) * kRecvStride) * 4);
        const uint32_t rbar = ptx::mapa(ptx::smem_u32(rs_bar), (uint32_t)rr);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = __uint_as_float(vh[j + e]) + __uint_as_float(vl[j + e]);
            o[e] = v + __shfl_down_sync(0xffffffffu, v, 8);
          }
          if (pusher)
            ptx::st_async_v4(dst + j * 4, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]), rbar);
        }
      }
      ptx::tc_fence_before();
      ptx::bar_sync(1, kTcCompute);
      if (tid == 0) ptx::mbar_arrive(acc1_free);  // every warp has read accumulator 1
      mbar_wait_warp(rs_bar, rs_parity);
      rs_parity ^= 1u;
      // this CTA's rows 32*crank .. +31 of the tile: sum the 4 K-slices, write the K-quarter partial
      const int jj = tid & 31, bg = tid >> 5;
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        float* out = a2 ? out1 : out0;
        const int ld = a2 ? ld1 : ld0;
        if (!out) continue;
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          const int bb = bg * 4 + i2;
          if (bb < B) {
            float s = 0.f;
#pragma unroll
            for (int src = 0; src < kDecCluster; ++src) s += recv[((a2 * kDecCluster + src) * kTcN + bb) * kRecvStride + jj];
            out[((size_t)kq * B + bb) * ld + tile * 128 + crank * 32 + jj] = s;
          }
        }
      }
    };

    // writes the 4 gate gradients of this thread as rows of the dG operand image and (deferred) as fp32
    auto publish_dG = [&](const CellGradTc& g, uint8_t* ximg) {
      const float gv[4] = {g.di, g.dj, g.df, g.dop};
      if (brow) {
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          __nv_bfloat16 hi, lo;
          split_bf16(gv[gi], hi, lo);
          xs[(gi * 2 + 0) * 256 + b * 8 + u8] = hi;
          xs[(gi * 2 + 1) * 256 + b * 8 + u8] = lo;
        }
      }
      ptx::bar_sync(1, kTcCompute);
      {
        const int gi = tid >> 6, hl = (tid >> 5) & 1, r2 = tid & 31;
        if (r2 < B) {
          const int col0 = gi * kCell + unit0;
          uint8_t* img = ximg + (size_t)(col0 >> 6) * kXTileBytes + (size_t)((col0 & 63) >> 3) * 128 + ximg_row_offset(r2, hl);
          *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(xs + (gi * 2 + hl) * 256 + r2 * 8);
        }
      }
    };

    // Stages everything of attention'(t) that depends only on saved forward data: alignment, cumulative alignment
    // (shifted by 15, zero borders), zero borders of the d pre-activation buffer, and s = tanh(keys + q + loc) of this
    // warp's 16 positions.  Runs inside the barrier wait that precedes phase C'(t).
    // (split in two: the staging half reads HBM and runs inside the wait of barrier 4 -- nothing reads cum_s / a_s after the
    //  wait of barrier 2 -- so that the wait of barrier 5 holds only the convolution and the tanh)
    float qf_next = 0.f;
    auto attention_stage = [&](int t, bool with_halo) {
      if (arow >= B) return;
      if (with_halo && halo && tid < 15) {
        // d cum_t: conv-transpose sums that the other position half's d pre-activations of step t + 1 send into this half's
        // 15 border positions (written by the peer cluster during its barrier-2 wait, two grid barriers ago)
        const float* hp = P.halo_part + (size_t)((arow * 2 + (1 - ph)) * kDecCluster) * 16 + tid;
        const float add = ((__ldcg(hp) + __ldcg(hp + 16)) + __ldcg(hp + 32)) + __ldcg(hp + 48);
        dcum_s[ph == 0 ? 113 + tid : tid] += add;
      }
      const int bb = arow;
      const float* al = P.align_tm + ((size_t)t * B + bb) * Teg + x0;
      const float* cum_prev = P.cum + ((size_t)t * B + bb) * Teg;
      qf_next = P.qf[((size_t)t * B + bb) * kAtt + crank * 32 + lane];
      for (int i = tid; i < TeP + 32; i += kTcCompute) {
        const int x = x0 + i - 15;   // the 15-position borders reach into the other half (TE2)
        cum_s[i] = (x >= 0 && x < Teg) ? cum_prev[x] : 0.f;
      }
      for (int x = tid; x < TeP; x += kTcCompute) a_s[x] = (x < Te) ? al[x] : 0.f;
      for (int i = tid; i < 15 * 32; i += kTcCompute) dps[i] = 0.f;
      for (int i = (15 + tl) * 32 + tid; i < (TeP + 32) * 32; i += kTcCompute) dps[i] = 0.f;
    };
    auto attention_prologue = [&]() {  // needs a barrier among the compute warps after attention_stage
      if (arow >= B) return;
      const float qf = qf_next;
      const int t0 = warp * 16;
      if (t0 < tl) {
        float acc[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) acc[p] = 0.f;
#pragma unroll
        for (int c = 0; c < 16 + kConvK - 1; ++c) {
          const float cv = cum_s[t0 + c];
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int k = c - p;
            if (k >= 0 && k < kConvK) acc[p] = fmaf(cv, F_reg[k], acc[p]);
          }
        }
        uint32_t kv[16];
        ptx::tmem_ld16(tmem_keys + ((uint32_t)(q4 * 32) << 16) + half * 16, kv);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int p = 0; p < 16; ++p) s_s[p * kTcCompute + tid] = tanhf(__uint_as_float(kv[p]) + qf + acc[p]);
      }
    };
    attention_stage(T - 1, false);
    ptx::bar_sync(1, kTcCompute);
    attention_prologue();
    ptx::bar_sync(1, kTcCompute);
    // Saved forward activations come from HBM (2.4 GB per launch, nothing of it is L2-resident): every phase issues the
    // loads of its saved operands BEFORE the grid-barrier wait that precedes it, so their latency rides in the wait.
    float pf_act[4] = {0.f, 0.f, 0.f, 0.f}, pf_cn = 0.f, pf_cz = 0.f, pf_mc = 0.f, pf_mh = 0.f, pf_dm = 0.f, pf_dh = 0.f;
    float pf_dctx = 0.f;
    if (arow < B && tid < Dq) pf_dctx = P.dctx_in[((size_t)(T - 1) * B + arow) * D + crank * Dq + tid];

    for (int t = T - 1; t >= 0; --t) {
      const uint8_t* zm = P.zone_mask + (size_t)t * 4 * BC;
      const uint32_t sp = (uint32_t)(T - 1 - t) & 1u;
      const bool last = (t == T - 1);
      STAMP(0);
      // ================= phase C': attention backward, batch row = cluster index =================
      // (alignment, cumulative alignment and tanh(keys + q + loc) of this step were staged by attention_prologue(t)
      //  inside the previous barrier wait: they depend only on saved forward data)
      float dav[4] = {0.f, 0.f, 0.f, 0.f}, ldot = 0.f;
      if (arow < B) {
        const int bb = arow;
        if (tid < Dq) {  // total gradient w.r.t. ctx_t: projection part + the 4 K-quarter partials of JA1(t+1)
          const size_t gi = ((size_t)t * B + bb) * D + crank * Dq + tid;
          float v = pf_dctx;
          if (!last) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) v += __ldcg(P.pctx + ((size_t)k4 * B + bb) * D + crank * Dq + tid);
            if (ph == 0) P.dctx[gi] = v;
          }
          dctx_s[tid] = v;
        }
        ptx::bar_sync(1, kTcCompute);
        {  // partial d a[x] over this CTA's context dims: thread = text position, columns = dims (two halves of warps)
          const int x = q4 * 32 + lane;
          float s = 0.f;
          const uint32_t va = tmem_val + ((uint32_t)(q4 * 32) << 16);
          for (int d0 = half * (Dq / 2); d0 < (half + 1) * (Dq / 2); d0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(va + d0, v);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) s = fmaf(dctx_s[d0 + j], __uint_as_float(v[j]), s);
          }
          if (x < TeP) ehalf[half * TeP + x] = s;
        }
        ptx::bar_sync(1, kTcCompute);
        if (tid == 0) ptx::mbar_arrive_expect_tx(e_bar, (uint32_t)(kDecCluster * ((tl + 3) & ~3) * 4));  // whole groups of four
        {
          const uint32_t ep = ptx::smem_u32(e_parts1) + (uint32_t)(crank * TeP * 4);
          const uint32_t eb = ptx::smem_u32(e_bar);
          for (int x = tid * 4; x < tl; x += kTcCompute * 4) {  // 16-byte remote stores (scalar ones are disproportionately expensive)
            const float4 a = *reinterpret_cast<const float4*>(ehalf + x), b = *reinterpret_cast<const float4*>(ehalf + TeP + x);
#pragma unroll
            for (uint32_t dst = 0; dst < (uint32_t)kDecCluster; ++dst)
              ptx::st_async_v4(ptx::mapa(ep, dst) + x * 4, __float_as_uint(a.x + b.x), __float_as_uint(a.y + b.y), __float_as_uint(a.z + b.z),
                               __float_as_uint(a.w + b.w), ptx::mapa(eb, dst));
          }
        }
        mbar_wait_warp(e_bar, e_parity);
        e_parity ^= 1u;
        // softmax backward (every warp reduces all positions redundantly: Te <= 128 -> 4 per lane); this warp's
        // 16 positions of d e stay in registers (position t0 + p comes from lane (t0 + p) & 31, slot (t0 + p) >> 5)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = lane + 32 * j;
          dav[j] = 0.f;
          if (x < tl) {
            dav[j] = (((e_parts1[x] + e_parts1[TeP + x]) + e_parts1[2 * TeP + x]) + e_parts1[3 * TeP + x]) + dcum_s[x];
            ldot = fmaf(a_s[x], dav[j], ldot);
          }
        }
        ldot = warp_sum(ldot);
        // TE2: the inner product runs over BOTH position halves of the row: publish this half's part (every warp of every CTA
        // of the cluster holds the same value), meet the other cluster at a grid barrier, add the two parts in a fixed order
        if (TE2 && crank == 0 && tid == 0) P.ldot_part[ph * 16 + bb] = ldot;
      }
      if (TE2) {
        grid_arrive_compute(P.barrier, bar_target, gridDim.x);
        grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      }
      if (arow < B) {
        const int bb = arow;
        if (TE2) ldot = __ldcg(P.ldot_part + bb) + __ldcg(P.ldot_part + 16 + bb);
        float de_blk[16];
        {
          float dev[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = lane + 32 * j;
            dev[j] = (x < tl) ? a_s[x] * (dav[j] - ldot) : 0.f;
          }
          const int t0 = warp * 16;
          const float mine = dev[0] * (float)((t0 >> 5) == 0) + dev[1] * (float)((t0 >> 5) == 1) + dev[2] * (float)((t0 >> 5) == 2) +
                             dev[3] * (float)((t0 >> 5) == 3);
#pragma unroll
          for (int p = 0; p < 16; ++p) de_blk[p] = __shfl_sync(0xffffffffu, mine, (t0 + p) & 31);
        }
        // d pre-activation of the energy for this CTA's 32 attention units: warp = 16-position block, lane = unit
        {
          const int t0 = warp * 16;
          float dq_acc = 0.f;
          if (t0 < tl) {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float sv = s_s[p * kTcCompute + tid];
              const bool ok = t0 + p < tl;
              const float dpre = ok ? de_blk[p] * sw_l * (1.f - sv * sv) : 0.f;
              if (ok) dsw_acc = fmaf(de_blk[p], sv, dsw_acc);
              dq_acc += dpre;
              dps[(15 + t0 + p) * 32 + lane] = dpre;
            }
          }
          qred[warp * 32 + lane] = dq_acc;
        }
        ptx::bar_sync(1, kTcCompute);
        if (tid < 32) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += qred[w * 32 + tid];
          (ph ? P.dq2 : P.dq)[((size_t)t * B + bb) * kAtt + crank * 32 + tid] = s;
        }
      }
      STAMP(1);
      grid_arrive_compute(P.barrier, bar_target, gridDim.x);
      if (brow) {  // saved operands of phase B'e (d h1 partials of JB2(t+1) are complete since the last barrier of step t+1)
        const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) pf_act[gi] = ld_stream(P.act1 + ai + gi * kCell, P.l2_stream);
        pf_cn = ld_stream(P.c1n + (size_t)t * BC + si, P.l2_stream);
        pf_cz = ld_stream(P.cz1 + (size_t)t * BC + si, P.l2_stream);
        pf_mc = (float)zm[2 * BC + si];
        pf_mh = (float)zm[3 * BC + si];
        pf_dm = P.dm1_proj[(size_t)t * BC + si];
        pf_dh = dh1d;
        if (!last) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) pf_dh += __ldcg(P.ph1 + ((size_t)k4 * B + b) * kCell + unit);
        }
      }
      grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      STAMP(2);

      // ================= phase B'e: cell-1 gate backward for this CTA's 8 units =================
      {
        for (int i = tid; i < B * (kAtt / 4); i += kTcCompute) {  // rows padded to 132 floats: conflict-free float4 reads
          float4 v4 = __ldcg(reinterpret_cast<const float4*>(P.dq + (size_t)t * B * kAtt) + i);
          if (TE2) {  // + the upper position half's part
            const float4 u4 = __ldcg(reinterpret_cast<const float4*>(P.dq2 + (size_t)t * B * kAtt) + i);
            v4 = make_float4(v4.x + u4.x, v4.y + u4.y, v4.z + u4.z, v4.w + u4.w);
          }
          reinterpret_cast<float4*>(dq_s)[(i >> 5) * 33 + (i & 31)] = v4;
        }
        ptx::bar_sync(1, kTcCompute);
        CellGradTc g = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (brow) {
          float s = 0.f;
          const float4* dq4 = reinterpret_cast<const float4*>(dq_s + b * (kAtt + 4));
          const float4* wq4 = reinterpret_cast<const float4*>(wq_s + u8 * (kAtt + 4));
#pragma unroll 8
          for (int a = 0; a < kAtt / 4; ++a) {
            const float4 x4 = dq4[a], w4 = wq4[a];
            s = fmaf(x4.x, w4.x, s);
            s = fmaf(x4.y, w4.y, s);
            s = fmaf(x4.z, w4.z, s);
            s = fmaf(x4.w, w4.w, s);
          }
          g = cell_backward_tc(pf_dm + s, pf_dh, dc1, pf_act[0], pf_act[1], pf_act[2], pf_act[3], pf_cn, pf_cz, pf_mc, pf_mh);
          dc1 = g.dc_prev;
          dh1d = g.dh_prev;
        }
        publish_dG(g, P.ximg_g1);
        STAMP(3);
        grid_arrive_compute(P.barrier, bar_target, gridDim.x);
        if (brow) {
          const size_t ai = (size_t)t * BG + (size_t)b * kGates + unit;
          st_stream(P.dG1 + ai, g.di, P.l2_stream);
          st_stream(P.dG1 + ai + kCell, g.dj, P.l2_stream);
          st_stream(P.dG1 + ai + 2 * kCell, g.df, P.l2_stream);
          st_stream(P.dG1 + ai + 3 * kCell, g.dop, P.l2_stream);
        }
        // ---- deferred halves of attention'(t), spread over the barrier waits so that none holds more than the barrier's own
        //      ~1.3 us (all of it inside the wait of barrier 1 took ~3.5 us): here the conv transpose that feeds d cum_{t-1};
        //      the wait of barrier 3 collects it and accumulates d F / d keys from the d pre-activations kept in `dps` ----
        if (arow < B) {
          const int t0 = warp * 16;
          if (t0 < tl) {
            // conv transpose: gradient reaching cum_{t-1} through the location features (partial over this CTA's units)
            float G[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) G[p] = 0.f;
#pragma unroll
            for (int c = 0; c < 16 + kConvK - 1; ++c) {
              const float v = dps[(t0 + c) * 32 + lane];
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                const int k = p + (kConvK - 1) - c;
                if (k >= 0 && k < kConvK) G[p] = fmaf(v, F_reg[k], G[p]);
              }
            }
            const float tot = warp_sum16(G, lane);
            // lanes L, L+2, L+4, L+6 (L % 8 == 0) hold four consecutive positions: one 16-byte remote store
            const float a1 = __shfl_down_sync(0xffffffffu, tot, 2), a2 = __shfl_down_sync(0xffffffffu, tot, 4),
                        a3 = __shfl_down_sync(0xffffffffu, tot, 6);
            if ((lane & 7) == 0) {
              const int x = t0 + warp_sum16_index(lane);
              if (x < tl) {
                const uint32_t ep = ptx::smem_u32(e_parts2) + (uint32_t)((crank * TeP + x) * 4);
                const uint32_t eb = ptx::smem_u32(e_bar);
#pragma unroll
                for (uint32_t dst = 0; dst < (uint32_t)kDecCluster; ++dst)
                  ptx::st_async_v4(ptx::mapa(ep, dst), __float_as_uint(tot), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3),
                                   ptx::mapa(eb, dst));
              }
            }
          }
          if (tid == 0) ptx::mbar_arrive_expect_tx(e_bar, (uint32_t)(kDecCluster * ((tl + 3) & ~3) * 4));  // collected in the wait of barrier 3
          if (halo && warp == 7) {
            // the same transpose for the 15 border positions of the OTHER half (|x - x'| <= 15 reaches across): partial over
            // this CTA's 32 units; the peer cluster adds the four CTA parts to its d cum before its next phase C'
            float H[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) H[j] = 0.f;
            if (ph == 0) {  // positions 128 + j  <-  own positions 113 + j + d, tap 30 - d
#pragma unroll
              for (int j = 0; j < 15; ++j)
#pragma unroll
                for (int d = 0; d < 15 - j; ++d) H[j] = fmaf(dps[(15 + 113 + j + d) * 32 + lane], F_reg[30 - d], H[j]);
            } else {        // positions 113 + j (own -15 + j)  <-  own positions x' <= j, tap j - x'
#pragma unroll
              for (int j = 0; j < 15; ++j)
#pragma unroll
                for (int xs2 = 0; xs2 <= j; ++xs2) H[j] = fmaf(dps[(15 + xs2) * 32 + lane], F_reg[j - xs2], H[j]);
            }
            const float tot = warp_sum16(H, lane);
            if ((lane & 1) == 0) {
              const int j = warp_sum16_index(lane);
              if (j < 15) P.halo_part[(size_t)((arow * 2 + ph) * kDecCluster + crank) * 16 + j] = tot;
            }
          }
        }
        grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      }
      STAMP(4);

      // ================= phase B'g: JB1 epilogue -> d m0_t partials (+ d h0 partials of JA2(t+1)) =================
      drain(&job_done[0], sp, P.pm0, kCell, last ? nullptr : P.ph0, kCell);
      STAMP(5);
      grid_arrive_compute(P.barrier, bar_target, gridDim.x);
      if (brow) {  // saved operands of phase A'e
        const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) pf_act[gi] = ld_stream(P.act0 + ai + gi * kCell, P.l2_stream);
        pf_cn = ld_stream(P.c0n + (size_t)t * BC + si, P.l2_stream);
        pf_cz = ld_stream(P.cz0 + (size_t)t * BC + si, P.l2_stream);
        pf_mc = (float)zm[si];
        pf_mh = (float)zm[BC + si];
      }
      if (arow < B) {  // d cum_{t-1}: the cluster's partial conv-transpose sums pushed during the wait of barrier 2
        mbar_wait_warp(e_bar, e_parity);
        e_parity ^= 1u;
        for (int x = tid; x < tl; x += kTcCompute)
          dcum_s[x] += ((e_parts2[x] + e_parts2[TeP + x]) + e_parts2[2 * TeP + x]) + e_parts2[3 * TeP + x];
      }
      if (arow < B && warp * 16 < tl) {  // d F of attention'(t) (deferred from phase C')
        const int t0 = warp * 16;
        float dp[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) dp[p] = dps[(15 + t0 + p) * 32 + lane];
#pragma unroll
        for (int c = 0; c < 16 + kConvK - 1; ++c) {
          const float cv = cum_s[t0 + c];
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int k = c - p;
            if (k >= 0 && k < kConvK) dF_reg[k] = fmaf(cv, dp[p], dF_reg[k]);
          }
        }
      }
      if (arow < B && warp * 16 < tl) {  // d keys += d pre-activations of attention'(t) (deferred from phase C')
        uint32_t dk[16];
        const uint32_t ka = ((uint32_t)(q4 * 32) << 16) + half * 16;
        ptx::tmem_ld16(tmem_dkeys + ka, dk);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int p = 0; p < 16; ++p) dk[p] = __float_as_uint(__uint_as_float(dk[p]) + dps[(15 + warp * 16 + p) * 32 + lane]);
        ptx::tmem_st16(tmem_dkeys + ka, dk);
        ptx::tmem_wait_st();
      }
      grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      STAMP(6);

      // ================= phase A'e: cell-0 gate backward =================
      {
        CellGradTc g = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (brow) {
          float dm0 = 0.f, dh = dh0d;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) dm0 += __ldcg(P.pm0 + ((size_t)k4 * B + b) * kCell + unit);
          if (!last) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) dh += __ldcg(P.ph0 + ((size_t)k4 * B + b) * kCell + unit);
          }
          g = cell_backward_tc(dm0, dh, dc0, pf_act[0], pf_act[1], pf_act[2], pf_act[3], pf_cn, pf_cz, pf_mc, pf_mh);
          dc0 = g.dc_prev;
          dh0d = g.dh_prev;
        }
        publish_dG(g, P.ximg_g0);
        STAMP(7);
        if (t == 0) {  // nothing upstream of step 0 needs d ctx_{-1} / d h_{-1}
          if (brow) {
            const size_t ai = (size_t)b * kGates + unit;
            st_stream(P.dG0 + ai, g.di, P.l2_stream);
            st_stream(P.dG0 + ai + kCell, g.dj, P.l2_stream);
            st_stream(P.dG0 + ai + 2 * kCell, g.df, P.l2_stream);
            st_stream(P.dG0 + ai + 3 * kCell, g.dop, P.l2_stream);
          }
          break;
        }
        grid_arrive_compute(P.barrier, bar_target, gridDim.x);
        if (brow) {
          const size_t ai = (size_t)t * BG + (size_t)b * kGates + unit;
          st_stream(P.dG0 + ai, g.di, P.l2_stream);
          st_stream(P.dG0 + ai + kCell, g.dj, P.l2_stream);
          st_stream(P.dG0 + ai + 2 * kCell, g.df, P.l2_stream);
          st_stream(P.dG0 + ai + 3 * kCell, g.dop, P.l2_stream);
        }
        attention_stage(t - 1, true);
        grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      }
      STAMP(8);

      // ================= phase A'g: JA1 epilogue -> d ctx_{t-1} partials (+ d h1 partials of JB2(t)) =================
      drain(isctx ? &job_done[2] : &job_done[1], sp, isctx ? P.pctx : nullptr, D, P.ph1, kCell);
      STAMP(9);
      grid_arrive_compute(P.barrier, bar_target, gridDim.x);
      attention_prologue();
      if (arow < B && tid < Dq) pf_dctx = P.dctx_in[((size_t)(t - 1) * B + arow) * D + crank * Dq + tid];
      grid_wait_compute(P.barrier, bar_target, ready_seq, ++ev);
      STAMP(10);
    }
#undef STAMP

    // ---- flush the per-CTA accumulators ----
    ptx::bar_sync(1, kTcCompute);
    if (arow < B) {  // d keys: TMEM -> global
      uint32_t dk[16];
      ptx::tmem_ld16(tmem_dkeys + ((uint32_t)(q4 * 32) << 16) + half * 16, dk);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int x = warp * 16 + p;
        if (x < Te) P.dkeys[((size_t)arow * Teg + x0 + x) * kAtt + crank * 32 + lane] = __uint_as_float(dk[p]);
      }
    }
    // d F / d w: cross-warp reduction in shared memory (recv is free now), then this cluster's sums go to its own slot of
    // dF_part; a small kernel adds the 32 slots in cluster order (atomics would make the last bits depend on arrival order)
    float* red = recv;  // [8 warps][32 taps][32 units] = 32 KB
#pragma unroll
    for (int k = 0; k < kConvK; ++k) red[(warp * 32 + k) * 32 + lane] = dF_reg[k];
    red[(warp * 32 + kConvK) * 32 + lane] = dsw_acc;
    ptx::bar_sync(1, kTcCompute);
    for (int i = tid; i < 32 * 32; i += kTcCompute) {
      const int k = i >> 5, u = i & 31;
      float v = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) v += red[(w2 * 32 + k) * 32 + u];
      P.dF_part[((size_t)cid * 32 + k) * kAtt + crank * 32 + u] = v;
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kTcMmaWarp) ptx::tmem_dealloc(tmem, 512);
  cluster.sync();
}

// ---- weight image for the reverse GEMMs: rows = input index (already K-major in the reference layout) ----
// CTA (c, r): tile = c & 7, gate columns 256*(4*(c>>3)+r) .. ; 16 tiles: JB1 (m0 rows of K1) | JB2 (h1 rows of K1) |
// JA1 (ctx rows of W0 = the two context blocks of K0 added) | JA2 (h0 rows of K0)
__global__ void prep_wimg_bwd_kernel(const float* __restrict__ K0, const float* __restrict__ K1, uint8_t* __restrict__ wimg,
                                     int D) {
  const size_t total = (size_t)kDecGrid * 16 * 8 * 128;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int kc = (int)(idx & 7);          // 8-column chunk inside the 64-column tile (fastest: coalesced 32 B reads)
    const int i = (int)((idx >> 3) & 127);  // row of the tile
    const size_t tq = idx >> 10;
    const int q = (int)(tq & 15), cta = (int)(tq >> 4);
    const int c = cta >> 2, r = cta & 3;
    const int tile = c & 7, kslice = (c >> 3) * 4 + r;
    const int job = q >> 2, kt = q & 3;
    const int col = kslice * 256 + kt * kTcKT + kc * 8;
    const int j = tile * 128 + i;
    float v[8];
    if (job == 0 || job == 1) {
      const float* src = K1 + (size_t)((job == 1 ? kCell : 0) + j) * kGates + col;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = src[e];
    } else if (job == 2) {
      if (j < D) {
        const float* s1 = K0 + (size_t)(kPrenet + j) * kGates + col;
        const float* s2 = K0 + (size_t)(kPrenet + D + j) * kGates + col;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = s1[e] + s2[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
    } else {
      const float* src = K0 + (size_t)(kPrenet + 2 * D + j) * kGates + col;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = src[e];
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
    uint8_t* tile_p = wimg + tq * kWTileBytes + (size_t)(i >> 3) * 1024 + (size_t)kc * 128 + (size_t)(i & 7) * 16;
    *reinterpret_cast<uint4*>(tile_p) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(tile_p + kWTileBytes / 2) = *reinterpret_cast<const uint4*>(lo);
  }
}

__global__ void add_dq2_kernel(float* __restrict__ dq, const float* __restrict__ dq2, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dq)[i];
    const float4 b = reinterpret_cast<const float4*>(dq2)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dq)[i] = a;
  }
}

// ======================================== host side ================================================
bool dec_tc_supported(int B, int Te, int D);  // decoder_fwd_tc.cu

template <int NS, int TE2>
static int launch_bwd_tc(const DecBwdTcParams& P, cudaStream_t stream, size_t smem, bool* ok) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  int max_optin = 0;
  MSTTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem > (size_t)max_optin) {
    *ok = false;
    return MSTTS_OK;
  }
  *ok = true;
  MSTTS_CUDA(cudaFuncSetAttribute(decoder_bwd_tc_kernel<NS, TE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(kDecGrid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // cooperative launch: the runtime gang-schedules the whole grid (all 32 clusters resident before any CTA starts), so the
  // grid barrier cannot deadlock behind a concurrent kernel that holds SMs (an NCCL kernel on another stream, MPS)
  cudaLaunchAttribute coop_attr[1];
  dec_cooperative_attr(&cfg, coop_attr);
  int nclusters = 0;
  MSTTS_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, decoder_bwd_tc_kernel<NS, TE2>, &cfg));
  MSTTS_REQUIRE(nclusters * kDecCluster >= kDecGrid, MSTTS_E_DEVICE,
                "decoder_bwd_tc: device co-schedules only %d clusters of %d (need %d)", nclusters, kDecCluster,
                kDecGrid / kDecCluster);
  // (an access-policy window that pins the weight image in persisting L2 was measured and removed: the carve-out shrinks the
  //  L2 of everything else and the step got slower; the eviction-priority hints of DecBwdTcParams::l2_stream do the job)
  mstts_timer_start(1, stream);
  MSTTS_CUDA(cudaLaunchKernelEx(&cfg, decoder_bwd_tc_kernel<NS, TE2>, P));
  mstts_timer_stop(1, stream);
  return MSTTS_OK;
}

// runs the reverse loop; expects dm1_proj / dctx (projection parts) prepared and dF / dsw zeroed by the caller
// hi | lo tile images of the two cell kernels in the reverse loop's streaming order: depends on the weights only
int dec_bwd_tc_prep_weights(const MsttsDecoderWeights* w, const DecLayout& l, char* ws, int D, cudaStream_t s) {
  prep_wimg_bwd_kernel<<<148 * 8, 256, 0, s>>>(w->cell0_kernel, w->cell1_kernel, (uint8_t*)(ws + l.wimg_b), D);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

int dec_bwd_tc_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws, cudaStream_t s,
                     bool wimg_ready) {
  auto F = [&](size_t off) { return (float*)(ws + off); };
  const int D = io->D, B = io->B;
  MSTTS_REQUIRE(dec_tc_supported(B, io->Te, D), MSTTS_E_UNSUPPORTED, "decoder_bwd bf16x3: unsupported shape B=%d Te=%d D=%d", B,
                io->Te, D);
  DecBwdTcParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.Te = io->Te; P.T = io->n_steps; P.D = D; P.nct = D / 128;
  P.wimg = (const uint8_t*)(ws + l.wimg_b);
  P.ximg_g1 = (uint8_t*)(ws + l.ximg_g1); P.ximg_g0 = (uint8_t*)(ws + l.ximg_g0);
  P.pm0 = F(l.pm0); P.ph1 = F(l.ph1); P.ph0 = F(l.ph0); P.pctx = F(l.pctx);
  P.Wq = w->query_kernel; P.F = F(l.locF); P.sw = w->score_w;
  P.keys = F(l.keys); P.values = F(l.values); P.text_len = io->text_len; P.zone_mask = io->zone_mask;
  P.act0 = F(l.act0); P.act1 = F(l.act1); P.c0n = F(l.c0n); P.c1n = F(l.c1n); P.cz0 = F(l.cz0); P.cz1 = F(l.cz1);
  P.qf = F(l.qf); P.cum = F(l.cum); P.align_tm = F(l.align_tm); P.dm1_proj = F(l.dm1_proj);
  P.dctx = F(l.dctx); P.dG0 = F(l.dG0); P.dG1 = F(l.dG1); P.dq = F(l.dq); P.dkeys = F(l.dkeys);
  P.dF = F(l.dF); P.dsw = F(l.dsw); P.dF_part = F(l.dF_part);
  P.dctx_in = P.dctx;
  {
    const char* e = getenv("MSTTS_LOOP_L2");
    P.w_evict_last = e ? atoi(e) : 0;
    e = getenv("MSTTS_LOOP_STREAM");
    P.l2_stream = e ? atoi(e) : 1;
  }
  P.barrier = (unsigned*)(ws + l.barrier);
  P.dbg = (long long*)(ws + l.dbg_b);
  if (!wimg_ready) {
    int rcw = dec_bwd_tc_prep_weights(w, l, ws, D, s);
    if (rcw) return rcw;
  }
  MSTTS_CUDA(cudaMemsetAsync(ws + l.ximg_g1, 0, l.ximg_g_end - l.ximg_g1, s));
  bool ok = false;
  int rc;
  if (io->Te > 128) {  // two clusters per row, 128 positions each: the shared-memory layout of a 128-position text
    P.dq2 = F(l.dq2); P.ldot_part = F(l.xexch); P.halo_part = F(l.xexch) + 32;
    P.dctx_in = F(l.dctx_in);
    MSTTS_CUDA(cudaMemcpyAsync(ws + l.dctx_in, ws + l.dctx, (size_t)io->n_steps * B * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    rc = launch_bwd_tc<3, 1>(P, s, tc_bwd_smem(3, 128, D).total, &ok);
    if (rc) return rc;
    if (!ok) rc = launch_bwd_tc<2, 1>(P, s, tc_bwd_smem(2, 128, D).total, &ok);
    if (rc) return rc;
    MSTTS_REQUIRE(ok, MSTTS_E_UNSUPPORTED, "decoder_bwd_tc: shared memory does not fit for Te=%d", io->Te);
    // d q = lower-half part + upper-half part (the weight-gradient product of the query layer reads the sum)
    const size_t n = (size_t)io->n_steps * B * kAtt;
    add_dq2_kernel<<<(int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 : 148 * 8), 256, 0, s>>>(P.dq, P.dq2, n / 4);
    MSTTS_CUDA(cudaGetLastError());
    return MSTTS_OK;
  }
  rc = launch_bwd_tc<3, 0>(P, s, tc_bwd_smem(3, io->Te, D).total, &ok);  // 3 of a critical job's 4 weight tiles prefetched
  if (rc) return rc;
  if (!ok) rc = launch_bwd_tc<2, 0>(P, s, tc_bwd_smem(2, io->Te, D).total, &ok);
  if (rc) return rc;
  MSTTS_REQUIRE(ok, MSTTS_E_UNSUPPORTED, "decoder_bwd_tc: shared memory does not fit for Te=%d", io->Te);
  return MSTTS_OK;
}
