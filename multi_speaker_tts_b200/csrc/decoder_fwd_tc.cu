// Tacotron2 decoder loop, forward, teacher-forced: ONE persistent kernel with the two LSTM cells on tcgen05
// tensor cores (bf16x3 split, fp32 accumulators in TMEM), weights re-streamed from L2 by bulk async copies,
// attention memory resident on chip (keys in shared memory, values in TMEM).
//
// Replaces the tf.while_loop body of Modules.py:397-443 and what it calls per step: ZoneoutLSTMCell.call x2
// (ZoneoutLSTMCell.py:228-264), Location_Sensitive_Attention.__call__ (Location_Sensitive_Attention.py:43-85)
// and the AttentionWrapper context.  Same saved-activation layout as the fp32 kernel (decoder_fwd.cu), so
// either reverse kernel can follow it.
//
// Grid: 128 CTAs = 32 clusters x 4, 352 threads:
//   warps 0-7   compute: LSTM epilogues (gates, zoneout), attention, grid barriers
//   warp 8      lane 0: weight-tile producer      (cp.async.bulk, runs ahead of the barriers)
//   warp 9      lane 0: activation-tile producer  (waits for the grid barrier that publishes the data)
//   warp 10     TMEM allocation; lane 0: MMA issuer (tcgen05.mma, commits free the ring slots)
// LSTM partition: cluster c owns units 32c..32c+31 (one M=128 tile of gate rows), CTA r of the cluster owns
// K-slice r of every operand; the four partial accumulators are reduced through distributed shared memory
// and CTA r finishes units 32c+8r..+7.  Per step each CTA runs four MMA jobs in weight-stream order:
//   J0  D0 += W0[ctx rows]  . ctx_{t-1}     (needs the barrier after attention t-1)
//   J1  D1 += W1[m0 rows]   . m0_t          (needs the barrier after cell 0)
//   J2  D0  = W0[h rows]    . h0_t          (for step t+1; off the critical path)
//   J3  D1  = W1[h rows]    . h1_t          (for step t+1; off the critical path)
// Attention: cluster b owns batch row b; CTA r owns attention units 32r..32r+31 and context dims r*D/4...
#include <cooperative_groups.h>

#include "decoder_layout.h"
#include "decoder_tc.cuh"

namespace cg = cooperative_groups;

struct DecFwdTcParams {
  int B, Te, T, D, training, n0;  // n0 = ctx tiles per CTA = D/256
  const uint8_t* wimg;            // [128][n0+12][32 KB]
  uint8_t *ximg_ctx, *ximg_m0, *ximg_h0, *ximg_h1;  // [2 parities][tiles][8 KB]
  const float *b0, *b1, *Wq, *F, *fb, *sw;
  const float *g0pre, *keys, *values;
  const int* text_len;
  const uint8_t* zone_mask;
  float *act0, *act1, *c0n, *c1n, *cz0, *hz0, *cz1, *hz1, *m0, *m1, *ctx, *cum, *align_tm, *qpart, *qf;
  unsigned* barrier;
  long long* dbg;  // [T][32] phase time stamps of CTA 0 (may be null)
  // free-running decode (Modules.py:212-237): projection + prenet inside the loop, stop-token exit
  int infer;
  const float *Wp, *bp, *P0, *pb0, *P1, *pb1;
  const uint8_t* prenet_mask;  // [T,2,B,256]
  uint8_t* ximg_pre;           // [2 parities][4 tiles][8 KB] prenet output of the coming step
  float* proj_tm;              // [T,B,81] bias-free projection (finish_outputs adds the bias)
  int* steps_done;
  int l2_stream;     // saved activations are stored / g0pre loaded with the evict-first hint (MSTTS_LOOP_STREAM=0: off)
  int w_evict_last;  // weight tiles loaded with the L2 evict_last hint (MSTTS_LOOP_L2=1)
};

struct TcSmem {
  // byte offsets into dynamic shared memory
  uint32_t ring, xbuf, recv, keys, wq, xs, m_s, qred, qf_s, cum_s, e_loc, e_parts, a_s, ctx_s, bred, bars, inf, pre2, total;
};
// free-running extras (floats): x_s [256] | pred [8][96] | pparts [4][96] | frame_s [96] | h1_s [256] | p2red [256] | pre_s [64]
constexpr uint32_t kTcInferFloats = 256 + 8 * 96 + 4 * 96 + 96 + 256 + 256 + 64;

__host__ __device__ inline TcSmem tc_fwd_smem(int NS, int Te, int D, int infer = 0) {
  const int TeP = (Te + 31) & ~31;
  TcSmem s;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t o = off;
    off += (bytes + 127) & ~127u;
    return o;
  };
  s.ring = take(NS * kWTileBytes);
  s.xbuf = take(4 * kXTileBytes);  // ONE activation slice: the four jobs of a step read it in turn
  s.recv = take(kDecCluster * kTcN * kRecvStride * 4);
  s.keys = take(Te * 32 * 4);
  s.wq = take(kUnitsPerCta * kAtt * 4);
  s.xs = take(2 * 2 * kTcN * 8 * 2);
  s.m_s = take(kTcN * kUnitsPerCta * 4);
  s.qred = take(8 * 32 * 4);
  s.qf_s = take(32 * 4);
  s.cum_s = take((TeP + 32) * 4);
  s.e_loc = take(TeP * 4);
  s.e_parts = take(kDecCluster * TeP * 4);
  s.a_s = take(TeP * 4);
  s.ctx_s = take((D / kDecCluster) * 4);
  s.bred = take(16 * 4);
  s.bars = take(256);
  s.inf = take(infer ? kTcInferFloats * 4 : 0);
  s.pre2 = take(Te > 128 ? 8 * 16 * 32 * 4 : 0);  // TE2: location features of each warp's second 16-position block
  s.total = off;
  return s;
}

// ---- gate math + zoneout of one (batch, unit): ZoneoutLSTMCell.py:230-264 ----
struct CellOut {
  float ig, jg, fg, og, c, m, cz, hz;
};
__device__ __forceinline__ CellOut cell_forward(const float (&g)[4], float cp, float hp, float mc, float mh) {
  CellOut r;
  r.ig = sigmoidf_precise(g[0]);
  r.jg = tanhf(g[1]);
  r.fg = sigmoidf_precise(g[2] + kForgetBias);
  r.og = sigmoidf_precise(g[3]);
  r.c = r.fg * cp + r.ig * r.jg;
  r.m = r.og * tanhf(r.c);
  r.cz = kZoneKeep * ((r.c - cp) * mc) + cp;
  r.hz = kZoneKeep * ((r.m - hp) * mh) + hp;
  return r;
}

// INFER = 1: free-running decode (a compile-time switch: the training instantiation carries none of its branches)
// TE2 = 1: texts of 129 .. 256 positions.  A CTA cannot hold its D/4 x Te slice of the attention values any more (TMEM has
// 256 free columns x 128 lanes), so TWO clusters serve a batch row (B <= 16): both compute the energies and the softmax over
// all positions -- redundantly and bit-identically, no exchange between them -- and each forms the context for half of the
// dims of its CTAs (values slice: 2 position blocks x D/8 dims).  Training only (the free-running projection needs the whole
// context inside one cluster).
template <int NS, int INFER, int TE2>
__global__ void __cluster_dims__(kDecCluster, 1, 1) __launch_bounds__(kTcThreads, 1)
    decoder_fwd_tc_kernel(const DecFwdTcParams P) {
  static_assert(!(INFER && TE2), "free-running decode runs on one cluster per row");
  extern __shared__ __align__(1024) uint8_t smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = P.B, Te = P.Te, D = P.D, Dq = D / kDecCluster, Dh = Dq / 2;
  const int crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kDecCluster;
  const int arow = TE2 ? (cid & 15) : cid;   // batch row this cluster's attention phase serves
  const int dsel = TE2 ? (cid >> 4) : 0;      // TE2: which half of the CTA's context dims
  const int TeP = (Te + 31) & ~31;
  const int n0 = P.n0;
  constexpr int infer = INFER;
  const int tps = n0 + 12 + infer;  // weight tiles per step (free-running: + the prenet rows of cell 0's kernel)
  const int total_tiles = tps * (P.T - 1) + n0 + infer + 4;
  const TcSmem L = tc_fwd_smem(NS, Te, D, infer);

  uint8_t* ring = smem + L.ring;   // [NS][32 KB] weight tiles
  uint8_t* xbuf = smem + L.xbuf;   // [32 KB]  activation operand of the running job (J0, J1, J2, J3 in turn)
  float* recv = reinterpret_cast<float*>(smem + L.recv);
  float* keys_s = reinterpret_cast<float*>(smem + L.keys);
  float* wq_s = reinterpret_cast<float*>(smem + L.wq);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem + L.xs);  // [vec 2][hi/lo][32][8]
  float* m_s = reinterpret_cast<float*>(smem + L.m_s);
  float* qred = reinterpret_cast<float*>(smem + L.qred);
  float* cum_s = reinterpret_cast<float*>(smem + L.cum_s);
  float* e_parts = reinterpret_cast<float*>(smem + L.e_parts);
  float* a_s = reinterpret_cast<float*>(smem + L.a_s);
  float* ctx_s = reinterpret_cast<float*>(smem + L.ctx_s);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + L.bars);  // [NS]
  uint64_t* empty = wfull + NS;        // [NS]
  uint64_t* xfull = empty + NS;        // [2]
  uint64_t* job_done = xfull + 2;      // [4]
  uint64_t* rs_bar = job_done + 4;     // K-split reduction pushes (st.async complete_tx)
  uint64_t* e_bar = rs_bar + 1;        // partial-energy all-gather
  uint64_t* pp_bar = e_bar + 1;        // free-running: projection partials all-gather
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pp_bar + 1);
  unsigned* ready_seq = tmem_slot + 1;
  unsigned* exit_flag = ready_seq + 1;  // free-running: every row has emitted its stop token -> producers / MMA issuer leave
  float* x_s = reinterpret_cast<float*>(smem + L.inf);  // free-running scratch (see kTcInferFloats)
  float* pred = x_s + 256;
  float* pparts = pred + 8 * 96;
  float* frame_s = pparts + 4 * 96;
  float* h1_s = frame_s + 96;
  float* p2red = h1_s + 256;
  float* pre_s = p2red + 256;
  float* pre2_s = reinterpret_cast<float*>(smem + L.pre2);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&wfull[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&xfull[0], 1);
    ptx::mbar_init(&xfull[1], 1);
    for (int j = 0; j < 4; ++j) ptx::mbar_init(&job_done[j], 1);
    ptx::mbar_init(rs_bar, 1);
    ptx::mbar_init(e_bar, 1);
    ptx::mbar_init(pp_bar, 1);
    *ready_seq = 0;
    *exit_flag = 0;
    ptx::fence_mbar_init();
  }
  if (warp == kTcMmaWarp) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // accumulators (M=64 uses lanes 0-15 of every 32-lane quarter): cell 0 at lane offset 0, cell 1 at lane offset 16,
  // both columns 0..255 (0..127: W_hi rows, 128..255: W_lo rows); values: block h at columns 256 + h*TeP
  const uint32_t tmem_val = tmem + 256;

  // ---- one-time staging by the compute warps ----
  float F_reg[kConvK];
  float sw_l = 0.f, fb_l = 0.f;
  if (warp < 8) {
    const int unit0 = cid * 32 + crank * kUnitsPerCta;
    for (int i = tid; i < kUnitsPerCta * kAtt; i += kTcCompute)
      wq_s[i] = P.Wq[(size_t)(unit0 + i / kAtt) * kAtt + (i % kAtt)];
#pragma unroll
    for (int k = 0; k < kConvK; ++k) F_reg[k] = P.F[k * kAtt + crank * 32 + lane];
    sw_l = P.sw[crank * 32 + lane];
    fb_l = P.fb[crank * 32 + lane];
    for (int i = tid; i < TeP + 32; i += kTcCompute) cum_s[i] = 0.f;
    for (int i = tid; i < TeP; i += kTcCompute) a_s[i] = 0.f;
    if (arow < B) {
      const float* kg = P.keys + (size_t)arow * Te * kAtt;
      for (int i = tid; i < Te * 32; i += kTcCompute) keys_s[i] = kg[(size_t)(i >> 5) * kAtt + crank * 32 + (i & 31)];
      // values slice -> TMEM, lane = context dim, column = text position.  Two blocks (warps 0-3 / 4-7):
      //   TE2 = 0: block = half of the CTA's Dq dims, all TeP positions (columns blk * TeP ..)
      //   TE2 = 1: block = 128-position half, the Dq/2 dims selected by dsel (columns blk * 128 ..)
      const int q = warp & 3, blk = warp >> 2, dloc = q * 32 + lane;
      if (q * 32 < Dh) {
        const float* vg = P.values + (size_t)arow * Te * D + crank * Dq + (TE2 ? dsel : blk) * Dh + dloc;
        const int x0 = TE2 ? blk * 128 : 0, ncol = TE2 ? 128 : TeP;
        for (int c0 = 0; c0 < ncol; c0 += 32) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = (dloc < Dh && x0 + c0 + j < Te) ? __float_as_uint(vg[(size_t)(x0 + c0 + j) * D]) : 0u;
          ptx::tmem_st32(tmem_val + ((uint32_t)(q * 32) << 16) + blk * ncol + c0, v);
        }
        ptx::tmem_wait_st();
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < kConvK; ++k) F_reg[k] = 0.f;
  }
  ptx::tc_fence_before();
  cluster.sync();  // all threads: peers' mbarriers are initialised before any remote store / arrive
  ptx::tc_fence_after();

  if (warp == 8 || warp == 9) {
    // =========================== weight-tile producers (even / odd tiles) ===========================
    // one thread sustains ~80 GB/s of 32 KB bulk copies, two reach the ~150 GB/s an SM can pull from L2
    if (lane == 0) {
      const uint8_t* wsrc = P.wimg + (size_t)blockIdx.x * tps * kWTileBytes;
      for (int i = warp - 8; i < total_tiles; i += 2) {
        const int s = i % NS, round = i / NS;
        if (infer) {
          // free-running: whether step i / tps runs at all is known only after the barrier that ends the previous step, and
          // a tile issued for a step that never runs would land in the shared memory of an exiting CTA -> no prefetch across steps
          const unsigned need = 3u * (unsigned)(i / tps);
          bool stop = false;
          while (ld_volatile_shared(ready_seq) < need) {
            if (ld_volatile_shared(exit_flag)) {
              stop = true;
              break;
            }
          }
          if (stop) break;
        }
        if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
        ptx::mbar_arrive_expect_tx(&wfull[s], kWTileBytes);
        if (P.w_evict_last)
          ptx::bulk_g2s_hint(ring + (size_t)s * kWTileBytes, wsrc + (size_t)(i % tps) * kWTileBytes, kWTileBytes, &wfull[s],
                             ptx::l2_policy_evict_last());
        else
          ptx::bulk_g2s(ring + (size_t)s * kWTileBytes, wsrc + (size_t)(i % tps) * kWTileBytes, kWTileBytes, &wfull[s]);
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // =========================== activation producer: one bulk copy per job ===========================
    if (lane == 0) {
      const size_t ctx_img = (size_t)(D / kTcKT) * kXTileBytes, vec_img = (size_t)(kCell / kTcKT) * kXTileBytes;
      for (int t = 0; t < P.T; ++t) {
        const int njobs = (t == P.T - 1) ? 2 : 4;
        for (int job = 0; job < njobs; ++job) {
          const uint8_t* src;
          uint32_t bytes = 4 * kXTileBytes;
          unsigned need;
          if (job == 0) {  // ctx_{t-1}, K-slice crank
            src = P.ximg_ctx + (size_t)((t + 1) & 1) * ctx_img + (size_t)(crank * n0) * kXTileBytes;
            bytes = (uint32_t)n0 * kXTileBytes;
            need = 3u * t;
          } else if (job == 1) {  // m0_t
            src = P.ximg_m0 + (size_t)(t & 1) * vec_img + (size_t)(crank * 4) * kXTileBytes;
            need = 3u * t + 1;
          } else if (job == 2) {  // h0_t
            src = P.ximg_h0 + (size_t)(t & 1) * vec_img + (size_t)(crank * 4) * kXTileBytes;
            need = 3u * t + 1;
          } else {  // h1_t
            src = P.ximg_h1 + (size_t)(t & 1) * vec_img + (size_t)(crank * 4) * kXTileBytes;
            need = 3u * t + 2;
          }
          // the buffer is free once the previous job has completed (J2 / J3 are deferred work and J0 / J1 wait for a grid
          // barrier that passes long after their predecessor ended, so one buffer costs nothing and pays for a 4th ring slot)
          if (job >= 1) ptx::mbar_wait(&job_done[job - 1], t & 1);
          else if (t > 0) ptx::mbar_wait(&job_done[3], (t - 1) & 1);
          bool stop = false;
          while (ld_volatile_shared(ready_seq) < need) {
            if (infer && ld_volatile_shared(exit_flag)) {
              stop = true;
              break;
            }
          }
          if (stop) goto act_done;
          if (job == 0 && infer) {  // + the K-slice of prenet(frame_{t-1}) behind the context tiles
            ptx::mbar_arrive_expect_tx(&xfull[0], bytes + kXTileBytes);
            ptx::bulk_g2s(xbuf, src, bytes, &xfull[0]);
            ptx::bulk_g2s(xbuf + bytes, P.ximg_pre + (size_t)(t & 1) * (kPrenet / kTcKT) * kXTileBytes + (size_t)crank * kXTileBytes, kXTileBytes,
                          &xfull[0]);
          } else {
            ptx::mbar_arrive_expect_tx(&xfull[job & 1], bytes);
            ptx::bulk_g2s(xbuf, src, bytes, &xfull[job & 1]);
          }
        }
      }
    }
  act_done:
    __syncwarp();
  } else if (warp == kTcMmaWarp) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(64, 256);
      int i = 0;
      for (int t = 0; t < P.T; ++t) {
        const int njobs = (t == P.T - 1) ? 2 : 4;
        for (int job = 0; job < njobs; ++job) {
          const int nt = job == 0 ? n0 + infer : 4;
          // two barriers over the one buffer: even jobs J0, J2, J0, ... ; odd jobs J1, J3, ...
          if (!infer) {
            ptx::mbar_wait(&xfull[job & 1], (uint32_t)(job >> 1));
          } else {
            bool stop = false;
            while (!ptx::mbar_try_wait(&xfull[job & 1], (uint32_t)(job >> 1))) {
              if (ld_volatile_shared(exit_flag)) {
                stop = true;
                break;
              }
            }
            if (stop) goto mma_done;
          }
          const uint32_t d = tmem + ((job & 1) ? (16u << 16) : 0u);
          const uint32_t xbase = ptx::smem_u32(xbuf);
          for (int kt = 0; kt < nt; ++kt, ++i) {
            const int s = i % NS, round = i / NS;
            ptx::mbar_wait(&wfull[s], round & 1);
            ptx::tc_fence_after();
            const uint32_t wbase = ptx::smem_u32(ring + (size_t)s * kWTileBytes);
            const bool fresh = (kt == 0) && (job >= 2 || t == 0);
#pragma unroll
            for (int k = 0; k < kTcKT / 16; ++k) {
              // A = [X_hi ; X_lo] (64 rows), B = [W_hi ; W_lo] (256 rows): hh, hl, lh, ll in one instruction
              const uint64_t a = ptx::umma_desc(xbase + kt * kXTileBytes + k * 256, kTcLBO, kTcSBO);
              const uint64_t bd = ptx::umma_desc(wbase + k * 256, kTcLBO, kTcSBO);
              ptx::umma_bf16(d, a, bd, idesc, (fresh && k == 0) ? 0u : 1u);
            }
            ptx::umma_commit(&empty[s]);
          }
          ptx::umma_commit(&job_done[job]);
          if (P.dbg && blockIdx.x == 0 && job < 2) P.dbg[(size_t)t * 32 + 20 + 5 * job] = clock64();
        }
      }
    }
  mma_done:
    __syncwarp();
  } else {
    // =========================== compute warps ===========================
    const int u8 = tid & 7, b = tid >> 3;  // this thread's (unit, batch row) in both cells
    const int unit = cid * 32 + crank * kUnitsPerCta + u8;
    const bool brow = b < B;
    const size_t BC = (size_t)B * kCell, BG = (size_t)B * kGates;
    const uint32_t recv_addr = ptx::smem_u32(recv);
    const int q4 = warp & 3, half = warp >> 2;  // TMEM lane quarter / accumulator column half of this warp
    uint32_t rs_parity = 0, e_parity = 0, pp_parity = 0;
    int fin_row = 0;  // free-running: warp 0, lane b: row b has emitted stop >= 0
    float c0 = 0.f, h0 = 0.f, c1 = 0.f, h1 = 0.f;  // zoned state of this (batch, unit), AttentionWrapper.zero_state
    const float bias0[4] = {P.b0[unit], P.b0[kCell + unit], P.b0[2 * kCell + unit], P.b0[3 * kCell + unit]};
    const float bias1[4] = {P.b1[unit], P.b1[kCell + unit], P.b1[2 * kCell + unit], P.b1[3 * kCell + unit]};
    float wq_r[kUnitsPerCta];  // query_layer rows of this CTA's units, column tid & 127
#pragma unroll
    for (int u = 0; u < kUnitsPerCta; ++u) wq_r[u] = wq_s[u * kAtt + (tid & 127)];
    unsigned bar_target = 0;
    const int tl = (arow < B) ? min(P.text_len[arow], Te) : 0;
    const size_t vec_img = (size_t)(kCell / kTcKT) * kXTileBytes, ctx_img = (size_t)(D / kTcKT) * kXTileBytes;
    // byte offset of this CTA's 8-unit chunk inside a [32 x 1024] activation image, for batch row r: + (r/8)*1024 + (r%8)*16
    const int unit0 = cid * 32 + crank * kUnitsPerCta;
    const size_t img_chunk = (size_t)(unit0 >> 6) * kXTileBytes + (size_t)((unit0 & 63) >> 3) * 128;
    long long* dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
    // at the middle step EVERY CTA also stamps the global timer (rows 0..127 of the same buffer, one row per CTA) so that
    // tools/phase_times.py can show the arrival spread at each barrier
    long long* dbg_all = (P.dbg && tid == 0) ? P.dbg + (size_t)blockIdx.x * 32 : nullptr;
#define STAMP(k)                                                 \
  do {                                                           \
    if (dbg) dbg[(size_t)t * 32 + (k)] = clock64();              \
    if (dbg_all && t == P.T / 2) {                               \
      unsigned long long gt;                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));     \
      dbg_all[(k)] = (long long)gt;                              \
    }                                                            \
  } while (0)

    // Pulls this CTA's partial accumulator (cell = job & 1) out of TMEM, adds its four hi/lo quadrants and scatters
    // the 128 gate rows to the CTAs that own the units.  X-image rows are ordered so that TMEM lane quarter q holds
    // batch rows 8q..8q+7: x_hi rows on lanes 0-7, x_lo rows on lanes 8-15 (cell 1: +16).  Warp w reads quarter
    // q4 = w & 3 and gate-row columns 64*half .. +63 (W_hi) and 128+64*half .. (W_lo).
    auto reduce_scatter = [&](int job, int t) {
      if (tid == 0) ptx::mbar_arrive_expect_tx(rs_bar, kDecCluster * 32 * kTcN * 4);
      mbar_wait_warp(&job_done[job], t & 1);
      STAMP(1 + 4 * job);
      ptx::tc_fence_after();
      const int l16 = lane - ((job & 1) ? 16 : 0);
      const bool pusher = l16 >= 0 && l16 < 8;          // lanes holding the x_hi rows of this cell's accumulator
      const int bq = 8 * q4 + (l16 & 7);                // batch row of this lane
      const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16) + half * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t vh[32], vl[32];
        ptx::tmem_ld32(ta + ch * 32, vh);
        ptx::tmem_ld32(ta + 128 + ch * 32, vl);
        ptx::tmem_wait_ld();
        const int rr = half * 2 + ch;  // 32 gate rows = the units of cluster CTA rr
        const uint32_t dst = ptx::mapa(recv_addr, (uint32_t)rr) + (uint32_t)(((crank * kTcN + bq) * kRecvStride) * 4);
        const uint32_t rbar = ptx::mapa(ptx::smem_u32(rs_bar), (uint32_t)rr);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = __uint_as_float(vh[j + e]) + __uint_as_float(vl[j + e]);
            o[e] = v + __shfl_down_sync(0xffffffffu, v, 8);  // + the x_lo row of the same batch
          }
          if (pusher)
            ptx::st_async_v4(dst + j * 4, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]), rbar);
        }
      }
      ptx::tc_fence_before();
      mbar_wait_warp(rs_bar, rs_parity);
      rs_parity ^= 1u;
    };

    // keys + location features of the NEXT attention step: depends only on the cumulative alignment, so it runs
    // inside the barrier wait that follows the softmax (Location_Sensitive_Attention.py:48-61,82)
    // one 16-position block per warp (positions 16 w ..); TE2: a second block (128 + 16 w ..) whose values wait in shared
    // memory (pre2_s [warp][16][32]) instead of registers
    float pre_e[16];
    auto location_block = [&](int t0, float (&out)[16]) {
#pragma unroll
      for (int p = 0; p < 16; ++p) out[p] = 0.f;
      if (t0 < tl) {
#pragma unroll
        for (int c = 0; c < 16 + kConvK - 1; ++c) {
          const float cv = cum_s[t0 + c];
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int k = c - p;
            if (k >= 0 && k < kConvK) out[p] = fmaf(cv, F_reg[k], out[p]);
          }
        }
#pragma unroll
        for (int p = 0; p < 16; ++p)
          if (t0 + p < tl) out[p] += keys_s[(t0 + p) * 32 + lane];
      }
    };
    auto location_features = [&]() {
      if (TE2) {
        float tmp[16];
        location_block(128 + warp * 16, tmp);
#pragma unroll
        for (int p = 0; p < 16; ++p) pre2_s[(warp * 16 + p) * 32 + lane] = tmp[p];
      }
      location_block(warp * 16, pre_e);
    };
    ptx::bar_sync(1, kTcCompute);
    location_features();

    for (int t = 0; t < P.T; ++t) {
      const uint8_t* zm = P.training ? P.zone_mask + (size_t)t * 4 * BC : nullptr;  // inference: mask = 1, the (1 - r) factor stays
      const int par = t & 1;
      STAMP(0);
      // ================= phase A: LSTM cell 0 =================
      {
        float add[4] = {0.f, 0.f, 0.f, 0.f};
        float mc = 1.f, mh = 1.f;
        if (brow) {
          if (!infer) {
            const float* gp = P.g0pre + (size_t)t * BG + (size_t)b * kGates + unit;
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) add[gi] = ld_stream(gp + gi * kCell, P.l2_stream) + bias0[gi];
          } else {  // the prenet rows of cell 0's kernel are part of job J0
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) add[gi] = bias0[gi];
          }
          if (zm) {
            mc = (float)zm[(size_t)b * kCell + unit];
            mh = (float)zm[BC + (size_t)b * kCell + unit];
          }
        }
        reduce_scatter(0, t);
        STAMP(2);
        float g[4];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          float s = 0.f;
#pragma unroll
          for (int src = 0; src < kDecCluster; ++src) s += recv[(src * kTcN + b) * kRecvStride + gi * 8 + u8];
          g[gi] = s + add[gi];
        }
        const CellOut r = cell_forward(g, c0, h0, mc, mh);
        if (brow) {
          __nv_bfloat16 hi, lo;
          c0 = r.cz;
          h0 = r.hz;
          split_bf16(r.m, hi, lo);
          xs[(0 * 2 + 0) * 256 + b * 8 + u8] = hi;
          xs[(0 * 2 + 1) * 256 + b * 8 + u8] = lo;
          split_bf16(r.hz, hi, lo);
          xs[(1 * 2 + 0) * 256 + b * 8 + u8] = hi;
          xs[(1 * 2 + 1) * 256 + b * 8 + u8] = lo;
        }
        ptx::bar_sync(1, kTcCompute);
        if (tid < 128) {  // 16-byte rows of the m0 / h0 images: (vector, hi|lo, batch row)
          const int vec = tid >> 6, hl = (tid >> 5) & 1, r2 = tid & 31;
          if (r2 < B) {
            uint8_t* img = (vec ? P.ximg_h0 : P.ximg_m0) + (size_t)par * vec_img + img_chunk + ximg_row_offset(r2, hl);
            *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(xs + (vec * 2 + hl) * 256 + r2 * 8);
          }
        }
        STAMP(3);
        grid_arrive_compute(P.barrier, bar_target, gridDim.x);
        if (brow && !infer) {  // activations saved for the reverse pass: nobody waits for these inside the loop
          const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
          st_stream(P.act0 + ai, r.ig, P.l2_stream);
          st_stream(P.act0 + ai + kCell, r.jg, P.l2_stream);
          st_stream(P.act0 + ai + 2 * kCell, r.fg, P.l2_stream);
          st_stream(P.act0 + ai + 3 * kCell, r.og, P.l2_stream);
          st_stream(P.c0n + (size_t)t * BC + si, r.c, P.l2_stream);
          st_stream(P.cz0 + (size_t)(t + 1) * BC + si, r.cz, P.l2_stream);
          st_stream(P.hz0 + (size_t)(t + 1) * BC + si, r.hz, P.l2_stream);
          st_stream(P.m0 + (size_t)t * BC + si, r.m, P.l2_stream);
        }
        grid_wait_compute(P.barrier, bar_target, ready_seq, 3u * t + 1);
      }
      STAMP(4);

      // ================= phase B: LSTM cell 1 (+ partial query projection) =================
      {
        float mc = 1.f, mh = 1.f;
        if (brow && zm) {
          mc = (float)zm[2 * BC + (size_t)b * kCell + unit];
          mh = (float)zm[3 * BC + (size_t)b * kCell + unit];
        }
        reduce_scatter(1, t);
        STAMP(6);
        float g[4];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          float s = 0.f;
#pragma unroll
          for (int src = 0; src < kDecCluster; ++src) s += recv[(src * kTcN + b) * kRecvStride + gi * 8 + u8];
          g[gi] = s + bias1[gi];
        }
        const CellOut r = cell_forward(g, c1, h1, mc, mh);
        m_s[b * kUnitsPerCta + u8] = brow ? r.m : 0.f;
        if (brow) {
          __nv_bfloat16 hi, lo;
          c1 = r.cz;
          h1 = r.hz;
          split_bf16(r.hz, hi, lo);
          xs[(1 * 2 + 0) * 256 + b * 8 + u8] = hi;
          xs[(1 * 2 + 1) * 256 + b * 8 + u8] = lo;
        }
        ptx::bar_sync(1, kTcCompute);
        if (tid < 64) {
          const int hl = tid >> 5, r2 = tid & 31;
          if (r2 < B) {
            uint8_t* img = P.ximg_h1 + (size_t)par * vec_img + img_chunk + ximg_row_offset(r2, hl);
            *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(xs + (2 + hl) * 256 + r2 * 8);
          }
        }
        // partial q[b][a] = sum over this CTA's 8 units of m1[b][unit] * Wq[unit][a]; thread = (a, half of the batch)
        {
          const int a = tid & 127;
#pragma unroll 4
          for (int bb = (tid >> 7) * 16; bb < (tid >> 7) * 16 + 16; ++bb) {
            if (bb < B) {
              const float4 ma = *reinterpret_cast<const float4*>(m_s + bb * kUnitsPerCta);
              const float4 mb = *reinterpret_cast<const float4*>(m_s + bb * kUnitsPerCta + 4);
              float s = ma.x * wq_r[0];
              s = fmaf(ma.y, wq_r[1], s);
              s = fmaf(ma.z, wq_r[2], s);
              s = fmaf(ma.w, wq_r[3], s);
              s = fmaf(mb.x, wq_r[4], s);
              s = fmaf(mb.y, wq_r[5], s);
              s = fmaf(mb.z, wq_r[6], s);
              s = fmaf(mb.w, wq_r[7], s);
              P.qpart[((size_t)blockIdx.x * B + bb) * kAtt + a] = s;
            }
          }
        }
        STAMP(7);
        // free-running: the projection of phase C reads every unit of m1_t from global memory -> store it before the barrier
        if (infer && brow) P.m1[(size_t)t * BC + (size_t)b * kCell + unit] = r.m;
        grid_arrive_compute(P.barrier, bar_target, gridDim.x);
        if (brow && !infer) {
          const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
          st_stream(P.act1 + ai, r.ig, P.l2_stream);
          st_stream(P.act1 + ai + kCell, r.jg, P.l2_stream);
          st_stream(P.act1 + ai + 2 * kCell, r.fg, P.l2_stream);
          st_stream(P.act1 + ai + 3 * kCell, r.og, P.l2_stream);
          st_stream(P.c1n + (size_t)t * BC + si, r.c, P.l2_stream);
          st_stream(P.cz1 + (size_t)(t + 1) * BC + si, r.cz, P.l2_stream);
          st_stream(P.hz1 + (size_t)(t + 1) * BC + si, r.hz, P.l2_stream);
          st_stream(P.m1 + (size_t)t * BC + si, r.m, P.l2_stream);
        }
        grid_wait_compute(P.barrier, bar_target, ready_seq, 3u * t + 2);
      }
      STAMP(8);

      // ================= phase C: location-sensitive attention, batch row = cluster index ==========
      float ctx_keep = 0.f;
      if (arow < B) {
        const int bb = arow;
        {  // q slice = sum of the 128 per-CTA partials (fixed order)
          float pq[kDecGrid / 8];
#pragma unroll
          for (int j = 0; j < kDecGrid / 8; ++j)
            pq[j] = __ldcg(P.qpart + ((size_t)(warp + 8 * j) * B + bb) * kAtt + crank * 32 + lane);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < kDecGrid / 8; ++j) s += pq[j];
          qred[warp * 32 + lane] = s;
        }
        ptx::bar_sync(1, kTcCompute);
        float qf = fb_l;
#pragma unroll
        for (int w = 0; w < 8; ++w) qf += qred[w * 32 + lane];
        if (warp == 0 && dsel == 0) P.qf[((size_t)t * B + bb) * kAtt + crank * 32 + lane] = qf;
        STAMP(14);
#pragma unroll
        for (int blk = 0; blk < (TE2 ? 2 : 1); ++blk) {
          const int t0 = blk * 128 + warp * 16;
          if (t0 < tl) {
            float v[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float pe = blk ? pre2_s[(warp * 16 + p) * 32 + lane] : pre_e[p];
              v[p] = (t0 + p < tl) ? sw_l * tanhf(pe + qf) : 0.f;
            }
            // sum over the 32 lanes (attention units) of 16 independent values: halve the value count at every
            // butterfly step (31 shuffles instead of 80, all independent within a level)
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
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            // lane L now holds position ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1) of the block:
            // push it straight to the four CTAs of the cluster (st.async counts the bytes on their e_bar)
            // lanes L, L+2, L+4, L+6 (L % 8 == 0) hold four consecutive positions: gathered into one 16-byte remote store
            // (scalar remote stores are disproportionately expensive: zlstm.cu measured it)
            const float a1 = __shfl_down_sync(0xffffffffu, v[0], 2), a2 = __shfl_down_sync(0xffffffffu, v[0], 4),
                        a3 = __shfl_down_sync(0xffffffffu, v[0], 6);
            if ((lane & 7) == 0) {
              const int x = t0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4;
              if (x < tl) {
                const uint32_t ep = ptx::smem_u32(e_parts) + (uint32_t)((crank * TeP + x) * 4);
                const uint32_t eb = ptx::smem_u32(e_bar);
#pragma unroll
                for (uint32_t dst = 0; dst < (uint32_t)kDecCluster; ++dst)
                  ptx::st_async_v4(ptx::mapa(ep, dst), __float_as_uint(v[0]), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3),
                                   ptx::mapa(eb, dst));
              }
            }
          }
        }
        if (tid == 0) ptx::mbar_arrive_expect_tx(e_bar, (uint32_t)(kDecCluster * ((tl + 3) & ~3) * 4));  // whole groups of four
        STAMP(11);
        mbar_wait_warp(e_bar, e_parity);
        e_parity ^= 1u;
        STAMP(12);
        // masked softmax over positions < tl (score_mask_value = -inf => exactly 0 beyond tl); Te <= 128: every
        // warp redundantly reduces all positions (4 per lane), no block-level exchange
        constexpr int NJ = TE2 ? 8 : 4;  // positions per lane
        float ev[NJ];
        float lmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int x = lane + 32 * j;
          ev[j] = -INFINITY;
          if (x < tl) ev[j] = ((e_parts[x] + e_parts[TeP + x]) + e_parts[2 * TeP + x]) + e_parts[3 * TeP + x];
          lmax = fmaxf(lmax, ev[j]);
        }
        lmax = warp_max(lmax);
        float lsum = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          ev[j] = (lane + 32 * j < tl) ? expf(ev[j] - lmax) : 0.f;
          lsum += ev[j];
        }
        // fixed-order sum (same in every warp and every CTA of the cluster)
        lsum = warp_sum(lsum);
        if (warp == 0) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int x = lane + 32 * j;
            if (x < Te) {
              const float a = ev[j] / lsum;
              a_s[x] = a;
              cum_s[15 + x] += a;
            }
          }
        }
        ptx::bar_sync(1, kTcCompute);
        STAMP(13);
        // context slice from the TMEM-resident values: thread = one context dim, columns = text positions
        if (q4 * 32 < Dh) {
          const int dloc = q4 * 32 + lane;
          float s = 0.f;
          // TE2 = 0: block `half` = half of the dims over all positions; TE2 = 1: block `half` = positions 128 half .. of dims dsel
          const int xb = TE2 ? half * 128 : 0, xe = TE2 ? min(tl, xb + 128) : tl;
          const uint32_t va = tmem_val + ((uint32_t)(q4 * 32) << 16) + half * (TE2 ? 128 : TeP);
          for (int c0 = 0; xb + c0 < xe; c0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(va + c0, v);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) s = fmaf(a_s[xb + c0 + j], __uint_as_float(v[j]), s);
          }
          if (dloc < Dh) ctx_s[half * Dh + dloc] = s;   // TE2: the two position-half partial sums of dim dloc
          ctx_keep = s;
        }
        ptx::bar_sync(1, kTcCompute);
        if (TE2) {  // combine the two position halves (fixed order); ctx_s[0 .. Dh) = this CTA's context dims
          float tot = 0.f;
          if (tid < Dh) tot = ctx_s[tid] + ctx_s[Dh + tid];
          ptx::bar_sync(1, kTcCompute);
          if (tid < Dh) ctx_s[tid] = tot;
          ptx::bar_sync(1, kTcCompute);
        }
        constexpr int kImgDiv = TE2 ? 8 : 4;  // dims of this CTA in the image: Dq (TE2: Dq / 2)
        if (tid < Dq / kImgDiv) {  // 16-byte rows of the ctx image: (hi|lo, 8-dim chunk)
          const int hl = tid / (Dq / (2 * kImgDiv)), ch = tid % (Dq / (2 * kImgDiv));
          const int k = crank * Dq + dsel * Dh + ch * 8;
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat16 hi, lo;
            split_bf16(ctx_s[ch * 8 + j], hi, lo);
            o[j] = hl ? lo : hi;
          }
          uint8_t* img = P.ximg_ctx + (size_t)par * ctx_img + (size_t)(k >> 6) * kXTileBytes + (size_t)((k & 63) >> 3) * 128 +
                         ximg_row_offset(bb, hl);
          *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(o);
        }
        if (infer) {
          // ---- free running (Modules.py:212-237,309-321): frame_t = [m1_t | ctx_t] . Wp + bp, K split over the cluster;
          //      then prenet(frame_t) = the decoder input of step t + 1 (dropout stays on, Modules.py:252) ----
          x_s[tid] = __ldcg(P.m1 + ((size_t)t * B + bb) * kCell + crank * 256 + tid);
          if (tid == 0) ptx::mbar_arrive_expect_tx(pp_bar, kDecCluster * (kMel + 1) * 4);
          ptx::bar_sync(1, kTcCompute);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
          for (int r = warp; r < 256 + Dq; r += 8) {
            const int grow = r < 256 ? crank * 256 + r : kCell + crank * Dq + (r - 256);
            const float* wr = P.Wp + (size_t)grow * (kMel + 1);
            const float xv = r < 256 ? x_s[r] : ctx_s[r - 256];
            a0 = fmaf(xv, __ldg(wr + lane), a0);
            a1 = fmaf(xv, __ldg(wr + 32 + lane), a1);
            if (lane < kMel + 1 - 64) a2 = fmaf(xv, __ldg(wr + 64 + lane), a2);
          }
          pred[warp * 96 + lane] = a0;
          pred[warp * 96 + 32 + lane] = a1;
          pred[warp * 96 + 64 + lane] = a2;
          ptx::bar_sync(1, kTcCompute);
          if (tid < kMel + 1) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sacc += pred[w * 96 + tid];
            const uint32_t pa = ptx::smem_u32(pparts) + (uint32_t)((crank * 96 + tid) * 4);
            const uint32_t pb = ptx::smem_u32(pp_bar);
#pragma unroll
            for (uint32_t dst = 0; dst < (uint32_t)kDecCluster; ++dst) ptx::st_async_f32(ptx::mapa(pa, dst), sacc, ptx::mapa(pb, dst));
          }
          mbar_wait_warp(pp_bar, pp_parity);
          pp_parity ^= 1u;
          if (tid < kMel + 1) {
            const float sacc = ((pparts[tid] + pparts[96 + tid]) + pparts[192 + tid]) + pparts[288 + tid];
            if (crank == 0) P.proj_tm[((size_t)t * B + bb) * (kMel + 1) + tid] = sacc;  // bias-free; read by every CTA after the barrier
            frame_s[tid] = sacc + P.bp[tid];
          }
          ptx::bar_sync(1, kTcCompute);
          if (t + 1 < P.T) {
            const uint8_t* pm = P.prenet_mask + (size_t)(t + 1) * 2 * B * kPrenet;
            {  // layer 0: every CTA of the cluster computes all 256 outputs (80 x 256)
              float sacc = P.pb0[tid];
              for (int k = 0; k < kMel; ++k) sacc = fmaf(frame_s[k], __ldg(P.P0 + k * kPrenet + tid), sacc);
              h1_s[tid] = (fmaxf(sacc, 0.f) / 0.5f) * (float)pm[(size_t)bb * kPrenet + tid];
            }
            ptx::bar_sync(1, kTcCompute);
            {  // layer 1: CTA r computes outputs 64 r .. 64 r + 63 = its K-slice of the next step's job J0
              const int j = crank * 64 + (tid & 63), kq = tid >> 6;
              float sacc = 0.f;
              for (int k = kq * 64; k < kq * 64 + 64; ++k) sacc = fmaf(h1_s[k], __ldg(P.P1 + k * kPrenet + j), sacc);
              p2red[kq * 64 + (tid & 63)] = sacc;
            }
            ptx::bar_sync(1, kTcCompute);
            if (tid < 64) {
              const int j = crank * 64 + tid;
              const float sacc = ((p2red[tid] + p2red[64 + tid]) + p2red[128 + tid]) + p2red[192 + tid] + P.pb1[j];
              pre_s[tid] = (fmaxf(sacc, 0.f) / 0.5f) * (float)pm[((size_t)B + bb) * kPrenet + j];
            }
            ptx::bar_sync(1, kTcCompute);
            if (tid < 16) {  // 16-byte rows of the pre image: (hi|lo, 8-value chunk)
              const int hl = tid >> 3, ch = tid & 7;
              __align__(16) __nv_bfloat16 o[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                __nv_bfloat16 hi, lo;
                split_bf16(pre_s[ch * 8 + jj], hi, lo);
                o[jj] = hl ? lo : hi;
              }
              uint8_t* img = P.ximg_pre + (size_t)((t + 1) & 1) * (kPrenet / kTcKT) * kXTileBytes + (size_t)crank * kXTileBytes +
                             (size_t)ch * 128 + ximg_row_offset(bb, hl);
              *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(o);
            }
          }
        }
      }
      STAMP(9);
      grid_arrive_compute(P.barrier, bar_target, gridDim.x);
      if (arow < B) {  // saved outputs + next step's location features ride in the barrier wait
        const int bb = arow;
        if (!TE2) {
          if (q4 * 32 < Dh && q4 * 32 + lane < Dh)
            P.ctx[((size_t)(t + 1) * B + bb) * D + crank * Dq + half * Dh + q4 * 32 + lane] = ctx_keep;
        } else if (tid < Dh) {
          P.ctx[((size_t)(t + 1) * B + bb) * D + crank * Dq + dsel * Dh + tid] = ctx_s[tid];
        }
        if (crank == 0 && dsel == 0) {
          for (int x = tid; x < Te; x += kTcCompute) {
            P.align_tm[((size_t)t * B + bb) * Te + x] = a_s[x];
            P.cum[((size_t)(t + 1) * B + bb) * Te + x] = cum_s[15 + x];
          }
        }
        location_features();
      }
      if (!infer) {
        grid_wait_compute(P.barrier, bar_target, ready_seq, 3u * t + 3);
      } else {
        // wait WITHOUT publishing the event: whether a next step exists is decided first.  finished |= stop >= 0
        // (Modules.py:216-219, OR-ed at :409); every CTA evaluates the same data => uniform exit.  The step cap
        // (time >= Max_Inference_Length) is the loop bound T = cap + 1.
        if ((tid & 31) == 0 && tid < 128) {
          while (ld_acquire_gpu(P.barrier) < bar_target) {
          }
        }
        ptx::bar_sync(1, kTcCompute);
        if (warp == 0) {
          if (lane < B) {
            const float st = __ldcg(P.proj_tm + ((size_t)t * B + lane) * (kMel + 1) + kMel) + P.bp[kMel];
            if (st >= 0.f) fin_row = 1;
          }
          const int all = __all_sync(0xffffffffu, lane < B ? fin_row : 1);
          if (lane == 0) {
            if (blockIdx.x == 0) *P.steps_done = t + 1;
            if (all) st_volatile_shared(exit_flag, 1u);
            else st_volatile_shared(ready_seq, 3u * t + 3);
          }
        }
        ptx::bar_sync(1, kTcCompute);
        if (ld_volatile_shared(exit_flag)) {
          // the deferred jobs J2 / J3 of this step were issued already: let them finish before the accumulators go away
          if (t < P.T - 1) {
            mbar_wait_warp(&job_done[2], t & 1);
            mbar_wait_warp(&job_done[3], t & 1);
          }
          break;
        }
      }
      STAMP(10);
    }
#undef STAMP
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kTcMmaWarp) ptx::tmem_dealloc(tmem, 512);
  cluster.sync();  // no CTA exits while a peer could still address its shared memory
}

// ---- weight image: bf16 hi/lo tiles in stream order for every CTA ----------------------------------------
// tile rows i = rr*32 + g*8 + u  <->  gate column g*1024 + 32c + 8rr + u ; k runs over this CTA's K-slice
__global__ void prep_wimg_fwd_kernel(const float* __restrict__ K0, const float* __restrict__ K1, uint8_t* __restrict__ wimg,
                                     int D, int infer) {
  const int n0 = D / 256, tps = n0 + 12 + infer, Dq = D / kDecCluster;
  const size_t total = (size_t)kDecGrid * tps * 8 * 128;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx & 127);
    const int kc = (int)((idx >> 7) & 7);
    const size_t tq = idx >> 10;
    const int q = (int)(tq % tps), cta = (int)(tq / tps);
    const int c = cta >> 2, r = cta & 3;
    const int rr = i >> 5, g = (i >> 3) & 3, u = i & 7;
    const int col = g * kCell + c * 32 + rr * 8 + u;
    float v[8];
    if (q < n0) {
      const int k = r * Dq + q * kTcKT + kc * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = K0[(size_t)(kPrenet + k + j) * kGates + col] + K0[(size_t)(kPrenet + D + k + j) * kGates + col];
    } else if (infer && q == n0) {  // free-running: prenet rows of cell 0's kernel, K-slice r (64 of the 256 rows)
      const int k = r * kTcKT + kc * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = K0[(size_t)(k + j) * kGates + col];
    } else {
      const int jq = (q - n0 - infer) >> 2, kt = (q - n0 - infer) & 3;
      const int k = r * 256 + kt * kTcKT + kc * 8;
      const float* src = jq == 0 ? K1 + (size_t)k * kGates
                                 : (jq == 1 ? K0 + (size_t)(kPrenet + 2 * D + k) * kGates : K1 + (size_t)(kCell + k) * kGates);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = src[(size_t)j * kGates + col];
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(v[j], hi[j], lo[j]);
    uint8_t* tile = wimg + tq * kWTileBytes + (size_t)(i >> 3) * 1024 + (size_t)kc * 128 + (size_t)(i & 7) * 16;
    *reinterpret_cast<uint4*>(tile) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(tile + kWTileBytes / 2) = *reinterpret_cast<const uint4*>(lo);
  }
}

// free-running decode: prenet of the all-zero first frame for batch row blockIdx.x -> parity-0 operand image of job J0
__global__ void prenet_zero_frame_kernel(const float* __restrict__ pb0, const float* __restrict__ P1, const float* __restrict__ pb1,
                                         const uint8_t* __restrict__ mask, int B, uint8_t* __restrict__ ximg_pre) {
  __shared__ float h1[kPrenet], pre[kPrenet];
  const int b = blockIdx.x, j = threadIdx.x;
  h1[j] = (fmaxf(pb0[j], 0.f) / 0.5f) * (float)mask[(size_t)b * kPrenet + j];
  __syncthreads();
  float sacc = pb1[j];
  for (int k = 0; k < kPrenet; ++k) sacc = fmaf(h1[k], P1[k * kPrenet + j], sacc);
  pre[j] = (fmaxf(sacc, 0.f) / 0.5f) * (float)mask[((size_t)B + b) * kPrenet + j];
  __syncthreads();
  if (j < 64) {  // (hi|lo, 8-value chunk): 2 x 32 chunks of 16 bytes
    const int hl = j >> 5, ch = j & 31;
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      __nv_bfloat16 hi, lo;
      split_bf16(pre[ch * 8 + jj], hi, lo);
      o[jj] = hl ? lo : hi;
    }
    const int k = ch * 8;
    uint8_t* img = ximg_pre + (size_t)(k >> 6) * kXTileBytes + (size_t)((k & 63) >> 3) * 128 + ximg_row_offset(b, hl);
    *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(o);
  }
}

// ======================================== host side ================================================
// one launch: 32 rows of texts up to 128 positions, or 16 rows of texts up to 256 (two clusters per row); larger batches run
// as row chunks (decoder_layout.h: DecChunkPlan)
bool dec_tc_supported(int B, int Te, int D) {
  return B >= 1 && Te >= 1 && Te <= 256 && B <= (Te > 128 ? kTcN / 2 : kTcN) && D % 256 == 0 && D <= 768;
}

template <int NS, int INFER, int TE2>
static int launch_fwd_tc(const DecFwdTcParams& P, cudaStream_t stream, size_t smem, bool* ok) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  int max_optin = 0;
  MSTTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem > (size_t)max_optin) {
    *ok = false;
    return MSTTS_OK;
  }
  *ok = true;
  MSTTS_CUDA(cudaFuncSetAttribute(decoder_fwd_tc_kernel<NS, INFER, TE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  MSTTS_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, decoder_fwd_tc_kernel<NS, INFER, TE2>, &cfg));
  MSTTS_REQUIRE(nclusters * kDecCluster >= kDecGrid, MSTTS_E_DEVICE,
                "decoder_fwd_tc: device co-schedules only %d clusters of %d (need %d)", nclusters, kDecCluster,
                kDecGrid / kDecCluster);
  mstts_timer_start(0, stream);
  MSTTS_CUDA(cudaLaunchKernelEx(&cfg, decoder_fwd_tc_kernel<NS, INFER, TE2>, P));
  mstts_timer_stop(0, stream);
  return MSTTS_OK;
}

int dec_fwd_tc_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws, cudaStream_t s) {
  auto F = [&](size_t off) { return (float*)(ws + off); };
  const int D = io->D;
  MSTTS_REQUIRE(dec_tc_supported(io->B, io->Te, D), MSTTS_E_UNSUPPORTED,
                "decoder bf16x3 mode needs Te<=256, D in {256,512,768} and B<=32 (16 for Te>128) per chunk (got B=%d Te=%d D=%d); use mode fp32",
                io->B, io->Te, D);
  DecFwdTcParams P;
  memset(&P, 0, sizeof(P));
  P.B = io->B; P.Te = io->Te; P.T = io->n_steps; P.D = D; P.training = io->is_training; P.n0 = D / 256;
  P.wimg = (const uint8_t*)(ws + l.wimg_f);
  P.ximg_ctx = (uint8_t*)(ws + l.ximg_ctx); P.ximg_m0 = (uint8_t*)(ws + l.ximg_m0);
  P.ximg_h0 = (uint8_t*)(ws + l.ximg_h0); P.ximg_h1 = (uint8_t*)(ws + l.ximg_h1);
  P.b0 = w->cell0_bias; P.b1 = w->cell1_bias; P.Wq = w->query_kernel;
  P.F = F(l.locF); P.fb = F(l.locFb); P.sw = w->score_w;
  P.g0pre = F(l.g0pre); P.keys = F(l.keys); P.values = F(l.values);
  {
    const char* e = getenv("MSTTS_LOOP_STREAM");
    P.l2_stream = e ? atoi(e) : 1;
    e = getenv("MSTTS_LOOP_L2");
    P.w_evict_last = e ? atoi(e) : 0;
  }
  P.text_len = io->text_len; P.zone_mask = io->zone_mask;
  P.act0 = F(l.act0); P.act1 = F(l.act1); P.c0n = F(l.c0n); P.c1n = F(l.c1n);
  P.cz0 = F(l.cz0); P.hz0 = F(l.hz0); P.cz1 = F(l.cz1); P.hz1 = F(l.hz1);
  P.m0 = F(l.m0); P.m1 = F(l.m1); P.ctx = F(l.ctx); P.cum = F(l.cum); P.align_tm = F(l.align_tm); P.qpart = F(l.qpart); P.qf = F(l.qf);
  P.barrier = (unsigned*)(ws + l.barrier);
  P.dbg = (long long*)(ws + l.dbg);
  // weight image (weights change every optimiser step) and zeroed activation images (initial state = 0)
  const int infer = io->is_training ? 0 : 1;
  prep_wimg_fwd_kernel<<<148 * 8, 256, 0, s>>>(w->cell0_kernel, w->cell1_kernel, (uint8_t*)(ws + l.wimg_f), D, infer);
  MSTTS_CUDA(cudaMemsetAsync(ws + l.ximg_ctx, 0, l.ximg_end - l.ximg_ctx, s));
  if (infer) {
    P.infer = 1;
    P.Wp = w->proj_kernel; P.bp = w->proj_bias; P.P0 = w->prenet0_kernel; P.pb0 = w->prenet0_bias;
    P.P1 = w->prenet1_kernel; P.pb1 = w->prenet1_bias; P.prenet_mask = io->prenet_mask;
    P.ximg_pre = (uint8_t*)(ws + l.ximg_pre); P.proj_tm = F(l.proj_tm); P.steps_done = io->steps_done;
    // decoder input of step 0 = prenet(zero frame) (Modules.py:178-185), straight into the operand image
    prenet_zero_frame_kernel<<<io->B, kPrenet, 0, s>>>(w->prenet0_bias, w->prenet1_kernel, w->prenet1_bias, io->prenet_mask, io->B, P.ximg_pre);
  }
  bool ok = false;
  int rc;
#define MSTTS_TRY_NS(NS_)                                                                                                   \
  if (!ok) {                                                                                                                 \
    rc = infer ? launch_fwd_tc<NS_, 1, 0>(P, s, tc_fwd_smem(NS_, io->Te, D, 1).total, &ok)                                   \
               : (io->Te > 128 ? launch_fwd_tc<NS_, 0, 1>(P, s, tc_fwd_smem(NS_, io->Te, D, 0).total, &ok)                   \
                               : launch_fwd_tc<NS_, 0, 0>(P, s, tc_fwd_smem(NS_, io->Te, D, 0).total, &ok));                 \
    if (rc) return rc;                                                                                                       \
  }
  MSTTS_TRY_NS(4)  // all 4 weight tiles of J1 resident when m0 arrives
  MSTTS_TRY_NS(3)
  MSTTS_TRY_NS(2)
#undef MSTTS_TRY_NS
  MSTTS_REQUIRE(ok, MSTTS_E_UNSUPPORTED, "decoder_fwd_tc: shared memory does not fit for Te=%d", io->Te);
  return MSTTS_OK;
}
