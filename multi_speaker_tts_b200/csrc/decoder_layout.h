// Workspace layout of the decoder op: prepared weights, hoisted pre-GEMM results and every activation
// the reverse pass needs.  One function computes the offsets so fwd, bwd and workspace_bytes agree.
#pragma once
#include "common.cuh"

struct DecLayout {
  // ---- prepared operands (rebuilt every fwd call: weights change each optimiser step) ----
  size_t values;   // [B,Te,D]   memory * sequence_mask(text_len)
  size_t keys;     // [B,Te,128] values @ memory_layer
  size_t W0r;      // [D+1024,4096] recurrent rows of cell0: (ctx block a + ctx block b) ; h rows
  size_t locF;     // [31,128]  location conv (x) dense composed
  size_t locFb;    // [128]     conv bias @ dense + score bias_b
  // ---- hoisted, input-only work (training: teacher-forced frames are known up front) ----
  size_t frames;   // [T,B,80]   frame fed at step t (zeros at t=0, mel[:,t-1] after)
  size_t pre_h;    // [T,B,256]  prenet layer 0 after relu*mask*2
  size_t pre;      // [T,B,256]  prenet output
  size_t g0pre;    // [T,B,4096] pre @ cell0_kernel[0:256]   (bias added in the cell epilogue)
  // ---- recurrent state / saved activations, time-major; "slot" arrays have T+1 entries, slot 0 = 0 ----
  size_t act0, act1;          // [T,B,4096] gate activations sigma(i) tanh(j) sigma(f+1) sigma(o)
  size_t c0n, c1n;            // [T,B,1024] new cell value before zoneout
  size_t cz0, hz0, cz1, hz1;  // [T+1,B,1024] zoned state after step t at slot t+1
  size_t m0, m1;              // [T,B,1024] cell outputs (un-zoned)
  size_t ctx;                 // [T+1,B,D]
  size_t cum;                 // [T+1,B,Te]
  size_t align_tm;            // [T,B,Te]
  size_t qpart;               // [128,B,128] per-CTA partial query projections of the current step
  size_t qf;                  // [T,B,128]  query projection + composed bias, saved for the reverse pass
  size_t proj_tm;             // [T,B,81]
  size_t barrier;             // 64 B of counters
  // ---- bf16x3 (tcgen05) mode only: operand images in the tensor-core shared-memory layout (decoder_tc.cuh) ----
  size_t wimg_f;              // [128 CTAs][D/256+12 tiles][32 KB] forward weight stream (hi|lo bf16)
  size_t ximg_ctx, ximg_m0, ximg_h0, ximg_h1, ximg_pre, ximg_end;  // [2 parities][K/64 tiles][8 KB] activation images (pre: free-running decode)
  size_t dbg;                 // [T][32] int64 clock64 stamps of CTA 0 at the phase boundaries (profiling aid)
  // ---- reverse pass scratch ----
  size_t bwd_begin;
  size_t W0rT;      // [4096, D+1024]  transpose of W0r
  size_t W1T;       // [4096, 2048]    transpose of cell1_kernel
  size_t dproj_tm;  // [T,B,81]   upstream gradient, time-major
  size_t dm1_proj;  // [T,B,1024] dproj @ Wp[0:1024]^T
  size_t dctx;      // [T,B,D]    in: dproj @ Wp[1024:]^T ; out: total gradient w.r.t. ctx_t
  size_t dG0, dG1;  // [T,B,4096] gradients w.r.t. the gate pre-activations
  size_t dq;        // [T,B,128]
  size_t dkeys;     // [B,Te,128]
  size_t dvalues;   // [B,Te,D]
  size_t dF;        // [31,128] (+ dfb [128] + dsw [128] right behind, one zeroed block)
  size_t dfb, dsw;
  size_t dF_part;   // [32 clusters][31 taps + d score_w][128] per-cluster sums of d F / d score_w, added in cluster order after the loop
  size_t dcum;      // [B,Te]
  size_t dpre;      // [T,B,256]
  size_t dpre_h;    // [T,B,256]
  size_t colsum_scratch;  // [64][4096] partial column sums (bias gradients)
  // ---- bf16x3 reverse kernel only ----
  size_t wimg_b;                         // [128 CTAs][16 tiles][32 KB] reverse weight stream
  size_t ximg_g1, ximg_g0, ximg_g_end;   // [64 k-tiles][8 KB] operand images of dG1_t / dG0_t
  size_t pm0, ph1, ph0, pctx;            // K-quarter partials [4][B][1024] / [4][B][D]
  size_t dbg_b;                          // [T][32] int64 phase stamps of the reverse kernel
  size_t dq2, xexch, dctx_in;            // texts > 128 positions: (dctx_in [T,B,D]: copy of the projection part of d ctx) upper-half d q [T,B,128]; [2][16] inner products + [16][2][4][16] border sums
  size_t total;
};

static inline DecLayout dec_layout(int B, int Te, int L, int D, int T, int mode) {
  (void)L;
  DecLayout l;
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    size_t o = off;
    off += align_up(nfloats * sizeof(float), 256);
    return o;
  };
  const size_t TB = (size_t)T * B, SB = (size_t)(T + 1) * B;
  l.values = take((size_t)B * Te * D);
  l.keys = take((size_t)B * Te * kAtt);
  l.W0r = take((size_t)(D + kCell) * kGates);
  l.locF = take(kConvK * kAtt);
  l.locFb = take(kAtt);
  l.frames = take(TB * kMel);
  l.pre_h = take(TB * kPrenet);
  l.pre = take(TB * kPrenet);
  l.g0pre = take(TB * kGates);
  l.act0 = take(TB * kGates);
  l.act1 = take(TB * kGates);
  l.c0n = take(TB * kCell);
  l.c1n = take(TB * kCell);
  l.cz0 = take(SB * kCell);
  l.hz0 = take(SB * kCell);
  l.cz1 = take(SB * kCell);
  l.hz1 = take(SB * kCell);
  l.m0 = take(TB * kCell);
  l.m1 = take(TB * kCell);
  l.ctx = take(SB * D);
  l.cum = take(SB * Te);
  l.align_tm = take(TB * Te);
  l.qpart = take((size_t)kDecGrid * B * kAtt);
  l.qf = take(TB * kAtt);
  l.proj_tm = take(TB * (kMel + 1));
  l.barrier = take(16);
  l.wimg_f = l.ximg_ctx = l.ximg_m0 = l.ximg_h0 = l.ximg_h1 = l.ximg_pre = l.ximg_end = l.dbg = off;
  if (mode == MSTTS_MODE_BF16X3) {
    auto take_bytes = [&](size_t nbytes) {
      size_t o = off;
      off += align_up(nbytes, 1024);
      return o;
    };
    l.wimg_f = take_bytes((size_t)kDecGrid * (D / 256 + 13) * 32768);  // + the prenet-row tile of the free-running decode
    l.ximg_ctx = take_bytes((size_t)2 * (D / 64) * 8192);
    l.ximg_m0 = take_bytes((size_t)2 * (kCell / 64) * 8192);
    l.ximg_h0 = take_bytes((size_t)2 * (kCell / 64) * 8192);
    l.ximg_h1 = take_bytes((size_t)2 * (kCell / 64) * 8192);
    l.ximg_pre = take_bytes((size_t)2 * (kPrenet / 64) * 8192);
    l.ximg_end = off;
    l.dbg = take_bytes((size_t)(T > kDecGrid ? T : kDecGrid) * 32 * 8);  // also [128 CTAs][32] stamps of the middle step
  }
  l.bwd_begin = off;
  l.W0rT = take((size_t)kGates * (D + kCell));
  l.W1T = take((size_t)kGates * 2 * kCell);
  l.dproj_tm = take(TB * (kMel + 1));
  l.dm1_proj = take(TB * kCell);
  l.dctx = take(TB * D);
  l.dG0 = take(TB * kGates);
  l.dG1 = take(TB * kGates);
  l.dq = take(TB * kAtt);
  l.dkeys = take((size_t)B * Te * kAtt);
  l.dvalues = take((size_t)B * Te * D);
  l.dF = take(kConvK * kAtt + 2 * kAtt);
  l.dfb = l.dF + (size_t)kConvK * kAtt * sizeof(float);
  l.dsw = l.dfb + (size_t)kAtt * sizeof(float);
  l.dF_part = take((size_t)(kDecGrid / kDecCluster) * 32 * kAtt);
  l.dcum = take((size_t)B * Te);
  l.dpre = take(TB * kPrenet);
  l.dpre_h = take(TB * kPrenet);
  l.colsum_scratch = take((size_t)64 * kGates);
  l.wimg_b = l.ximg_g1 = l.ximg_g0 = l.ximg_g_end = l.pm0 = l.ph1 = l.ph0 = l.pctx = l.dbg_b = l.dq2 = l.xexch = l.dctx_in = off;
  if (mode == MSTTS_MODE_BF16X3) {
    auto take_bytes = [&](size_t nbytes) {
      size_t o = off;
      off += align_up(nbytes, 1024);
      return o;
    };
    l.wimg_b = take_bytes((size_t)kDecGrid * 16 * 32768);
    l.ximg_g1 = take_bytes((size_t)(kGates / 64) * 8192);
    l.ximg_g0 = take_bytes((size_t)(kGates / 64) * 8192);
    l.ximg_g_end = off;
    l.pm0 = take((size_t)4 * B * kCell);
    l.ph1 = take((size_t)4 * B * kCell);
    l.ph0 = take((size_t)4 * B * kCell);
    l.pctx = take((size_t)4 * B * D);
    l.dbg_b = take_bytes((size_t)(T > kDecGrid ? T : kDecGrid) * 32 * 8);  // also [128 CTAs][32] stamps of the middle step
    if (Te > 128) {
      l.dq2 = take(TB * kAtt);
      l.xexch = take(32 + 16 * 2 * kDecCluster * 16);
      l.dctx_in = take(TB * D);
    }
  }
  l.total = off;
  return l;
}

// ---- batches beyond one tensor-core tile (bf16x3 mode, B > 32) ----------------------------------------------------
// The persistent tcgen05 loops handle up to 32 batch rows (one M = 64 operand tile of hi + lo rows; 16 rows for texts of
// 129 .. 256 positions, where two clusters serve a row).  Decoder rows are independent, so a larger batch runs as balanced row chunks, each with its own workspace (the saved
// activations of every chunk must survive until the reverse pass) and its own contiguous copy of the dropout / zoneout
// masks (time-major with the batch inside); the weight gradients of the chunks are summed.
struct DecChunkPlan {
  int nchunks, bc;          // chunks of bc rows (the last one may be smaller)
  size_t chunk_ws;          // bytes per chunk workspace (layout of bc rows)
  size_t pm_off, zm_off;    // per-chunk mask copies: [nchunks][T,2,bc,256] / [nchunks][T,2,2,bc,1024]
  size_t pm_bytes, zm_bytes;
  size_t dw_off;            // scratch weight gradients of chunks >= 1 (summed into the caller's)
  size_t dw_floats;
  size_t total;
};
// rows one launch of the tcgen05 loops takes: 32, or 16 for texts beyond 128 positions (two clusters serve a row there)
static inline int dec_chunk_cap(int Te) { return Te > 128 ? 16 : 32; }
static inline bool dec_is_chunked(int B, int Te, int mode) { return mode == MSTTS_MODE_BF16X3 && B > dec_chunk_cap(Te); }
static inline size_t dec_weight_floats(int D) {  // every tensor padded to 64 floats (256-byte aligned slices)
  const size_t n[17] = {(size_t)kMel * kPrenet, kPrenet, (size_t)kPrenet * kPrenet, kPrenet, (size_t)(kPrenet + 2 * D + kCell) * kGates, kGates,
                        (size_t)2 * kCell * kGates, kGates, (size_t)D * kAtt, (size_t)kCell * kAtt, (size_t)kConvK * kConvC, kConvC,
                        (size_t)kConvC * kAtt, kAtt, kAtt, (size_t)(kCell + D) * (kMel + 1), kMel + 1};
  size_t tot = 0;
  for (int i = 0; i < 17; ++i) tot += (n[i] + 63) / 64 * 64;
  return tot;
}
static inline DecChunkPlan dec_chunk_plan(int B, int Te, int L, int D, int T, int mode) {
  DecChunkPlan p;
  const int cap = dec_chunk_cap(Te);
  p.nchunks = (B + cap - 1) / cap;
  p.bc = (B + p.nchunks - 1) / p.nchunks;
  p.chunk_ws = align_up(dec_layout(p.bc, Te, L, D, T, mode).total, 1024);
  size_t off = p.chunk_ws * p.nchunks;
  p.pm_bytes = align_up((size_t)T * 2 * p.bc * kPrenet, 256);
  p.zm_bytes = align_up((size_t)T * 4 * p.bc * kCell, 256);
  p.pm_off = off;
  off += p.pm_bytes * p.nchunks;
  p.zm_off = off;
  off += p.zm_bytes * p.nchunks;
  p.dw_off = off;
  p.dw_floats = dec_weight_floats(D);
  off += p.dw_floats * sizeof(float);
  p.total = off;
  return p;
}
