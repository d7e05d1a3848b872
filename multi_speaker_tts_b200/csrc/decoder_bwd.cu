// Reverse pass of the Tacotron2 decoder loop: ONE persistent reverse-time kernel for the recurrent
// gradients (d h/c of both cells, d context, d cumulative alignment) + hoisted batched GEMMs for every
// weight gradient.  Gradient contract: SURVEY.md A-11 (derived from MSTTS_SV.py:58-98,180-191).
//
// Per reverse step t (same 128-CTA / 32-cluster partition as the forward kernel):
//   phase C'  attention backward, one batch row per cluster: d ctx -> d alignment -> softmax' -> d energy ->
//             d query slice, d keys, d location filter, d cumulative alignment (conv transpose)
//   phase B'e cell-1 gate backward for this CTA's 8 units -> dG1[t]
//   phase B'g dG1[t] @ K1^T for this CTA's 16 input columns (8 of m0, 8 of h1) ; cell-0 gate backward -> dG0[t]
//   phase A'g dG0[t] @ W0r^T for this CTA's ctx / h0 columns -> d ctx_{t-1}, d h0_{t-1}
#include <cooperative_groups.h>

#include "common.cuh"
#include "decoder_gemv.cuh"
#include "decoder_layout.h"
#include "gemm.h"
#include "scratch_pool.h"
#include "tc_gemm.h"

namespace cg = cooperative_groups;

struct DecBwdParams {
  int B, Te, T, D, training, resident;
  const float *W0rT, *W1T, *Wq, *F, *sw;
  const float *keys, *values;
  const int* text_len;
  const uint8_t* zone_mask;
  const float *act0, *act1, *c0n, *c1n, *cz0, *cz1, *qf, *cum, *align_tm;
  const float* dm1_proj;
  float *dctx, *dG0, *dG1, *dq, *dkeys, *dF, *dsw, *dcum;
  float* dF_part;  // [32 clusters][32][128]
  unsigned* barrier;
};

constexpr int kRed16 = 24;  // floats per (slice, batch) row of the 16-column reduction buffer

// gate backward of one (batch, unit): ZoneoutLSTMCell.py:237-260 differentiated
struct CellGrad {
  float di, dj, df, dop, dc_prev, dh_prev;
};
__device__ __forceinline__ CellGrad cell_backward(float dm_direct, float dhz, float dcz, float ig, float jg, float fg,
                                                  float og, float cn, float cp, float mc, float mh) {
  CellGrad r;
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

template <int NB>
__global__ void __cluster_dims__(kDecCluster, 1, 1) __launch_bounds__(kDecThreads, 1)
    decoder_bwd_kernel(const DecBwdParams P) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, l16 = lane & 15;
  const int B = P.B, Te = P.Te, D = P.D, Dq = D / kDecCluster;
  const int unit0 = blockIdx.x * kUnitsPerCta;
  const int crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kDecCluster;
  const int nclusters = gridDim.x / kDecCluster;
  const int TeP = (Te + 15) & ~15;
  const int Dc = (D + kDecGrid - 1) / kDecGrid;  // ctx columns of W0r^T owned by this CTA

  // ---- shared memory carve-up: persistent part, then a scratch region shared by the GEMV reduction buffer
  //      (phases B'g, A'g) and the attention scratch (phase C') ----
  float* wq_s = smem;                                 // [8][129] (padded: read with the unit index varying per lane)
  float* dcz0_s = wq_s + kUnitsPerCta * (kAtt + 1) + 8;// [B][8] carries (gradient w.r.t. zoned states)
  float* dhz0_s = dcz0_s + B * kUnitsPerCta;
  float* dcz1_s = dhz0_s + B * kUnitsPerCta;
  float* dhz1_s = dcz1_s + B * kUnitsPerCta;
  float* qred = dhz1_s + B * kUnitsPerCta;            // [8][32]
  float* bred = qred + 8 * 32;                        // [32]
  float* scratch = bred + 32;
  const int red_floats = max(16 * NB * kRed16, 8 * 32 * 32);  // also holds the final dF/dsw reduction
  const int att_floats = (TeP + 32) + 7 * TeP + 2 * kDecCluster * TeP + (TeP + 32) * 32 + ((Dq + 31) & ~31);
  float* red = scratch;                               // [16][NB][24]
  float* cum_s = scratch;                             // [TeP+32]
  float* e_loc = cum_s + TeP + 32;                    // [TeP]
  float* a_s = e_loc + TeP;                           // [TeP]
  float* da_s = a_s + TeP;                            // [TeP]
  float* de_s = da_s + TeP;                           // [TeP]
  float* dcum_s = de_s + TeP;                         // [TeP]
  float* g_loc = dcum_s + TeP;                        // [TeP]
  float* spare = g_loc + TeP;                         // [TeP]
  float* e_parts1 = spare + TeP;                      // [4][TeP]
  float* e_parts2 = e_parts1 + kDecCluster * TeP;     // [4][TeP]
  float* dps = e_parts2 + kDecCluster * TeP;          // [TeP+32][32]
  float* dctx_s = dps + (TeP + 32) * 32;              // [Dq]
  float* after = scratch + (red_floats > att_floats ? red_floats : att_floats);
  float* keys_s = after;                              // [Te][32]  (resident)
  float* vals_s = keys_s + (P.resident ? Te * 32 : 0);// [Te][Dq]  (resident)
  float* dkeys_s = vals_s + (P.resident ? Te * Dq : 0);// [Te][32] (resident accumulator)

  for (int i = tid; i < kUnitsPerCta * kAtt; i += kDecThreads)
    wq_s[(i / kAtt) * (kAtt + 1) + (i % kAtt)] = P.Wq[(size_t)(unit0 + i / kAtt) * kAtt + (i % kAtt)];
  for (int i = tid; i < 4 * B * kUnitsPerCta; i += kDecThreads) dcz0_s[i] = 0.f;
  float F_reg[kConvK], dF_reg[kConvK];
#pragma unroll
  for (int k = 0; k < kConvK; ++k) {
    F_reg[k] = P.F[k * kAtt + crank * 32 + lane];
    dF_reg[k] = 0.f;
  }
  const float sw_l = P.sw[crank * 32 + lane];
  float dsw_acc = 0.f;
  if (P.resident && cid < B) {
    const float* kg = P.keys + (size_t)cid * Te * kAtt;
    for (int i = tid; i < Te * 32; i += kDecThreads) {
      keys_s[i] = kg[(size_t)(i >> 5) * kAtt + crank * 32 + (i & 31)];
      dkeys_s[i] = 0.f;
    }
    const float* vg = P.values + (size_t)cid * Te * D;
    for (int i = tid; i < Te * Dq; i += kDecThreads) vals_s[i] = vg[(size_t)(i / Dq) * D + crank * Dq + (i % Dq)];
  }
  __syncthreads();

  unsigned bar_target = 0;
  const size_t BC = (size_t)B * kCell, BG = (size_t)B * kGates;

  for (int t = P.T - 1; t >= 0; --t) {
    const uint8_t* zm = P.training ? P.zone_mask + (size_t)t * 4 * BC : nullptr;

    // ================= phase C': attention backward, one batch row per cluster =================
    for (int b = cid; b < B; b += nclusters) {
      const int tl = min(P.text_len[b], Te);
      const float* keys_b = P.resident ? keys_s : P.keys + (size_t)b * Te * kAtt + crank * 32;
      const int kstride = P.resident ? 32 : kAtt;
      const float* vals_b = P.resident ? vals_s : P.values + (size_t)b * Te * D + crank * Dq;
      const int vstride = P.resident ? Dq : D;
      float* dkeys_b = P.resident ? dkeys_s : P.dkeys + (size_t)b * Te * kAtt + crank * 32;
      const float* al = P.align_tm + ((size_t)t * B + b) * Te;
      const float* cum_prev = P.cum + ((size_t)t * B + b) * Te;
      for (int i = tid; i < TeP + 32; i += kDecThreads) cum_s[i] = 0.f;
      for (int i = tid; i < (TeP + 32) * 32; i += kDecThreads) dps[i] = 0.f;
      __syncthreads();
      for (int x = tid; x < Te; x += kDecThreads) {
        a_s[x] = al[x];
        cum_s[15 + x] = cum_prev[x];
        dcum_s[x] = (t == P.T - 1) ? 0.f : __ldcg(P.dcum + (size_t)b * Te + x);
      }
      if (tid < Dq) dctx_s[tid] = P.dctx[((size_t)t * B + b) * D + crank * Dq + tid];
      const float qf = P.qf[((size_t)t * B + b) * kAtt + crank * 32 + lane];
      __syncthreads();
      // partial d a[x] over this CTA's context dims
      for (int x = warp; x < tl; x += 8) {
        float s = 0.f;
        for (int d = lane; d < Dq; d += 32) s = fmaf(dctx_s[d], vals_b[(size_t)x * vstride + d], s);
        s = warp_sum(s);
        if (lane == 0) e_loc[x] = s;
      }
      __syncthreads();
#pragma unroll
      for (int dst = 0; dst < kDecCluster; ++dst) {
        float* remote = cluster.map_shared_rank(e_parts1, dst) + crank * TeP;
        for (int x = tid; x < tl; x += kDecThreads) remote[x] = e_loc[x];
      }
      cluster.sync();
      // softmax backward
      float ldot = 0.f;
      for (int x = tid; x < tl; x += kDecThreads) {
        const float da = (((e_parts1[x] + e_parts1[TeP + x]) + e_parts1[2 * TeP + x]) + e_parts1[3 * TeP + x]) + dcum_s[x];
        da_s[x] = da;
        ldot = fmaf(a_s[x], da, ldot);
      }
      ldot = warp_sum(ldot);
      if (lane == 0) bred[warp] = ldot;
      __syncthreads();
      float dot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) dot += bred[w];
      for (int x = tid; x < tl; x += kDecThreads) de_s[x] = a_s[x] * (da_s[x] - dot);
      __syncthreads();
      // energy backward over this CTA's 32 attention units
      float dq_acc = 0.f;
      for (int blk = warp; blk * 16 < tl; blk += 8) {
        const int t0 = blk * 16;
        float acc[16], dp[16];
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
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int x = t0 + p;
          float dpre = 0.f;
          if (x < tl) {
            const float s = tanhf(keys_b[(size_t)x * kstride + lane] + qf + acc[p]);
            const float de = de_s[x];
            dpre = de * sw_l * (1.f - s * s);
            dsw_acc = fmaf(de, s, dsw_acc);
            dq_acc += dpre;
            dkeys_b[(size_t)x * kstride + lane] += dpre;
            dps[(15 + x) * 32 + lane] = dpre;
          }
          dp[p] = dpre;
        }
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
      qred[warp * 32 + lane] = dq_acc;
      __syncthreads();
      if (tid < 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += qred[w * 32 + tid];
        P.dq[((size_t)t * B + b) * kAtt + crank * 32 + tid] = s;
      }
      // conv transpose: gradient reaching cum_{t-1} through the location features (partial over units)
      for (int blk = warp; blk * 16 < tl; blk += 8) {
        const int t0 = blk * 16;
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
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const float g = warp_sum(G[p]);
          if (lane == 0) g_loc[t0 + p] = g;
        }
      }
      __syncthreads();
#pragma unroll
      for (int dst = 0; dst < kDecCluster; ++dst) {
        float* remote = cluster.map_shared_rank(e_parts2, dst) + crank * TeP;
        for (int x = tid; x < tl; x += kDecThreads) remote[x] = g_loc[x];
      }
      cluster.sync();
      if (crank == 0) {
        for (int x = tid; x < Te; x += kDecThreads) {
          float conv = 0.f;
          if (x < tl) conv = ((e_parts2[x] + e_parts2[TeP + x]) + e_parts2[2 * TeP + x]) + e_parts2[3 * TeP + x];
          P.dcum[(size_t)b * Te + x] = dcum_s[x] + conv;
        }
      }
      if (b + nclusters < B) cluster.sync();
      __syncthreads();
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);

    // ================= phase B'e: cell-1 gate backward for this CTA's units =================
    for (int p = tid; p < B * kUnitsPerCta; p += kDecThreads) {
      const int b = p >> 3, u = p & 7, unit = unit0 + u;
      const float* dqb = P.dq + ((size_t)t * B + b) * kAtt;
      float s = 0.f;
#pragma unroll 8
      for (int a = 0; a < kAtt; ++a) s = fmaf(dqb[a], wq_s[u * (kAtt + 1) + a], s);
      const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
      const float dm_direct = P.dm1_proj[(size_t)t * BC + si] + s;
      const float mc = zm ? (float)zm[2 * BC + si] : 1.f, mh = zm ? (float)zm[3 * BC + si] : 1.f;
      const CellGrad g = cell_backward(dm_direct, dhz1_s[p], dcz1_s[p], P.act1[ai], P.act1[ai + kCell], P.act1[ai + 2 * kCell],
                                       P.act1[ai + 3 * kCell], P.c1n[(size_t)t * BC + si], P.cz1[(size_t)t * BC + si], mc, mh);
      P.dG1[ai] = g.di;
      P.dG1[ai + kCell] = g.dj;
      P.dG1[ai + 2 * kCell] = g.df;
      P.dG1[ai + 3 * kCell] = g.dop;
      dcz1_s[p] = g.dc_prev;
      dhz1_s[p] = g.dh_prev;
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);

    // ================= phase B'g: dG1 @ K1^T (own 8 m0 + 8 h1 columns) ; cell-0 gate backward =================
    for (int b0 = 0; b0 < B; b0 += NB) {
      const int nb = min(NB, B - b0);
      {
        const int col = (l16 >> 3) * kCell + unit0 + (l16 & 7);
        const int slice = warp * 2 + (lane >> 4);
        float acc[NB];
        gemv_acc<NB>(P.W1T, 2 * kCell, col, slice * (kGates / 16), (slice + 1) * (kGates / 16),
                     P.dG1 + (size_t)t * BG + (size_t)b0 * kGates, kGates, nullptr, 0, nb, acc);
#pragma unroll
        for (int b = 0; b < NB; ++b) red[(slice * NB + b) * kRed16 + l16] = acc[b];
      }
      __syncthreads();
      for (int p = tid; p < nb * kUnitsPerCta; p += kDecThreads) {
        const int bl = p >> 3, u = p & 7, unit = unit0 + u, b = b0 + bl, pc = b * kUnitsPerCta + u;
        float dm0 = 0.f, dh1 = 0.f;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) {
          dm0 += red[(sl * NB + bl) * kRed16 + u];
          dh1 += red[(sl * NB + bl) * kRed16 + 8 + u];
        }
        dhz1_s[pc] += dh1;
        const size_t si = (size_t)b * kCell + unit, ai = (size_t)t * BG + (size_t)b * kGates + unit;
        const float mc = zm ? (float)zm[si] : 1.f, mh = zm ? (float)zm[BC + si] : 1.f;
        const CellGrad g = cell_backward(dm0, dhz0_s[pc], dcz0_s[pc], P.act0[ai], P.act0[ai + kCell], P.act0[ai + 2 * kCell],
                                         P.act0[ai + 3 * kCell], P.c0n[(size_t)t * BC + si], P.cz0[(size_t)t * BC + si], mc, mh);
        P.dG0[ai] = g.di;
        P.dG0[ai + kCell] = g.dj;
        P.dG0[ai + 2 * kCell] = g.df;
        P.dG0[ai + 3 * kCell] = g.dop;
        dcz0_s[pc] = g.dc_prev;
        dhz0_s[pc] = g.dh_prev;
      }
      __syncthreads();
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);

    // ================= phase A'g: dG0 @ W0r^T (own ctx columns + own 8 h0 columns) =================
    for (int b0 = 0; b0 < B; b0 += NB) {
      const int nb = min(NB, B - b0);
      {
        int col = -1;
        if (l16 < Dc) {
          const int c = blockIdx.x * Dc + l16;
          if (c < D) col = c;
        } else if (l16 < Dc + kUnitsPerCta) {
          col = D + unit0 + (l16 - Dc);
        }
        const int slice = warp * 2 + (lane >> 4);
        float acc[NB];
        gemv_acc<NB>(P.W0rT, D + kCell, col, slice * (kGates / 16), (slice + 1) * (kGates / 16),
                     P.dG0 + (size_t)t * BG + (size_t)b0 * kGates, kGates, nullptr, 0, nb, acc);
#pragma unroll
        for (int b = 0; b < NB; ++b) red[(slice * NB + b) * kRed16 + l16] = acc[b];
      }
      __syncthreads();
      for (int p = tid; p < nb * 16; p += kDecThreads) {
        const int bl = p >> 4, c16 = p & 15, b = b0 + bl;
        float s = 0.f;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) s += red[(sl * NB + bl) * kRed16 + c16];
        if (c16 < Dc) {
          const int c = blockIdx.x * Dc + c16;
          if (c < D && t > 0) P.dctx[((size_t)(t - 1) * B + b) * D + c] += s;
        } else if (c16 < Dc + kUnitsPerCta) {
          dhz0_s[b * kUnitsPerCta + (c16 - Dc)] += s;
        }
      }
      __syncthreads();
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);
  }

  // ---- flush the per-CTA accumulators ----
  if (P.resident && cid < B) {
    float* dk = P.dkeys + (size_t)cid * Te * kAtt;
    for (int i = tid; i < Te * 32; i += kDecThreads) dk[(size_t)(i >> 5) * kAtt + crank * 32 + (i & 31)] = dkeys_s[i];
  }
  // cross-warp reduction in shared memory, then the cluster's sums go to its own slot of dF_part (added in cluster order by
  // sum_dF_partials_kernel: deterministic, unlike atomics)
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kConvK; ++k) scratch[(warp * 32 + k) * 32 + lane] = dF_reg[k];
  scratch[(warp * 32 + kConvK) * 32 + lane] = dsw_acc;
  __syncthreads();
  for (int i = tid; i < 32 * 32; i += kDecThreads) {
    const int k = i >> 5, u = i & 31;
    float v = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) v += scratch[(w2 * 32 + k) * 32 + u];
    P.dF_part[((size_t)cid * 32 + k) * kAtt + crank * 32 + u] = v;
  }
}

// d F [31,128] and d score_w [128] = the per-cluster sums added in cluster order
__global__ void sum_dF_partials_kernel(const float* __restrict__ part, int nclusters, float* __restrict__ dF, float* __restrict__ dsw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (k, j)
  if (i >= 32 * kAtt) return;
  const int k = i / kAtt, j = i % kAtt;
  float v = 0.f;
  for (int c = 0; c < nclusters; ++c) v += part[((size_t)c * 32 + k) * kAtt + j];
  if (k < kConvK)
    dF[k * kAtt + j] = v;
  else if (k == kConvK)
    dsw[j] = v;
}

// ======================================== host side ================================================
static size_t dec_bwd_smem_bytes(int NB, int B, int Te, int D, int resident) {
  const int TeP = (Te + 15) & ~15, Dq = D / kDecCluster;
  size_t red_floats = (size_t)16 * NB * kRed16;
  if (red_floats < 8 * 32 * 32) red_floats = 8 * 32 * 32;
  const size_t att_floats = (size_t)(TeP + 32) + 7 * TeP + 2 * kDecCluster * TeP + (size_t)(TeP + 32) * 32 + ((Dq + 31) & ~31);
  size_t f = kUnitsPerCta * (kAtt + 1) + 8 + (size_t)4 * B * kUnitsPerCta + 8 * 32 + 32 + (red_floats > att_floats ? red_floats : att_floats);
  if (resident) f += (size_t)Te * 32 * 2 + (size_t)Te * Dq;
  return f * sizeof(float);
}

template <int NB>
static int launch_bwd(const DecBwdParams& P, cudaStream_t stream) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  int max_optin = 0;
  MSTTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DecBwdParams Q = P;
  Q.resident = (P.B <= kDecGrid / kDecCluster) && dec_bwd_smem_bytes(NB, P.B, P.Te, P.D, 1) <= (size_t)max_optin;
  const size_t smem = dec_bwd_smem_bytes(NB, P.B, P.Te, P.D, Q.resident);
  MSTTS_REQUIRE(smem <= (size_t)max_optin, MSTTS_E_UNSUPPORTED, "decoder_bwd: Te=%d needs %zu B smem > %d", P.Te, smem,
                max_optin);
  MSTTS_CUDA(cudaFuncSetAttribute(decoder_bwd_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(kDecGrid);
  cfg.blockDim = dim3(kDecThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // cooperative launch: the runtime gang-schedules the whole grid (all 32 clusters resident before any CTA starts), so the
  // grid barrier cannot deadlock behind a concurrent kernel that holds SMs (an NCCL kernel on another stream, MPS)
  cudaLaunchAttribute coop_attr[1];
  dec_cooperative_attr(&cfg, coop_attr);
  int nclusters = 0;
  MSTTS_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, decoder_bwd_kernel<NB>, &cfg));
  MSTTS_REQUIRE(nclusters * kDecCluster >= kDecGrid, MSTTS_E_DEVICE,
                "decoder_bwd: device co-schedules only %d clusters of %d (need %d)", nclusters, kDecCluster,
                kDecGrid / kDecCluster);
  mstts_timer_start(1, stream);
  MSTTS_CUDA(cudaLaunchKernelEx(&cfg, decoder_bwd_kernel<NB>, Q));
  mstts_timer_stop(1, stream);
  return MSTTS_OK;
}

int dec_bwd_tc_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws, cudaStream_t s,
                     bool wimg_ready);  // decoder_bwd_tc.cu
int dec_bwd_tc_prep_weights(const MsttsDecoderWeights* w, const DecLayout& l, char* ws, int D, cudaStream_t s);

static int dec_bwd_persistent(const DecBwdParams& P, cudaStream_t stream) {
  const int B = P.B;
  if (B <= 1) return launch_bwd<1>(P, stream);
  if (B <= 2) return launch_bwd<2>(P, stream);
  if (B <= 4) return launch_bwd<4>(P, stream);
  if (B <= 8) return launch_bwd<8>(P, stream);
  if (B <= 16) return launch_bwd<16>(P, stream);
  return launch_bwd<32>(P, stream);
}

// ---- small kernels around the loop -----------------------------------------------------------------
// out[c][r] = in[r][c]   (in: R x C row-major)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = in[(size_t)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = bx + j, r = by + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][j];
  }
}

// d_linear [B][T][80], d_stop [B][T] -> dproj_tm [T][B][81]
__global__ void gather_dproj_kernel(const float* __restrict__ dlin, const float* __restrict__ dstop,
                                    float* __restrict__ dproj, int B, int T) {
  const size_t n = (size_t)T * B * (kMel + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (kMel + 1));
    const size_t tb = i / (kMel + 1);
    const size_t t = tb / B, b = tb % B;
    dproj[i] = (c < kMel) ? dlin[(b * T + t) * kMel + c] : dstop[b * T + t];
  }
}

// out[c] = sum_r in[r][c]   (R x C row-major).  Two deterministic passes: kColsumSlices row slices per 32-column
// group (coalesced 128-byte rows, fixed order inside a slice), then a fixed-order sum over the slices.
constexpr int kColsumSlices = 64;
__global__ void colsum_partial_kernel(const float* __restrict__ in, float* __restrict__ part, size_t R, int C) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t r0 = R * blockIdx.y / kColsumSlices, r1 = R * (blockIdx.y + 1) / kColsumSlices;
  float s = 0.f;
  if (c < C)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) s += in[r * C + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += sh[j][threadIdx.x];
    part[(size_t)blockIdx.y * C + c] = tot;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float tot = 0.f;
  for (int j = 0; j < kColsumSlices; ++j) tot += part[(size_t)j * C + c];
  out[c] = tot;
}
static void colsum(cudaStream_t s, const float* in, float* out, size_t R, int C, float* scratch) {
  colsum_partial_kernel<<<dim3((C + 31) / 32, kColsumSlices), dim3(32, 8), 0, s>>>(in, scratch, R, C);
  colsum_final_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, out, C);
}

// prenet backward through relu + dropout: dz = 2 * dy * [y > 0]   (y = relu(z) * 2 * mask)
__global__ void prenet_act_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dy[i] = (y[i] > 0.f) ? 2.f * dy[i] : 0.f;
}

// location filter chain rule: F = Wc @ Wd, fb = bc @ Wd + bias_b
__global__ void location_grads_kernel(const float* __restrict__ dF, const float* __restrict__ dfb,
                                      const float* __restrict__ Wc, const float* __restrict__ bc,
                                      const float* __restrict__ Wd, float* __restrict__ dWc, float* __restrict__ dbc,
                                      float* __restrict__ dWd, float* __restrict__ dbias_b) {
  const int tid = threadIdx.x;  // 1 block, 1024 threads
  // dWc[k][c] = sum_u dF[k][u] Wd[c][u]
  for (int i = tid; i < kConvK * kConvC; i += blockDim.x) {
    const int k = i / kConvC, c = i % kConvC;
    float s = 0.f;
    for (int u = 0; u < kAtt; ++u) s = fmaf(dF[k * kAtt + u], Wd[c * kAtt + u], s);
    dWc[i] = s;
  }
  // dbc[c] = sum_u dfb[u] Wd[c][u]
  for (int c = tid; c < kConvC; c += blockDim.x) {
    float s = 0.f;
    for (int u = 0; u < kAtt; ++u) s = fmaf(dfb[u], Wd[c * kAtt + u], s);
    dbc[c] = s;
  }
  // dWd[c][u] = sum_k Wc[k][c] dF[k][u] + bc[c] dfb[u]
  for (int i = tid; i < kConvC * kAtt; i += blockDim.x) {
    const int c = i / kAtt, u = i % kAtt;
    float s = bc[c] * dfb[u];
    for (int k = 0; k < kConvK; ++k) s = fmaf(Wc[k * kConvC + c], dF[k * kAtt + u], s);
    dWd[i] = s;
  }
  for (int u = tid; u < kAtt; u += blockDim.x) dbias_b[u] = dfb[u];
}

// d_memory = (dvalues + dkeys @ Wm^T) * sequence_mask  -- dvalues already holds both terms; apply the mask
__global__ void mask_dmemory_kernel(const float* __restrict__ dvalues, const int* __restrict__ text_len,
                                    float* __restrict__ dmem, int B, int Te, int D) {
  const size_t n = (size_t)B * Te * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / D;
    const int b = (int)(row / Te), x = (int)(row % Te);
    dmem[i] = (x < text_len[b]) ? dvalues[i] : 0.f;
  }
}

// copy rows [r0, r0+n) of a [*,4096] gradient block to rows [r1, r1+n)  (the context enters cell 0 twice)
__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static inline int ew_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// ---- products beside the reverse loop (see the weight-gradient section of decoder_bwd_one) ----
constexpr int kDecIdleSMs = 20;  // 148 SMs - the 128 the persistent loop holds
struct DecSideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
static int dec_side_stream(DecSideStream** out) {
  static thread_local DecSideStream table[64];
  static thread_local bool made[64] = {false};
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  dev &= 63;
  if (!made[dev]) {
    MSTTS_CUDA(cudaStreamCreateWithFlags(&table[dev].stream, cudaStreamNonBlocking));
    MSTTS_CUDA(cudaEventCreateWithFlags(&table[dev].fork, cudaEventDisableTiming));
    MSTTS_CUDA(cudaEventCreateWithFlags(&table[dev].join, cudaEventDisableTiming));
    made[dev] = true;
  }
  *out = &table[dev];
  return MSTTS_OK;
}
// MSTTS_NO_OVERLAP=1 in the environment runs every weight-gradient product after the loop (A/B measurements)
static bool dec_overlap_enabled() {  // read per call: bench.py times both settings in one process
  const char* e = getenv("MSTTS_NO_OVERLAP");
  return !(e && e[0] == '1');
}
static int dec_env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}
// one warp; lane 0 polls the loop's barrier counter (monotonic) until it reaches `target`
__global__ void wait_counter_kernel(const unsigned* __restrict__ counter, unsigned target, unsigned sleep_ns) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(counter) < target) __nanosleep(sleep_ns);
    __threadfence();
  }
  __syncwarp();
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

static int decoder_bwd_one(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const MsttsDecoderGrads* g,
                           const MsttsDecoderWeightGrads* dw, void* ws_, size_t ws_bytes, cudaStream_t s);

// B > 32 in bf16x3 mode: the row chunks of the forward call (DecChunkPlan), weight gradients summed over the chunks
static int decoder_bwd_chunked(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const MsttsDecoderGrads* g,
                               const MsttsDecoderWeightGrads* dw, void* ws_, size_t ws_bytes, cudaStream_t s) {
  const int B = io->B, Te = io->Te, L = io->L, D = io->D, T = io->n_steps;
  const DecChunkPlan p = dec_chunk_plan(B, Te, L, D, T, io->mode);
  MSTTS_REQUIRE(ws_bytes >= p.total, MSTTS_E_WORKSPACE, "decoder_bwd: workspace %zu < %zu", ws_bytes, p.total);
  char* ws = (char*)ws_;
  const size_t n[17] = {(size_t)kMel * kPrenet, kPrenet, (size_t)kPrenet * kPrenet, kPrenet, (size_t)(kPrenet + 2 * D + kCell) * kGates, kGates,
                        (size_t)2 * kCell * kGates, kGates, (size_t)D * kAtt, (size_t)kCell * kAtt, (size_t)kConvK * kConvC, kConvC,
                        (size_t)kConvC * kAtt, kAtt, kAtt, (size_t)(kCell + D) * (kMel + 1), kMel + 1};
  static_assert(sizeof(MsttsDecoderWeightGrads) == 17 * sizeof(float*), "weight-gradient struct has 17 tensors");
  MsttsDecoderWeightGrads tmp;
  {
    float** tp = reinterpret_cast<float**>(&tmp);
    float* base = (float*)(ws + p.dw_off);
    size_t off = 0;
    for (int i = 0; i < 17; ++i) {
      tp[i] = base + off;
      off += (n[i] + 63) / 64 * 64;
    }
  }
  for (int c = 0; c < p.nchunks; ++c) {
    const int b0 = c * p.bc, bc = (B - b0 < p.bc) ? B - b0 : p.bc;
    MsttsDecoderIO sub = *io;
    sub.B = bc;
    sub.memory = io->memory + (size_t)b0 * Te * D;
    sub.text_len = io->text_len + b0;
    sub.mel = io->mel + (size_t)b0 * L * kMel;
    sub.mel_len = io->mel_len + b0;
    sub.prenet_mask = (const uint8_t*)(ws + p.pm_off + (size_t)c * p.pm_bytes);  // gathered by the forward call
    sub.zone_mask = (const uint8_t*)(ws + p.zm_off + (size_t)c * p.zm_bytes);
    sub.linear = io->linear + (size_t)b0 * T * kMel;
    sub.stop = io->stop + (size_t)b0 * T;
    sub.align = io->align + (size_t)b0 * T * Te;
    MsttsDecoderGrads gs;
    gs.d_linear = g->d_linear + (size_t)b0 * T * kMel;
    gs.d_stop = g->d_stop + (size_t)b0 * T;
    gs.d_memory = g->d_memory ? g->d_memory + (size_t)b0 * Te * D : nullptr;
    int rc = decoder_bwd_one(w, &sub, &gs, c == 0 ? dw : &tmp, ws + (size_t)c * p.chunk_ws, p.chunk_ws, s);
    if (rc) return rc;
    if (c > 0) {
      float* const* dst = reinterpret_cast<float* const*>(dw);
      float* const* src = reinterpret_cast<float* const*>(&tmp);
      for (int i = 0; i < 17; ++i) {
        size_t gsz = (n[i] + 255) / 256;
        add_inplace_kernel<<<(int)(gsz < 148 * 8 ? gsz : 148 * 8), 256, 0, s>>>(dst[i], src[i], n[i]);
      }
    }
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

extern "C" int mstts_decoder_bwd(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const MsttsDecoderGrads* g,
                                 const MsttsDecoderWeightGrads* dw, void* ws_, size_t ws_bytes, void* stream_) {
  MSTTS_REQUIRE(w && io && g && dw && ws_, MSTTS_E_INVALID, "decoder_bwd: null argument");
  MSTTS_REQUIRE(io->is_training && (io->mode == MSTTS_MODE_FP32 || io->mode == MSTTS_MODE_BF16X3), MSTTS_E_UNSUPPORTED,
                "decoder_bwd: only training in fp32 / bf16x3 mode is implemented");
  MSTTS_REQUIRE(g->d_linear && g->d_stop, MSTTS_E_INVALID, "decoder_bwd: null upstream gradient");
  {
    const float* const* gp = reinterpret_cast<const float* const*>(dw);
    for (size_t i = 0; i < sizeof(MsttsDecoderWeightGrads) / sizeof(float*); ++i)
      MSTTS_REQUIRE(gp[i], MSTTS_E_INVALID, "decoder_bwd: null weight-gradient pointer #%zu", i);
  }
  if (dec_is_chunked(io->B, io->Te, io->mode)) return decoder_bwd_chunked(w, io, g, dw, ws_, ws_bytes, (cudaStream_t)stream_);
  return decoder_bwd_one(w, io, g, dw, ws_, ws_bytes, (cudaStream_t)stream_);
}

static int decoder_bwd_one(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const MsttsDecoderGrads* g,
                           const MsttsDecoderWeightGrads* dw, void* ws_, size_t ws_bytes, cudaStream_t s) {
  const int B = io->B, Te = io->Te, L = io->L, D = io->D, T = io->n_steps;
  const DecLayout l = dec_layout(B, Te, L, D, T, io->mode);
  MSTTS_REQUIRE(ws_bytes >= l.total, MSTTS_E_WORKSPACE, "decoder_bwd: workspace %zu < %zu", ws_bytes, l.total);
  char* ws = (char*)ws_;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  const size_t TB = (size_t)T * B;
  const int K0r = D + kCell, NP = kMel + 1;
  int rc;

  const bool tc = io->mode == MSTTS_MODE_BF16X3;
  // the reverse loop's weight image depends on the weights only: built on the side stream while this one computes the
  // projection part of the gradients
  bool wimg_ready = false;
  if (tc && T >= 64 && dec_overlap_enabled()) {
    DecSideStream* sd = nullptr;
    if ((rc = dec_side_stream(&sd))) return rc;
    MSTTS_CUDA(cudaEventRecord(sd->fork, s));  // after the previous use of the workspace on this stream
    MSTTS_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
    if ((rc = dec_bwd_tc_prep_weights(w, l, ws, D, sd->stream))) return rc;
    MSTTS_CUDA(cudaEventRecord(sd->join, sd->stream));
    wimg_ready = true;
  }
  // ---- upstream gradient through the hoisted projection ----
  gather_dproj_kernel<<<ew_grid(TB * NP), 256, 0, s>>>(g->d_linear, g->d_stop, F(l.dproj_tm), B, T);
  // d m1 (projection part) and d ctx (projection part): what the loop needs.  The projection's own weight gradient does not
  // feed the loop and waits until after it (below, beside the second chain).
  if ((rc = gemm_rowmajor_ex(s, false, true, (int)TB, kCell, NP, F(l.dproj_tm), NP, w->proj_kernel, NP, F(l.dm1_proj), kCell, 0.f))) return rc;
  if ((rc = gemm_rowmajor_ex(s, false, true, (int)TB, D, NP, F(l.dproj_tm), NP, w->proj_kernel + (size_t)kCell * NP, NP,
                             F(l.dctx), D, 0.f))) return rc;
  // ---- transposed recurrent weights (the tcgen05 kernel reads the reference layout directly) ----
  if (!tc) {
    transpose_kernel<<<dim3(kGates / 32, (K0r + 31) / 32), dim3(32, 8), 0, s>>>(F(l.W0r), F(l.W0rT), K0r, kGates);
    transpose_kernel<<<dim3(kGates / 32, 2 * kCell / 32), dim3(32, 8), 0, s>>>(w->cell1_kernel, F(l.W1T), 2 * kCell, kGates);
  }
  MSTTS_CUDA(cudaMemsetAsync(ws + l.dF, 0, (kConvK * kAtt + 2 * kAtt) * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.dkeys, 0, (size_t)B * Te * kAtt * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.barrier, 0, 64, s));

  // ---- the reverse loop ----
  DecBwdParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.Te = Te; P.T = T; P.D = D; P.training = io->is_training;
  P.W0rT = F(l.W0rT); P.W1T = F(l.W1T); P.Wq = w->query_kernel; P.F = F(l.locF); P.sw = w->score_w;
  P.keys = F(l.keys); P.values = F(l.values); P.text_len = io->text_len; P.zone_mask = io->zone_mask;
  P.act0 = F(l.act0); P.act1 = F(l.act1); P.c0n = F(l.c0n); P.c1n = F(l.c1n); P.cz0 = F(l.cz0); P.cz1 = F(l.cz1);
  P.qf = F(l.qf); P.cum = F(l.cum); P.align_tm = F(l.align_tm); P.dm1_proj = F(l.dm1_proj);
  P.dctx = F(l.dctx); P.dG0 = F(l.dG0); P.dG1 = F(l.dG1); P.dq = F(l.dq); P.dkeys = F(l.dkeys);
  P.dF = F(l.dF); P.dsw = F(l.dsw); P.dcum = F(l.dcum); P.dF_part = F(l.dF_part);
  P.barrier = (unsigned*)(ws + l.barrier);
  // ---- weight gradients of the two cells: dW[rows, 4096] = X[T*B, rows]^T dG[T*B, 4096] ----
  // ~1 TFLOP at config 2, bf16x3 on the hand-written tcgen05 kernel (every product here is TC_FAST: the higher precision levels
  // were measured on the full-size gradients, tools/grad_probe.py, and change nothing -- the agreement with the oracle is
  // limited by the ReLU / L1 decisions of the forward pass, tests/test_full_size_gpu.py).  The contraction runs over time, and
  // the reverse loop finishes its steps from the last one down, so the products are cut into time chunks: chunk c (the rows of
  // steps [t_lo, t_hi)) is packed and accumulated ON A SIDE STREAM, ON THE 20 SMs THE PERSISTENT LOOP LEAVES IDLE, as soon as
  // the loop's barrier counter shows that step t_lo is done (a one-thread kernel polls it); only the last chunk runs after
  // the loop, on the whole device.  Every side launch is capped at 20 SMs' worth of CTAs (TcGridCap): the loop's cooperative
  // launch can always become resident, whatever the order in which the two streams start.
  const int KbT = (int)((TB + 63) / 64);
  const int xmax = D > kCell ? D : kCell;
  float* dK0_ctx = dw->cell0_kernel + (size_t)kPrenet * kGates;
  // rows [r0, r0 + nr) of the time-major operands -> both cells' products, accumulated when acc
  auto wgrad_rows = [&](cudaStream_t st, void* gimg, void* ximg, size_t r0, int nr, bool acc) -> int {
    const int kb = (nr + 63) / 64;
    const float beta = acc ? 1.f : 0.f;
    int r2;
    auto one = [&](const float* X, int rows, float* dW) -> int {
      int r3 = tc_pack_f32(st, X + r0 * rows, rows, true, rows, nr, 128, kb, ximg, 0, 0);
      if (r3) return r3;
      return tc_gemm_images(st, ximg, gimg, rows, kGates, nr, dW, kGates, beta);
    };
    // cell 1: rows [m0 | h1_prev]
    if ((r2 = tc_pack_f32(st, F(l.dG1) + r0 * kGates, kGates, true, kGates, nr, 256, kb, gimg, 0, 0))) return r2;
    if ((r2 = one(F(l.m0), kCell, dw->cell1_kernel))) return r2;
    if ((r2 = one(F(l.hz1), kCell, dw->cell1_kernel + (size_t)kCell * kGates))) return r2;
    // cell 0: rows [prenet | ctx | (ctx again: copied at the end) | h0_prev]
    if ((r2 = tc_pack_f32(st, F(l.dG0) + r0 * kGates, kGates, true, kGates, nr, 256, kb, gimg, 0, 0))) return r2;
    if ((r2 = one(F(l.pre), kPrenet, dw->cell0_kernel))) return r2;
    if ((r2 = one(F(l.ctx), D, dK0_ctx))) return r2;
    return one(F(l.hz0), kCell, dw->cell0_kernel + (size_t)(kPrenet + 2 * D) * kGates);
  };
  // time chunks: NCH - 1 of them beside the loop (highest steps first), the last one (lowest steps, incl. step 0) after it
  const int bars_per_step = (tc && Te > 128) ? 6 : 5;   // grid barriers of one reverse step (decoder_bwd_tc.cu)
  const int NCH = (tc && T >= 64 && dec_overlap_enabled()) ? dec_env_int("MSTTS_OVERLAP_CHUNKS", 8, 2, 64) : 1;
  const int t_split = NCH > 1 ? T / NCH : T;             // the after-loop chunk covers steps [0, t_split)
  DecSideStream* side = nullptr;
  if (NCH > 1) {
    if ((rc = dec_side_stream(&side))) return rc;
    MSTTS_CUDA(cudaEventRecord(side->fork, s));           // the barrier counter is zeroed, the saved activations are complete
    MSTTS_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  if (wimg_ready) MSTTS_CUDA(cudaStreamWaitEvent(s, side->join, 0));  // recorded after the weight-image kernel (above)
  if ((rc = tc ? dec_bwd_tc_entry(w, io, l, ws, s, wimg_ready) : dec_bwd_persistent(P, s))) return rc;
  sum_dF_partials_kernel<<<(32 * kAtt + 255) / 256, 256, 0, s>>>(F(l.dF_part), kDecGrid / kDecCluster, F(l.dF), F(l.dsw));
  if (NCH > 1) {
    TcGridCap cap(dec_env_int("MSTTS_OVERLAP_SMS", kDecIdleSMs, 1, kDecIdleSMs));
    ScratchScope ssc(side->stream);
    void *gimg_s = nullptr, *ximg_s = nullptr;
    const int max_steps = (T - t_split + NCH - 2) / (NCH - 1);
    if ((rc = ssc.get(&gimg_s, tc_image_bytes(kGates, max_steps * B, 256)))) return rc;
    if ((rc = ssc.get(&ximg_s, tc_image_bytes(xmax, max_steps * B, 128)))) return rc;
    int t_hi = T;
    for (int c = 0; c < NCH - 1; ++c) {
      const int t_lo = t_split + (int)((long long)(T - t_split) * (NCH - 2 - c) / (NCH - 1));
      // every CTA adds 1 per grid barrier; step t is complete (its dG rows written and released) once the last barrier of that
      // step has been passed: bars_per_step * (T - t) barriers since the start of the loop
      const unsigned target = (unsigned)kDecGrid * (unsigned)bars_per_step * (unsigned)(T - t_lo);
      wait_counter_kernel<<<1, 32, 0, side->stream>>>((const unsigned*)(ws + l.barrier), target,
                                                      (unsigned)dec_env_int("MSTTS_OVERLAP_POLL_NS", 200, 20, 100000));
      if ((rc = wgrad_rows(side->stream, gimg_s, ximg_s, (size_t)t_lo * B, (t_hi - t_lo) * B, c > 0))) return rc;
      t_hi = t_lo;
    }
    MSTTS_CUDA(cudaEventRecord(side->join, side->stream));
    MSTTS_CUDA(cudaStreamWaitEvent(s, side->join, 0));
  }
  // After the loop two chains that share nothing but read-only inputs run side by side: the last time chunk of the cell weight
  // gradients + the cell bias gradients on the caller's stream, and everything else (query layer, location filter, the prenet
  // chain, the memory side) on the side stream -- mostly small grids (one-block location kernel, narrow products) that leave
  // the device half empty when queued one after the other.
  cudaStream_t st = s;  // the stream of the second chain
  if (NCH > 1 && dec_env_int("MSTTS_TAIL_STREAMS", 1, 0, 1)) {
    st = side->stream;
    MSTTS_CUDA(cudaEventRecord(side->fork, s));  // the loop has finished: dG0 / dq / dF / dctx / dkeys are final
    MSTTS_CUDA(cudaStreamWaitEvent(st, side->fork, 0));
  }
  {
    ScratchScope sc(s);
    void *gimg = nullptr, *ximg = nullptr;
    const int nr = t_split * B;
    if ((rc = sc.get(&gimg, tc_image_bytes(kGates, nr, 256)))) return rc;
    if ((rc = sc.get(&ximg, tc_image_bytes(xmax, nr, 128)))) return rc;
    if ((rc = wgrad_rows(s, gimg, ximg, 0, nr, NCH > 1))) return rc;
  }
  (void)KbT;
  copy_rows_kernel<<<ew_grid((size_t)D * kGates), 256, 0, s>>>(dK0_ctx, dK0_ctx + (size_t)D * kGates, (size_t)D * kGates);
  colsum(s, F(l.dG1), dw->cell1_bias, TB, kGates, F(l.colsum_scratch));
  colsum(s, F(l.dG0), dw->cell0_bias, TB, kGates, F(l.colsum_scratch));
  // projection: dWp = [m1 | ctx]^T dproj ; dbp = colsum(dproj)
  if ((rc = gemm_rowmajor_ex(s, true, false, kCell, NP, (int)TB, F(l.m1), kCell, F(l.dproj_tm), NP, dw->proj_kernel, NP, 0.f))) return rc;
  if ((rc = gemm_rowmajor_ex(s, true, false, D, NP, (int)TB, F(l.ctx) + (size_t)B * D, D, F(l.dproj_tm), NP,
                             dw->proj_kernel + (size_t)kCell * NP, NP, 0.f))) return rc;
  colsum(s, F(l.dproj_tm), dw->proj_bias, TB, NP, F(l.colsum_scratch));
  {
    ScratchScope sc2(st);
    float* cs = F(l.colsum_scratch);  // the second chain's column sums are at most 256 wide: own partials when it runs concurrently
    if (st != s) {
      void* p = nullptr;
      if ((rc = sc2.get(&p, (size_t)kColsumSlices * 256 * sizeof(float)))) return rc;
      cs = (float*)p;
    }
    // query layer: dWq = m1^T dq ; composed-bias gradient dfb = colsum(dq)
    if ((rc = gemm_rowmajor_ex(st, true, false, kCell, kAtt, (int)TB, F(l.m1), kCell, F(l.dq), kAtt, dw->query_kernel, kAtt, 0.f))) return rc;
    colsum(st, F(l.dq), F(l.dfb), TB, kAtt, cs);
    location_grads_kernel<<<1, 1024, 0, st>>>(F(l.dF), F(l.dfb), w->loc_conv_kernel, w->loc_conv_bias, w->loc_dense_kernel,
                                              dw->loc_conv_kernel, dw->loc_conv_bias, dw->loc_dense_kernel, dw->score_b);
    MSTTS_CUDA(cudaMemcpyAsync(dw->score_w, F(l.dsw), kAtt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // prenet: d pre = dG0 @ K0[0:256]^T, then back through the two dense+relu+dropout layers
    if ((rc = gemm_rowmajor_ex(st, false, true, (int)TB, kPrenet, kGates, F(l.dG0), kGates, w->cell0_kernel, kGates, F(l.dpre), kPrenet, 0.f))) return rc;
    prenet_act_bwd_kernel<<<ew_grid(TB * kPrenet), 256, 0, st>>>(F(l.dpre), F(l.pre), TB * kPrenet);
    if ((rc = gemm_rowmajor_ex(st, true, false, kPrenet, kPrenet, (int)TB, F(l.pre_h), kPrenet, F(l.dpre), kPrenet, dw->prenet1_kernel, kPrenet, 0.f))) return rc;
    colsum(st, F(l.dpre), dw->prenet1_bias, TB, kPrenet, cs);
    if ((rc = gemm_rowmajor_ex(st, false, true, (int)TB, kPrenet, kPrenet, F(l.dpre), kPrenet, w->prenet1_kernel, kPrenet, F(l.dpre_h), kPrenet, 0.f))) return rc;
    prenet_act_bwd_kernel<<<ew_grid(TB * kPrenet), 256, 0, st>>>(F(l.dpre_h), F(l.pre_h), TB * kPrenet);
    if ((rc = gemm_rowmajor_ex(st, true, false, kMel, kPrenet, (int)TB, F(l.frames), kMel, F(l.dpre_h), kPrenet, dw->prenet0_kernel, kPrenet, 0.f))) return rc;
    colsum(st, F(l.dpre_h), dw->prenet0_bias, TB, kPrenet, cs);
    // memory side: dvalues[b] = A_b^T dctx_b (over steps) + dkeys[b] @ Wm^T ; dWm = values^T dkeys
    if ((rc = gemm_rowmajor_batched(st, true, false, Te, D, T, F(l.align_tm), B * Te, Te, F(l.dctx), B * D, D, F(l.dvalues), D,
                                    (long long)Te * D, 0.f, B))) return rc;
    if ((rc = gemm_rowmajor_ex(st, false, true, B * Te, D, kAtt, F(l.dkeys), kAtt, w->memory_kernel, kAtt, F(l.dvalues), D, 1.f))) return rc;
    if ((rc = gemm_rowmajor_ex(st, true, false, D, kAtt, B * Te, F(l.values), D, F(l.dkeys), kAtt, dw->memory_kernel, kAtt, 0.f))) return rc;
    if (g->d_memory)
      mask_dmemory_kernel<<<ew_grid((size_t)B * Te * D), 256, 0, st>>>(F(l.dvalues), io->text_len, g->d_memory, B, Te, D);
  }
  if (st != s) {
    MSTTS_CUDA(cudaEventRecord(side->join, st));
    MSTTS_CUDA(cudaStreamWaitEvent(s, side->join, 0));
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
