// Tacotron2 decoder loop, forward, as ONE persistent kernel: teacher-forced (training) and free-running
// (inference, Modules.py:212-237: the projected frame of step t is the prenet input of step t+1, the loop ends
// when every row has emitted stop >= 0 or the step cap is reached).
//
// Replaces the tf.while_loop body of Modules.py:397-443 (Decoder_Dynamic_Decode) and everything it calls
// per step: ZoneoutLSTMCell.call x2 (ZoneoutLSTMCell.py:228-264), Location_Sensitive_Attention.__call__
// (Location_Sensitive_Attention.py:43-85) and the AttentionWrapper context.  Input-only work (prenet,
// the prenet rows of cell0's kernel, memory_layer keys, the mel/stop projection in teacher-forced mode)
// is hoisted out of the loop into batched GEMMs.
//
// Grid: 128 CTAs = 32 clusters x 4 (B200: 148 SMs, cluster-4 can place 33 clusters).
//   phase A  cell 0 : CTA j owns LSTM units 8j..8j+7 (32 gate columns), K split over its 8 warps
//   phase B  cell 1 : same split; epilogue also emits this CTA's partial query projection
//   phase C  attention: cluster c owns batch row c; CTA r of the cluster owns attention units 32r..32r+31
//            (keys slice resident in smem) and context dims r*D/4.. (values slice resident in smem);
//            partial energies are exchanged through distributed shared memory.
//   one device-wide barrier after each phase.
#include <cooperative_groups.h>

#include "common.cuh"
#include "decoder_layout.h"

namespace cg = cooperative_groups;

struct DecFwdParams {
  int B, Te, T, D, training, resident;
  const float *W0r, *W1, *b0, *b1, *Wq, *F, *fb, *sw;
  const float *g0pre, *keys, *values;
  // free-running mode only: un-hoisted prenet + projection (Modules.py:222-255,309-321)
  const float *K0pre, *Wp, *bp, *P0, *pb0, *P1, *pb1;
  const uint8_t* prenet_mask;  // [T,2,B,256]
  float *pre, *proj_tm;        // [T,B,256], [T,B,81] (bias-free; finish_outputs adds it)
  int* steps_done;
  const int* text_len;
  const uint8_t* zone_mask;
  float *act0, *act1, *c0n, *c1n, *cz0, *hz0, *cz1, *hz1, *m0, *m1, *ctx, *cum, *align_tm, *qpart, *qf;
  unsigned* barrier;
};


#include "decoder_gemv.cuh"

// ---- gate math + zoneout for this CTA's 8 units (ZoneoutLSTMCell.py:230-264) ----------------------
template <int NB>
__device__ __forceinline__ void lstm_epilogue(const float* red, int nb, int unit0, const float* addend,
                                              const float* __restrict__ bias, const float* c_prev,
                                              const float* h_prev, const uint8_t* __restrict__ mask_c,
                                              const uint8_t* __restrict__ mask_h, float* act, float* cn,
                                              float* cz, float* hz, float* m, float* m_s) {
  for (int p = threadIdx.x; p < nb * kUnitsPerCta; p += blockDim.x) {
    const int b = p >> 3, u = p & 7, unit = unit0 + u;
    float g[4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[(w * NB + b) * kRedStride + gi * 8 + u];
      if (addend) s += addend[(size_t)b * kGates + gi * kCell + unit];
      s += bias[gi * kCell + unit];
      g[gi] = s;
    }
    const float ig = sigmoidf_precise(g[0]);
    const float jg = tanhf(g[1]);
    const float fg = sigmoidf_precise(g[2] + kForgetBias);
    const float og = sigmoidf_precise(g[3]);
    const size_t si = (size_t)b * kCell + unit;
    const float cp = c_prev[si], hp = h_prev[si];
    const float c = fg * cp + ig * jg;
    const float mm = og * tanhf(c);
    float dc = c - cp, dm = mm - hp;
    if (mask_c) {
      dc *= (float)mask_c[si];
      dm *= (float)mask_h[si];
    }
    const size_t ai = (size_t)b * kGates + unit;
    act[ai] = ig;
    act[ai + kCell] = jg;
    act[ai + 2 * kCell] = fg;
    act[ai + 3 * kCell] = og;
    cn[si] = c;
    cz[si] = kZoneKeep * dc + cp;
    hz[si] = kZoneKeep * dm + hp;
    m[si] = mm;
    if (m_s) m_s[b * kUnitsPerCta + u] = mm;
  }
}

template <int NB>
__global__ void __cluster_dims__(kDecCluster, 1, 1) __launch_bounds__(kDecThreads, 1)
    decoder_fwd_kernel(const DecFwdParams P) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = P.B, Te = P.Te, D = P.D, Dq = D / kDecCluster;
  const int unit0 = blockIdx.x * kUnitsPerCta;
  const int crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kDecCluster;
  const int nclusters = gridDim.x / kDecCluster;
  const int TeP = (Te + 15) & ~15;

  // ---- shared memory carve-up ----
  float* red = smem;                                  // [8][NB][40]
  float* wq_s = red + 8 * NB * kRedStride;            // [8][128]  query_layer rows of this CTA's units
  float* m_s = wq_s + kUnitsPerCta * kAtt;            // [NB][8]
  float* qred = m_s + NB * kUnitsPerCta;              // [8][32]
  float* qf_s = qred + 8 * 32;                        // [32]   q + composed bias, this CTA's unit slice
  float* cum_s = qf_s + 32;                           // [TeP+32] cum shifted by 15, zero padded
  float* e_loc = cum_s + TeP + 32;                    // [TeP]
  float* e_parts = e_loc + TeP;                       // [4][TeP]  written by the 4 CTAs of the cluster
  float* a_s = e_parts + kDecCluster * TeP;           // [TeP]
  float* bred = a_s + TeP;                            // [16]
  float* x_s = bred + 16;                             // [512]  free-running: [m1 slice (256) | ctx slice (Dq)]
  float* pred = x_s + 512;                            // [8][96] projection partials per warp
  float* pparts = pred + 8 * 96;                      // [4][96] projection partials per CTA of the cluster
  float* frame_s = pparts + kDecCluster * 96;         // [96]   projected frame (+ stop logit at [80])
  float* h1_s = frame_s + 96;                         // [256]  prenet layer 0 output
  float* p2red = h1_s + kPrenet;                      // [4][64]
  int* fin_s = reinterpret_cast<int*>(p2red + 256);   // [256]  free-running: finished flag per batch row
  float* keys_s = p2red + 256 + 256;                  // [Te][32]   (resident)
  float* vals_s = keys_s + (P.resident ? Te * 32 : 0);// [Te][Dq]   (resident)

  // ---- one-time staging ----
  for (int i = tid; i < kUnitsPerCta * kAtt; i += kDecThreads)
    wq_s[i] = P.Wq[(size_t)(unit0 + i / kAtt) * kAtt + (i % kAtt)];
  float F_reg[kConvK];
#pragma unroll
  for (int k = 0; k < kConvK; ++k) F_reg[k] = P.F[k * kAtt + crank * 32 + lane];
  const float sw_l = P.sw[crank * 32 + lane];
  const float fb_l = P.fb[crank * 32 + lane];
  if (P.resident && cid < B) {
    const float* kg = P.keys + (size_t)cid * Te * kAtt;
    for (int i = tid; i < Te * 32; i += kDecThreads) keys_s[i] = kg[(size_t)(i >> 5) * kAtt + crank * 32 + (i & 31)];
    const float* vg = P.values + (size_t)cid * Te * D;
    for (int i = tid; i < Te * Dq; i += kDecThreads) vals_s[i] = vg[(size_t)(i / Dq) * D + crank * Dq + (i % Dq)];
  }
  for (int i = tid; i < TeP + 32; i += kDecThreads) cum_s[i] = 0.f;
  fin_s[tid] = 0;
  __syncthreads();

  unsigned bar_target = 0;
  const size_t BC = (size_t)B * kCell, BG = (size_t)B * kGates;

  // free-running mode: prenet of frame_s (this cluster's batch row b) -> pre[slot][b][:]; layer 0 is computed by every
  // CTA of the cluster (80x256), layer 1 is split 64 outputs per CTA.  Dropout stays on (Modules.py:252).
  auto prenet_row = [&](int b, int slot) {
    {
      float s = P.pb0[tid];
      for (int k = 0; k < kMel; ++k) s = fmaf(frame_s[k], __ldg(P.P0 + k * kPrenet + tid), s);
      const float mk = (float)P.prenet_mask[(((size_t)slot * 2 + 0) * B + b) * kPrenet + tid];
      h1_s[tid] = (fmaxf(s, 0.f) / 0.5f) * mk;
    }
    __syncthreads();
    {
      const int j = crank * 64 + (tid & 63), kq = tid >> 6;
      float s = 0.f;
      for (int k = kq * 64; k < kq * 64 + 64; ++k) s = fmaf(h1_s[k], __ldg(P.P1 + k * kPrenet + j), s);
      p2red[kq * 64 + (tid & 63)] = s;
    }
    __syncthreads();
    if (tid < 64) {
      const int j = crank * 64 + tid;
      const float s = ((p2red[tid] + p2red[64 + tid]) + p2red[128 + tid]) + p2red[192 + tid] + P.pb1[j];
      const float mk = (float)P.prenet_mask[(((size_t)slot * 2 + 1) * B + b) * kPrenet + j];
      P.pre[((size_t)slot * B + b) * kPrenet + j] = (fmaxf(s, 0.f) / 0.5f) * mk;
    }
    __syncthreads();
  };
  if (!P.training) {  // first input = prenet(zeros) (Modules.py:178-185)
    if (tid < 96) frame_s[tid] = 0.f;
    __syncthreads();
    for (int b = cid; b < B; b += nclusters) prenet_row(b, 0);
    grid_barrier(P.barrier, bar_target, gridDim.x);
  }

  for (int t = 0; t < P.T; ++t) {
    const uint8_t* zm = P.training ? P.zone_mask + (size_t)t * 4 * BC : nullptr;
    // ================= phase A: LSTM cell 0 =================
    for (int b0 = 0; b0 < B; b0 += NB) {
      const int nb = min(NB, B - b0);
      if (P.training) {
        lstm_gemv<NB>(P.W0r, D + kCell, P.ctx + ((size_t)t * B + b0) * D, D, P.hz0 + (size_t)t * BC + (size_t)b0 * kCell,
                      kCell, nb, red, unit0);
      } else {  // the prenet rows of cell 0's kernel are not hoisted: the frame is last step's projection
        const int col = (lane >> 3) * kCell + unit0 + (lane & 7);
        const int kslice = (D + kCell) >> 3;
        float acc[NB], acc2[NB];
        gemv_acc<NB>(P.W0r, kGates, col, warp * kslice, (warp + 1) * kslice, P.ctx + ((size_t)t * B + b0) * D, D,
                     P.hz0 + (size_t)t * BC + (size_t)b0 * kCell, kCell, nb, acc);
        gemv_acc<NB>(P.K0pre, kGates, col, warp * (kPrenet / 8), (warp + 1) * (kPrenet / 8),
                     P.pre + ((size_t)t * B + b0) * kPrenet, kPrenet, nullptr, 0, nb, acc2);
#pragma unroll
        for (int b = 0; b < NB; ++b) red[(warp * NB + b) * kRedStride + lane] = acc[b] + acc2[b];
      }
      __syncthreads();
      lstm_epilogue<NB>(red, nb, unit0, P.training ? P.g0pre + (size_t)t * BG + (size_t)b0 * kGates : nullptr, P.b0,
                        P.cz0 + (size_t)t * BC + (size_t)b0 * kCell, P.hz0 + (size_t)t * BC + (size_t)b0 * kCell,
                        zm ? zm + (size_t)b0 * kCell : nullptr, zm ? zm + BC + (size_t)b0 * kCell : nullptr,
                        P.act0 + (size_t)t * BG + (size_t)b0 * kGates, P.c0n + (size_t)t * BC + (size_t)b0 * kCell,
                        P.cz0 + (size_t)(t + 1) * BC + (size_t)b0 * kCell,
                        P.hz0 + (size_t)(t + 1) * BC + (size_t)b0 * kCell, P.m0 + (size_t)t * BC + (size_t)b0 * kCell,
                        nullptr);
      __syncthreads();
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);

    // ================= phase B: LSTM cell 1 (+ partial query projection) =================
    for (int b0 = 0; b0 < B; b0 += NB) {
      const int nb = min(NB, B - b0);
      lstm_gemv<NB>(P.W1, 2 * kCell, P.m0 + (size_t)t * BC + (size_t)b0 * kCell, kCell,
                    P.hz1 + (size_t)t * BC + (size_t)b0 * kCell, kCell, nb, red, unit0);
      __syncthreads();
      lstm_epilogue<NB>(red, nb, unit0, nullptr, P.b1, P.cz1 + (size_t)t * BC + (size_t)b0 * kCell,
                        P.hz1 + (size_t)t * BC + (size_t)b0 * kCell, zm ? zm + 2 * BC + (size_t)b0 * kCell : nullptr,
                        zm ? zm + 3 * BC + (size_t)b0 * kCell : nullptr, P.act1 + (size_t)t * BG + (size_t)b0 * kGates,
                        P.c1n + (size_t)t * BC + (size_t)b0 * kCell, P.cz1 + (size_t)(t + 1) * BC + (size_t)b0 * kCell,
                        P.hz1 + (size_t)(t + 1) * BC + (size_t)b0 * kCell, P.m1 + (size_t)t * BC + (size_t)b0 * kCell,
                        m_s);
      __syncthreads();
      // partial q[b][a] = sum over this CTA's 8 units of m1[b][unit] * Wq[unit][a]
      for (int i = tid; i < nb * kAtt; i += kDecThreads) {
        const int b = i >> 7, a = i & 127;
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < kUnitsPerCta; ++u) s = fmaf(m_s[b * kUnitsPerCta + u], wq_s[u * kAtt + a], s);
        P.qpart[((size_t)blockIdx.x * B + b0 + b) * kAtt + a] = s;
      }
      __syncthreads();
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);

    // ================= phase C: location-sensitive attention, one batch row per cluster ==========
    for (int b = cid; b < B; b += nclusters) {
      const int tl = min(P.text_len[b], Te);
      const float* keys_b = P.resident ? keys_s : P.keys + (size_t)b * Te * kAtt + crank * 32;
      const int kstride = P.resident ? 32 : kAtt;
      const float* vals_b = P.resident ? vals_s : P.values + (size_t)b * Te * D + crank * Dq;
      const int vstride = P.resident ? Dq : D;
      // previous cumulative alignment (slot t), shifted by 15 so cum_s[x] = cum[x-15]
      const float* cum_prev = P.cum + ((size_t)t * B + b) * Te;
      for (int x = tid; x < Te; x += kDecThreads) cum_s[15 + x] = cum_prev[x];
      // q slice = sum of the 128 per-CTA partials (fixed order -> identical in every run)
      {
        float s = 0.f;
        for (int j = warp; j < kDecGrid; j += 8) s += __ldcg(P.qpart + ((size_t)j * B + b) * kAtt + crank * 32 + lane);
        qred[warp * 32 + lane] = s;
      }
      __syncthreads();
      if (tid < 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += qred[w * 32 + tid];
        qf_s[tid] = s + fb_l;
        P.qf[((size_t)t * B + b) * kAtt + crank * 32 + tid] = s + fb_l;  // saved for the reverse pass
      }
      __syncthreads();
      // partial energies over this CTA's 32 attention units, 16 positions per warp pass
      const float qf = qf_s[lane];
      for (int blk = warp; blk * 16 < tl; blk += 8) {
        const int t0 = blk * 16;
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
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int x = t0 + p;
          float v = 0.f;
          if (x < tl) v = sw_l * tanhf(keys_b[(size_t)x * kstride + lane] + qf + acc[p]);
          v = warp_sum(v);
          if (lane == 0) e_loc[x] = v;
        }
      }
      __syncthreads();
      // all-gather the partial energies across the cluster through DSMEM
#pragma unroll
      for (int dst = 0; dst < kDecCluster; ++dst) {
        float* remote = cluster.map_shared_rank(e_parts, dst) + crank * TeP;
        for (int x = tid; x < tl; x += kDecThreads) remote[x] = e_loc[x];
      }
      cluster.sync();
      // masked softmax over positions < tl (score_mask_value = -inf => exactly 0 beyond tl)
      float lmax = -INFINITY;
      for (int x = tid; x < tl; x += kDecThreads) {
        const float e = ((e_parts[x] + e_parts[TeP + x]) + e_parts[2 * TeP + x]) + e_parts[3 * TeP + x];
        a_s[x] = e;
        lmax = fmaxf(lmax, e);
      }
      lmax = warp_max(lmax);
      if (lane == 0) bred[warp] = lmax;
      __syncthreads();
      float gmax = bred[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) gmax = fmaxf(gmax, bred[w]);
      float lsum = 0.f;
      for (int x = tid; x < tl; x += kDecThreads) {
        const float ex = expf(a_s[x] - gmax);
        a_s[x] = ex;
        lsum += ex;
      }
      lsum = warp_sum(lsum);
      if (lane == 0) bred[8 + warp] = lsum;
      __syncthreads();
      float gsum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) gsum += bred[8 + w];
      for (int x = tid; x < Te; x += kDecThreads) {
        const float a = (x < tl) ? a_s[x] / gsum : 0.f;
        a_s[x] = a;
        if (crank == 0) {
          P.align_tm[((size_t)t * B + b) * Te + x] = a;
          P.cum[((size_t)(t + 1) * B + b) * Te + x] = cum_s[15 + x] + a;
        }
      }
      __syncthreads();
      // context slice: ctx[d] = sum_x a[x] * values[x][d]
      if (tid < Dq) {
        float s = 0.f;
        for (int x = 0; x < tl; ++x) s = fmaf(a_s[x], vals_b[(size_t)x * vstride + tid], s);
        P.ctx[((size_t)(t + 1) * B + b) * D + crank * Dq + tid] = s;
        if (!P.training) x_s[256 + tid] = s;
      }
      if (!P.training) {
        // ---- projection [m1 | ctx] @ Wp (Modules.py:309-321), K split over the cluster, then the next prenet ----
        x_s[tid] = __ldcg(P.m1 + ((size_t)t * B + b) * kCell + crank * 256 + tid);
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int nrow = 256 + Dq;
        for (int r = warp; r < nrow; r += 8) {
          const int grow = r < 256 ? crank * 256 + r : kCell + crank * Dq + (r - 256);
          const float* wr = P.Wp + (size_t)grow * (kMel + 1);
          const float xv = x_s[r];
          a0 = fmaf(xv, __ldg(wr + lane), a0);
          a1 = fmaf(xv, __ldg(wr + 32 + lane), a1);
          if (lane < kMel + 1 - 64) a2 = fmaf(xv, __ldg(wr + 64 + lane), a2);
        }
        pred[warp * 96 + lane] = a0;
        pred[warp * 96 + 32 + lane] = a1;
        pred[warp * 96 + 64 + lane] = a2;
        __syncthreads();
        if (tid < 96) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += pred[w * 96 + tid];
#pragma unroll
          for (int dst = 0; dst < kDecCluster; ++dst) cluster.map_shared_rank(pparts, dst)[crank * 96 + tid] = s;
        }
        cluster.sync();
        if (tid < kMel + 1) {
          const float s = ((pparts[tid] + pparts[96 + tid]) + pparts[192 + tid]) + pparts[288 + tid];
          if (crank == 0) P.proj_tm[((size_t)t * B + b) * (kMel + 1) + tid] = s;
          frame_s[tid] = s + P.bp[tid];
        }
        __syncthreads();
        if (t + 1 < P.T) prenet_row(b, t + 1);
        cluster.sync();  // pparts / e_parts are reused by the next row
      } else if (b + nclusters < B) {
        cluster.sync();  // e_parts is reused by the next row
      }
    }
    grid_barrier(P.barrier, bar_target, gridDim.x);
    if (!P.training) {
      // finished |= stop >= 0 (Modules.py:216-219, OR-ed at :409); every CTA evaluates the same data => uniform exit.
      // The step cap (time >= Max_Inference_Length) is the loop bound T = cap + 1.
      int all = 1;
      for (int b = tid; b < B; b += kDecThreads) {
        const float st = __ldcg(P.proj_tm + ((size_t)t * B + b) * (kMel + 1) + kMel) + P.bp[kMel];
        if (st >= 0.f) fin_s[b] = 1;
        all &= fin_s[b];
      }
      all = __syncthreads_and(all);
      if (blockIdx.x == 0 && tid == 0) *P.steps_done = t + 1;
      if (all) break;
    }
  }
}

// ======================================== host side ================================================
size_t dec_fwd_smem_bytes(int NB, int Te, int D, int resident) {
  const int TeP = (Te + 15) & ~15;
  size_t f = (size_t)8 * NB * kRedStride + kUnitsPerCta * kAtt + NB * kUnitsPerCta + 8 * 32 + 32 + (TeP + 32) + TeP +
             kDecCluster * TeP + TeP + 16 + (512 + 8 * 96 + kDecCluster * 96 + 96 + kPrenet + 256 + 256);
  if (resident) f += (size_t)Te * 32 + (size_t)Te * (D / kDecCluster);
  return f * sizeof(float);
}

template <int NB>
static int launch_fwd(const DecFwdParams& P, cudaStream_t stream) {
  int dev = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  int max_optin = 0;
  MSTTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DecFwdParams Q = P;
  Q.resident = (P.B <= kDecGrid / kDecCluster) && dec_fwd_smem_bytes(NB, P.Te, P.D, 1) <= (size_t)max_optin;
  const size_t smem = dec_fwd_smem_bytes(NB, P.Te, P.D, Q.resident);
  MSTTS_REQUIRE(smem <= (size_t)max_optin, MSTTS_E_UNSUPPORTED, "decoder_fwd: Te=%d needs %zu B smem > %d", P.Te, smem,
                max_optin);
  MSTTS_CUDA(cudaFuncSetAttribute(decoder_fwd_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  MSTTS_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, decoder_fwd_kernel<NB>, &cfg));
  MSTTS_REQUIRE(nclusters * kDecCluster >= kDecGrid, MSTTS_E_DEVICE,
                "decoder_fwd: device co-schedules only %d clusters of %d (need %d)", nclusters, kDecCluster,
                kDecGrid / kDecCluster);
  mstts_timer_start(0, stream);
  MSTTS_CUDA(cudaLaunchKernelEx(&cfg, decoder_fwd_kernel<NB>, Q));
  mstts_timer_stop(0, stream);
  return MSTTS_OK;
}

int dec_fwd_persistent(const DecFwdParams& P, cudaStream_t stream) {
  const int B = P.B;
  if (B <= 1) return launch_fwd<1>(P, stream);
  if (B <= 2) return launch_fwd<2>(P, stream);
  if (B <= 4) return launch_fwd<4>(P, stream);
  if (B <= 8) return launch_fwd<8>(P, stream);
  if (B <= 16) return launch_fwd<16>(P, stream);
  return launch_fwd<32>(P, stream);
}

int dec_fwd_persistent_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws,
                             cudaStream_t s) {
  auto F = [&](size_t off) { return (float*)(ws + off); };
  DecFwdParams P;
  memset(&P, 0, sizeof(P));
  P.B = io->B; P.Te = io->Te; P.T = io->n_steps; P.D = io->D; P.training = io->is_training;
  P.W0r = F(l.W0r); P.W1 = w->cell1_kernel; P.b0 = w->cell0_bias; P.b1 = w->cell1_bias; P.Wq = w->query_kernel;
  P.F = F(l.locF); P.fb = F(l.locFb); P.sw = w->score_w;
  P.g0pre = F(l.g0pre); P.keys = F(l.keys); P.values = F(l.values);
  P.text_len = io->text_len; P.zone_mask = io->is_training ? io->zone_mask : nullptr;
  P.K0pre = w->cell0_kernel; P.Wp = w->proj_kernel; P.bp = w->proj_bias; P.P0 = w->prenet0_kernel; P.pb0 = w->prenet0_bias;
  P.P1 = w->prenet1_kernel; P.pb1 = w->prenet1_bias; P.prenet_mask = io->prenet_mask; P.pre = F(l.pre);
  P.proj_tm = F(l.proj_tm); P.steps_done = io->steps_done;
  P.act0 = F(l.act0); P.act1 = F(l.act1); P.c0n = F(l.c0n); P.c1n = F(l.c1n);
  P.cz0 = F(l.cz0); P.hz0 = F(l.hz0); P.cz1 = F(l.cz1); P.hz1 = F(l.hz1);
  P.m0 = F(l.m0); P.m1 = F(l.m1); P.ctx = F(l.ctx); P.cum = F(l.cum); P.align_tm = F(l.align_tm); P.qpart = F(l.qpart); P.qf = F(l.qf);
  P.barrier = (unsigned*)(ws + l.barrier);
  return dec_fwd_persistent(P, s);
}
