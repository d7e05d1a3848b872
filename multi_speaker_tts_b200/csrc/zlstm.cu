// Zoneout-LSTM over a whole sequence as ONE launch (H = 256): the encoder BiLSTM (Modules.py:49-73) and the speaker-embedding
// stack (Speaker_Embedding/Modules.py:12-37) -- the callers either side of the decoder loop.  Replaces
// tf.nn.dynamic_rnn / stack_bidirectional_dynamic_rnn over ZoneoutLSTMCell.call (ZoneoutLSTMCell.py:188-271): ~15 tiny
// launches per time step in an op-by-op port, none here.
//
// The input rows of the cell kernel are applied to all steps by one GEMM outside (xk = x Kx + bias); this kernel runs the
// recurrence.  Batch rows are independent, so the grid is one 8-CTA cluster per 8 batch rows and there is no grid-wide
// synchronisation: CTA r of a cluster owns hidden units 32r..32r+31 (its 128 gate columns of Kh stay resident in shared
// memory, 128 KB), the new h slice is pushed into the 8 CTAs' shared memory through DSMEM and one hardware cluster barrier
// closes the step.  Sequence lengths: beyond length[b] the state is carried and the output is zero; reverse = 1 walks each
// row from its last valid frame down (what tf reverse_sequence -> rnn -> reverse_sequence computes).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

constexpr int kZH = 256, kZCluster = 8, kZRows = 8, kZUnits = kZH / kZCluster;  // 32 units per CTA
constexpr int kZThreads = 256;

struct ZlstmParams {
  const float* xk;       // [B,T,4H]
  const float* kh;       // [H,4H]
  const int* lengths;    // [B]
  const uint8_t* masks;  // [T,2,B,H] (c, h) or NULL (inference: no mask, the keep factor stays)
  const float* x_res;    // [B,T,H] or NULL: ResidualWrapper (output = m + x)
  float* out;            // [B,T,H]
  float *acts, *c_prev, *h_prev;  // saved for the reverse pass (NULL: not saved): [B,T,4H], [B,T,H], [B,T,H]
  int B, T, reverse;
  float keep;
};

__global__ void __cluster_dims__(kZCluster, 1, 1) __launch_bounds__(kZThreads, 1) zlstm_fwd_kernel(const ZlstmParams P) {
  extern __shared__ __align__(16) float zsm[];
  cg::cluster_group cluster = cg::this_cluster();
  float* W_s = zsm;                           // [256 k][128 cols]  (col = gate*32 + unit)
  float* h_s = W_s + kZH * 128;               // [2][8 rows][256]
  float* g_s = h_s + 2 * kZRows * kZH;        // [8 rows][128]
  const int tid = threadIdx.x, crank = (int)cluster.block_rank();
  const int row0 = (blockIdx.x / kZCluster) * kZRows;
  for (int i = tid; i < kZH * 128; i += kZThreads) {
    const int k = i >> 7, c = i & 127;
    W_s[i] = P.kh[(size_t)k * 4 * kZH + (c >> 5) * kZH + crank * kZUnits + (c & 31)];
  }
  for (int i = tid; i < 2 * kZRows * kZH; i += kZThreads) h_s[i] = 0.f;
  // element owned for the cell update: (row r8, unit u)
  const int r8 = tid >> 5, u = tid & 31;
  const int brow = row0 + r8;
  const int len_own = brow < P.B ? min(max(P.lengths[brow], 0), P.T) : 0;
  int maxlen = 0;
  for (int r = 0; r < kZRows; ++r)
    if (row0 + r < P.B) maxlen = max(maxlen, min(max(P.lengths[row0 + r], 0), P.T));
  // GEMV mapping: column c (128) x row half (2): 4 rows per thread
  const int gc = tid & 127, rh = tid >> 7;
  int len_r[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) len_r[r] = (row0 + rh * 4 + r < P.B) ? min(max(P.lengths[row0 + rh * 4 + r], 0), P.T) : 0;
  float c_state = 0.f, h_state = 0.f;
  cluster.sync();
  for (int s = 0; s < maxlen; ++s) {
    const float* hcur = h_s + (s & 1) * kZRows * kZH;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // this step's global operands do not depend on the recurrence: issue the loads before the GEMV, use them after it
    float xin[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (s < len_r[r]) {
        const int t = P.reverse ? len_r[r] - 1 - s : s;
        xin[r] = __ldg(P.xk + ((size_t)(row0 + rh * 4 + r) * P.T + t) * 4 * kZH + (gc >> 5) * kZH + crank * kZUnits + (gc & 31));
      }
    }
    float mk_c = 1.f, mk_h = 1.f;
    if (P.masks && s < len_own) {
      const size_t mi = (((size_t)s * 2) * P.B + brow) * kZH + crank * kZUnits + u;
      mk_c = (float)P.masks[mi];
      mk_h = (float)P.masks[mi + (size_t)P.B * kZH];
    }
#pragma unroll 4
    for (int k = 0; k < kZH; k += 4) {
      const float w0 = W_s[(k + 0) * 128 + gc], w1 = W_s[(k + 1) * 128 + gc], w2 = W_s[(k + 2) * 128 + gc], w3 = W_s[(k + 3) * 128 + gc];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + (rh * 4 + r) * kZH + k);
        acc[r] = fmaf(hv.x, w0, acc[r]);
        acc[r] = fmaf(hv.y, w1, acc[r]);
        acc[r] = fmaf(hv.z, w2, acc[r]);
        acc[r] = fmaf(hv.w, w3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = rh * 4 + r;
      g_s[rr * 128 + gc] = acc[r] + xin[r];
    }
    __syncthreads();
    float h_pub = h_state;
    if (s < len_own) {
      const int t = P.reverse ? len_own - 1 - s : s;
      const int unit = crank * kZUnits + u;
      const float ig = sigmoidf_precise(g_s[r8 * 128 + u]);
      const float jg = tanhf(g_s[r8 * 128 + 32 + u]);
      const float fg = sigmoidf_precise(g_s[r8 * 128 + 64 + u] + kForgetBias);
      const float og = sigmoidf_precise(g_s[r8 * 128 + 96 + u]);
      const float cn = fg * c_state + ig * jg;
      const float m = og * tanhf(cn);
      const float dc = (cn - c_state) * mk_c, dm = (m - h_state) * mk_h;
      const size_t oi = ((size_t)brow * P.T + t) * kZH + unit;
      if (P.acts) {
        const size_t ai = ((size_t)brow * P.T + t) * 4 * kZH + unit;
        P.acts[ai] = ig;
        P.acts[ai + kZH] = jg;
        P.acts[ai + 2 * kZH] = fg;
        P.acts[ai + 3 * kZH] = og;
        P.c_prev[oi] = c_state;
        P.h_prev[oi] = h_state;
      }
      P.out[oi] = P.x_res ? m + P.x_res[oi] : m;
      c_state = P.keep * dc + c_state;
      h_state = P.keep * dm + h_state;
      h_pub = h_state;
    }
    // publish this CTA's slice of the next h to every CTA of the cluster: four neighbouring units per 16-byte remote store
    // (scalar remote stores, 8 per thread, were the largest part of the step)
    {
      const float h1 = __shfl_down_sync(0xffffffffu, h_pub, 1), h2 = __shfl_down_sync(0xffffffffu, h_pub, 2),
                  h3 = __shfl_down_sync(0xffffffffu, h_pub, 3);
      if ((u & 3) == 0) {
        float* hnext = h_s + ((s + 1) & 1) * kZRows * kZH + r8 * kZH + crank * kZUnits + u;
        const float4 v4 = make_float4(h_pub, h1, h2, h3);
#pragma unroll
        for (int dst = 0; dst < kZCluster; ++dst) *reinterpret_cast<float4*>(cluster.map_shared_rank(hnext, dst)) = v4;
      }
    }
    cluster.sync();
  }
}

// ---- reverse pass of the recurrence: d xk (= d gate pre-activations) for every step; the weight / input gradients are
//      GEMMs over d xk outside.  CTA r owns units 32r..: its rows of Kh (transposed: [1024 cols][32 units]) stay in smem. ----
struct ZlstmBwdParams {
  const float* dout;     // [B,T,H] gradient w.r.t. the cell output m (the residual path is handled by the caller)
  const float* kh;       // [H,4H]
  const int* lengths;
  const uint8_t* masks;
  const float *acts, *c_prev;
  float* dxk;            // [B,T,4H] (zero beyond the lengths)
  int B, T, reverse;
  float keep;
};

__global__ void __cluster_dims__(kZCluster, 1, 1) __launch_bounds__(kZThreads, 1) zlstm_bwd_kernel(const ZlstmBwdParams P) {
  extern __shared__ __align__(16) float zsm[];
  cg::cluster_group cluster = cg::this_cluster();
  float* WT_s = zsm;                          // [1024 cols][32 units]
  float* dg_s = WT_s + 4 * kZH * kZUnits;     // [2][1024 cols][8 rows]   full d gates of the cluster's rows
  float* red_s = dg_s + 2 * 4 * kZH * kZRows; // [8 col-groups][8 rows][32 units]
  const int tid = threadIdx.x, crank = (int)cluster.block_rank();
  const int row0 = (blockIdx.x / kZCluster) * kZRows;
  for (int i = tid; i < 4 * kZH * kZUnits; i += kZThreads) {
    const int col = i >> 5, uu = i & 31;
    WT_s[i] = P.kh[(size_t)(crank * kZUnits + uu) * 4 * kZH + col];
  }
  const int r8 = tid >> 5, u = tid & 31;
  const int brow = row0 + r8;
  const int len_own = brow < P.B ? min(max(P.lengths[brow], 0), P.T) : 0;
  int maxlen = 0;
  for (int r = 0; r < kZRows; ++r)
    if (row0 + r < P.B) maxlen = max(maxlen, min(max(P.lengths[row0 + r], 0), P.T));
  const int unit = crank * kZUnits + u;
  float dcz = 0.f, dhz = 0.f;  // gradients w.r.t. the zoned state leaving the current step
  const int cq = tid >> 5;     // GEMV: this warp handles columns [128 cq, +128), lane = unit
  // saved operands of a step do not depend on the recurrence: those of step s - 1 are fetched while step s runs its GEMV
  float pf_act[4] = {0.f, 0.f, 0.f, 0.f}, pf_cp = 0.f, pf_dout = 0.f, pf_mc = 1.f, pf_mh = 1.f;
  auto prefetch = [&](int s) {
    if (s < 0 || s >= len_own) return;
    const int t = P.reverse ? len_own - 1 - s : s;
    const size_t oi = ((size_t)brow * P.T + t) * kZH + unit, ai = ((size_t)brow * P.T + t) * 4 * kZH + unit;
#pragma unroll
    for (int g = 0; g < 4; ++g) pf_act[g] = __ldg(P.acts + ai + g * kZH);
    pf_cp = __ldg(P.c_prev + oi);
    pf_dout = __ldg(P.dout + oi);
    if (P.masks) {
      const size_t mi = (((size_t)s * 2) * P.B + brow) * kZH + unit;
      pf_mc = (float)P.masks[mi];
      pf_mh = (float)P.masks[mi + (size_t)P.B * kZH];
    }
  };
  prefetch(maxlen - 1);
  cluster.sync();
  for (int s = maxlen - 1; s >= 0; --s) {
    float dgv[4] = {0.f, 0.f, 0.f, 0.f};
    float dh_direct = dhz, dc_direct = dcz;
    const bool live = s < len_own;
    int t = 0;
    if (live) {
      t = P.reverse ? len_own - 1 - s : s;
      const float ig = pf_act[0], jg = pf_act[1], fg = pf_act[2], og = pf_act[3];
      const float cp = pf_cp;
      const float kc = P.keep * pf_mc, kh = P.keep * pf_mh;
      const float cn = fg * cp + ig * jg;
      const float tc = tanhf(cn);
      const float dm = pf_dout + dhz * kh;
      dh_direct = dhz * (1.f - kh);
      const float dcn = dm * og * (1.f - tc * tc) + dcz * kc;
      dc_direct = dcz * (1.f - kc) + dcn * fg;
      dgv[0] = dcn * jg * ig * (1.f - ig);
      dgv[1] = dcn * ig * (1.f - jg * jg);
      dgv[2] = dcn * cp * fg * (1.f - fg);
      dgv[3] = dm * tc * og * (1.f - og);
      float* dx = P.dxk + ((size_t)brow * P.T + t) * 4 * kZH + unit;
      dx[0] = dgv[0];
      dx[kZH] = dgv[1];
      dx[2 * kZH] = dgv[2];
      dx[3 * kZH] = dgv[3];
    }
    // all-gather the d gates of this step into every CTA: dg_s[buf][col][row].  The CTA's 128 columns x 8 rows are first laid
    // out [col][row] in local shared memory (red_s is idle here), then pushed as 16-byte vectors: 8 remote stores per thread
    // instead of 32 scalar ones
    float* dgb = dg_s + (s & 1) * 4 * kZH * kZRows;
#pragma unroll
    for (int g = 0; g < 4; ++g) red_s[(g * kZUnits + u) * kZRows + r8] = dgv[g];
    __syncthreads();
    {
      const int lc = tid >> 1, hf = tid & 1;  // local column (gate * 32 + unit), row half
      const float4 v4 = *reinterpret_cast<const float4*>(red_s + lc * kZRows + 4 * hf);
      float* p = dgb + (size_t)((lc >> 5) * kZH + crank * kZUnits + (lc & 31)) * kZRows + 4 * hf;
#pragma unroll
      for (int dst = 0; dst < kZCluster; ++dst) *reinterpret_cast<float4*>(cluster.map_shared_rank(p, dst)) = v4;
    }
    prefetch(s - 1);
    cluster.sync();
    // d h_prev[row][unit] = sum_col dG[row][col] Kh[unit][col]: warp cq sums its 128 columns for all 8 rows, lane = unit
    float acc[kZRows];
#pragma unroll
    for (int r = 0; r < kZRows; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 128; ++c) {
      const int col = cq * 128 + c;
      const float wv = WT_s[col * kZUnits + u];
      const float4 d0 = *reinterpret_cast<const float4*>(dgb + (size_t)col * kZRows);
      const float4 d1 = *reinterpret_cast<const float4*>(dgb + (size_t)col * kZRows + 4);
      acc[0] = fmaf(d0.x, wv, acc[0]);
      acc[1] = fmaf(d0.y, wv, acc[1]);
      acc[2] = fmaf(d0.z, wv, acc[2]);
      acc[3] = fmaf(d0.w, wv, acc[3]);
      acc[4] = fmaf(d1.x, wv, acc[4]);
      acc[5] = fmaf(d1.y, wv, acc[5]);
      acc[6] = fmaf(d1.z, wv, acc[6]);
      acc[7] = fmaf(d1.w, wv, acc[7]);
    }
#pragma unroll
    for (int r = 0; r < kZRows; ++r) red_s[(cq * kZRows + r) * kZUnits + u] = acc[r];
    __syncthreads();
    float dhp = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) dhp += red_s[(q * kZRows + r8) * kZUnits + u];
    __syncthreads();
    dhz = dh_direct + dhp;  // a finished row contributes zero d gates, so dhp is 0 for it
    dcz = dc_direct;
  }
}

// zero-fill helper for d xk rows beyond the lengths is the caller's cudaMemset
extern "C" int mstts_zlstm_fwd(const float* xk, const float* kh, const int32_t* lengths, const uint8_t* masks, const float* x_res, int B, int T,
                               int H, int reverse, float keep, float* out, float* acts, float* c_prev, float* h_prev, void* stream) {
  MSTTS_REQUIRE(xk && kh && lengths && out, MSTTS_E_INVALID, "zlstm_fwd: null pointer");
  MSTTS_REQUIRE(H == kZH, MSTTS_E_UNSUPPORTED, "zlstm: H=%d (only %d is built: encoder BiLSTM and speaker-embedding cells)", H, kZH);
  MSTTS_REQUIRE(B >= 1 && T >= 1, MSTTS_E_INVALID, "zlstm_fwd: B=%d T=%d", B, T);
  MSTTS_REQUIRE(!acts || (c_prev && h_prev), MSTTS_E_INVALID, "zlstm_fwd: acts needs c_prev and h_prev");
  cudaStream_t s = (cudaStream_t)stream;
  MSTTS_CUDA(cudaMemsetAsync(out, 0, (size_t)B * T * H * sizeof(float), s));  // zero beyond the sequence lengths
  ZlstmParams P;
  P.xk = xk; P.kh = kh; P.lengths = lengths; P.masks = masks; P.x_res = x_res; P.out = out; P.acts = acts; P.c_prev = c_prev;
  P.h_prev = h_prev; P.B = B; P.T = T; P.reverse = reverse; P.keep = keep;
  const size_t smem = (size_t)(kZH * 128 + 2 * kZRows * kZH + kZRows * 128) * sizeof(float);
  MSTTS_CUDA(cudaFuncSetAttribute(zlstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nclusters = (B + kZRows - 1) / kZRows;
  zlstm_fwd_kernel<<<nclusters * kZCluster, kZThreads, smem, s>>>(P);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

extern "C" int mstts_zlstm_bwd(const float* dout, const float* kh, const int32_t* lengths, const uint8_t* masks, const float* acts,
                               const float* c_prev, int B, int T, int H, int reverse, float keep, float* dxk, void* stream) {
  MSTTS_REQUIRE(dout && kh && lengths && acts && c_prev && dxk, MSTTS_E_INVALID, "zlstm_bwd: null pointer");
  MSTTS_REQUIRE(H == kZH, MSTTS_E_UNSUPPORTED, "zlstm: H=%d (only %d is built)", H, kZH);
  cudaStream_t s = (cudaStream_t)stream;
  MSTTS_CUDA(cudaMemsetAsync(dxk, 0, (size_t)B * T * 4 * H * sizeof(float), s));
  ZlstmBwdParams P;
  P.dout = dout; P.kh = kh; P.lengths = lengths; P.masks = masks; P.acts = acts; P.c_prev = c_prev; P.dxk = dxk; P.B = B; P.T = T;
  P.reverse = reverse; P.keep = keep;
  const size_t smem = (size_t)(4 * kZH * kZUnits + 2 * 4 * kZH * kZRows + 8 * kZRows * kZUnits) * sizeof(float);
  MSTTS_CUDA(cudaFuncSetAttribute(zlstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nclusters = (B + kZRows - 1) / kZRows;
  zlstm_bwd_kernel<<<nclusters * kZCluster, kZThreads, smem, s>>>(P);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
