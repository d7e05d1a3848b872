// Host orchestration of the decoder op: operand preparation, hoisted GEMMs, persistent-kernel launch,
// output projection.  (C ABI: mstts_decoder_workspace_bytes / mstts_decoder_fwd.)
#include "common.cuh"
#include "decoder_layout.h"
#include "gemm.h"
#include "scratch_pool.h"
#include "tc_gemm.h"

struct DecFwdParams;  // decoder_fwd.cu
int dec_fwd_persistent_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws,
                             cudaStream_t s);
int dec_fwd_tc_entry(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const DecLayout& l, char* ws, cudaStream_t s);  // decoder_fwd_tc.cu

// ------------------------------------------------------------------------------------------------
// small prep / epilogue kernels (all HBM-bound elementwise work, grid-stride, coalesced)
// ------------------------------------------------------------------------------------------------
// values = memory * sequence_mask(text_len)   (BahdanauAttention._prepare_memory)
__global__ void mask_memory_kernel(const float* __restrict__ mem, const int* __restrict__ text_len, float* __restrict__ out,
                                   int B, int Te, int D) {
  const size_t n = (size_t)B * Te * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / D;
    const int b = (int)(row / Te), x = (int)(row % Te);
    out[i] = (x < text_len[b]) ? mem[i] : 0.f;
  }
}

// W0r rows 0..D-1 = K0[256+r] + K0[256+D+r] (the context enters cell 0 twice, Modules.py:234 + default
// cell_input_fn); rows D.. = K0[256+2D+..] (h rows)
__global__ void fold_cell0_kernel(const float* __restrict__ K0, float* __restrict__ W0r, int D) {
  const size_t n = (size_t)(D + kCell) * kGates;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / kGates, c = i % kGates;
    float v;
    if (r < (size_t)D)
      v = K0[(kPrenet + r) * kGates + c] + K0[(kPrenet + D + r) * kGates + c];
    else
      v = K0[(kPrenet + 2 * (size_t)D + (r - D)) * kGates + c];
    W0r[i] = v;
  }
}

// F[k][u] = sum_c Wc[k][0][c] * Wd[c][u];  fb[u] = sum_c bc[c] * Wd[c][u] + bias_b[u]
__global__ void compose_location_kernel(const float* __restrict__ Wc, const float* __restrict__ bc,
                                        const float* __restrict__ Wd, const float* __restrict__ bias_b,
                                        float* __restrict__ F, float* __restrict__ fb) {
  const int u = threadIdx.x;  // 128 threads
  for (int k = blockIdx.x; k <= kConvK; k += gridDim.x) {
    float s = 0.f;
    if (k < kConvK) {
      for (int c = 0; c < kConvC; ++c) s = fmaf(Wc[k * kConvC + c], Wd[c * kAtt + u], s);
      F[k * kAtt + u] = s;
    } else {
      for (int c = 0; c < kConvC; ++c) s = fmaf(bc[c], Wd[c * kAtt + u], s);
      fb[u] = s + bias_b[u];
    }
  }
}

// frames[t][b][:] = (t == 0) ? 0 : mel[b][t-1][:]      (Modules.py:180, :222-230 in teacher-forced mode)
__global__ void shift_frames_kernel(const float* __restrict__ mel, float* __restrict__ frames, int B, int L, int T) {
  const size_t n = (size_t)T * B * kMel;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i % kMel);
    const size_t tb = i / kMel;
    const int b = (int)(tb % B), t = (int)(tb / B);
    frames[i] = (t == 0 || t - 1 >= L) ? 0.f : mel[((size_t)b * L + (t - 1)) * kMel + m];
  }
}

// y = relu(y + bias) / 0.5 * mask       (tf.layers.dense(relu) + tf.layers.dropout(rate=.5, training=True))
// mask layout [T][2][B][256]; y layout [T*B][256]
__global__ void prenet_act_kernel(float* __restrict__ y, const float* __restrict__ bias, const uint8_t* __restrict__ mask,
                                  int layer, int B, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kPrenet);
    const size_t tb = i / kPrenet;
    const size_t t = tb / B, b = tb % B;
    const float v = fmaxf(y[i] + bias[c], 0.f);
    y[i] = (v / 0.5f) * (float)mask[((t * 2 + layer) * B + b) * kPrenet + c];
  }
}

// proj_tm [T][B][81] (+bias) -> linear [B][T][80], stop [B][T]; align_tm [T][B][Te] -> align [B][T][Te]
__global__ void finish_outputs_kernel(const float* __restrict__ proj_tm, const float* __restrict__ bias,
                                      const float* __restrict__ align_tm, float* __restrict__ linear,
                                      float* __restrict__ stop, float* __restrict__ align, int B, int T, int Te) {
  const size_t n1 = (size_t)T * B * (kMel + 1);
  const size_t n2 = (size_t)T * B * Te;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n1 + n2; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n1) {
      const int c = (int)(i % (kMel + 1));
      const size_t tb = i / (kMel + 1);
      const size_t t = tb / B, b = tb % B;
      const float v = proj_tm[i] + bias[c];
      if (c < kMel)
        linear[(b * T + t) * kMel + c] = v;
      else
        stop[b * T + t] = v;
    } else {
      const size_t j = i - n1;
      const int x = (int)(j % Te);
      const size_t tb = j / Te;
      const size_t t = tb / B, b = tb % B;
      align[(b * T + t) * Te + x] = align_tm[j];
    }
  }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

static inline int ew_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
extern "C" size_t mstts_decoder_workspace_bytes(int B, int Te, int L, int D, int n_steps, int mode) {
  if (B <= 0 || Te <= 0 || D <= 0 || n_steps <= 0) return 0;
  if (dec_is_chunked(B, Te, mode)) return dec_chunk_plan(B, Te, L, D, n_steps, mode).total;
  return dec_layout(B, Te, L, D, n_steps, mode).total;
}

// offset (bytes) of a named workspace region, or (size_t)-1: lets tests / profiling tools look at saved
// activations and the phase time stamps without knowing the layout
extern "C" size_t mstts_decoder_ws_offset(const char* name, int B, int Te, int L, int D, int n_steps, int mode) {
  if (!name || B <= 0 || Te <= 0 || D <= 0 || n_steps <= 0) return (size_t)-1;
  if (dec_is_chunked(B, Te, mode)) return (size_t)-1;  // several chunk workspaces: no single region
  const DecLayout l = dec_layout(B, Te, L, D, n_steps, mode);
#define REGION(x) if (!strcmp(name, #x)) return l.x;
  REGION(values) REGION(keys) REGION(g0pre) REGION(act0) REGION(act1) REGION(c0n) REGION(c1n) REGION(cz0) REGION(hz0)
  REGION(cz1) REGION(hz1) REGION(m0) REGION(m1) REGION(ctx) REGION(cum) REGION(align_tm) REGION(qf) REGION(dG0) REGION(dG1)
  REGION(dctx) REGION(dq) REGION(dbg) REGION(dbg_b) REGION(wimg_f) REGION(total)
  REGION(frames) REGION(pre_h) REGION(pre) REGION(dpre) REGION(dpre_h) REGION(dproj_tm) REGION(dm1_proj) REGION(proj_tm)
#undef REGION
  return (size_t)-1;
}

static int check_io(const MsttsDecoderWeights* w, const MsttsDecoderIO* io) {
  MSTTS_REQUIRE(w && io, MSTTS_E_INVALID, "decoder: null weights/io");
  MSTTS_REQUIRE(io->B >= 1 && io->B <= 256, MSTTS_E_INVALID, "decoder: B=%d out of range [1,256]", io->B);
  MSTTS_REQUIRE(io->Te >= 1 && io->Te <= 2048, MSTTS_E_INVALID, "decoder: Te=%d out of range", io->Te);
  MSTTS_REQUIRE(io->D >= 32 && io->D % 32 == 0 && io->D <= 1024, MSTTS_E_INVALID,
                "decoder: memory depth D=%d must be a multiple of 32 in [32,1024]", io->D);
  MSTTS_REQUIRE(io->n_steps >= 1, MSTTS_E_INVALID, "decoder: n_steps=%d", io->n_steps);
  MSTTS_REQUIRE(io->mode == MSTTS_MODE_FP32 || io->mode == MSTTS_MODE_BF16X3, MSTTS_E_UNSUPPORTED,
                "decoder: mode %d not implemented (fp32 = 0, bf16x3 = 1)", io->mode);
  MSTTS_REQUIRE(io->memory && io->text_len && io->prenet_mask && io->linear && io->stop && io->align, MSTTS_E_INVALID,
                "decoder: null io pointer");
  const float* const* wp = reinterpret_cast<const float* const*>(w);
  for (size_t i = 0; i < sizeof(MsttsDecoderWeights) / sizeof(float*); ++i)
    MSTTS_REQUIRE(wp[i], MSTTS_E_INVALID, "decoder: null weight pointer #%zu", i);
  return MSTTS_OK;
}

// Free-running (inference) decode, Modules.py:212-237: nothing can be hoisted except the memory layer; the persistent
// fp32 kernel computes projection -> prenet -> cells -> attention per step and stops when every row is finished.
// n_steps = step cap + 1 (Max_Inference_Length + 1); outputs beyond *steps_done are zero.
static int decoder_fwd_free_running(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws_, size_t ws_bytes,
                                    cudaStream_t s) {
  MSTTS_REQUIRE(io->steps_done, MSTTS_E_INVALID, "decoder: free-running mode needs steps_done");
  const bool tc = io->mode == MSTTS_MODE_BF16X3;  // tcgen05 loop with the projection + prenet inside (decoder_fwd_tc.cu)
  MSTTS_REQUIRE(!tc || (io->B <= 32 && io->Te <= 128 && io->D % 256 == 0 && io->D <= 768), MSTTS_E_UNSUPPORTED,
                "decoder: free-running bf16x3 mode needs B<=32, Te<=128, D in {256,512,768} (got B=%d Te=%d D=%d); use mode fp32", io->B, io->Te,
                io->D);
  const int B = io->B, Te = io->Te, D = io->D, T = io->n_steps;
  const DecLayout l = dec_layout(B, Te, io->L, D, T, io->mode);
  MSTTS_REQUIRE(ws_ && ws_bytes >= l.total, MSTTS_E_WORKSPACE, "decoder: workspace %zu < %zu", ws_bytes, l.total);
  MSTTS_REQUIRE(((uintptr_t)ws_ & 255) == 0, MSTTS_E_INVALID, "decoder: workspace must be 256-byte aligned");
  char* ws = (char*)ws_;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  const size_t TB = (size_t)T * B;
  mask_memory_kernel<<<ew_grid((size_t)B * Te * D), 256, 0, s>>>(io->memory, io->text_len, F(l.values), B, Te, D);
  int rc = gemm_rowmajor(s, B * Te, kAtt, D, F(l.values), D, w->memory_kernel, kAtt, F(l.keys), kAtt, 0.f);
  if (rc) return rc;
  fold_cell0_kernel<<<148 * 8, 256, 0, s>>>(w->cell0_kernel, F(l.W0r), D);
  compose_location_kernel<<<32, kAtt, 0, s>>>(w->loc_conv_kernel, w->loc_conv_bias, w->loc_dense_kernel, w->score_b,
                                              F(l.locF), F(l.locFb));
  const size_t BC = (size_t)B * kCell * sizeof(float);
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cz0, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.hz0, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cz1, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.hz1, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.ctx, 0, (size_t)B * D * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cum, 0, (size_t)B * Te * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.barrier, 0, 64, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.proj_tm, 0, TB * (kMel + 1) * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.align_tm, 0, TB * Te * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(io->steps_done, 0, sizeof(int), s));
  rc = tc ? dec_fwd_tc_entry(w, io, l, ws, s) : dec_fwd_persistent_entry(w, io, l, ws, s);
  if (rc) return rc;
  finish_outputs_kernel<<<ew_grid(TB * (kMel + 1 + Te)), 256, 0, s>>>(F(l.proj_tm), w->proj_bias, F(l.align_tm), io->linear,
                                                                    io->stop, io->align, B, T, Te);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

// [m1 | ctx_t] . Wp over all steps as ONE product: the two activation blocks are packed side by side along K
static int dec_projection(cudaStream_t s, const float* m1, const float* ctx, const float* Wp, float* out, int TB, int D) {
  const int K = kCell + D, NP = kMel + 1, Kb = (K + 63) / 64;
  if (D % 64 != 0) {  // k-block aligned blocks only; otherwise two accumulating products
    int rc = gemm_rowmajor(s, TB, NP, kCell, m1, kCell, Wp, NP, out, NP, 0.f);
    if (rc) return rc;
    return gemm_rowmajor(s, TB, NP, D, ctx, D, Wp + (size_t)kCell * NP, NP, out, NP, 1.f);
  }
  ScratchScope sc(s);
  void *ai = nullptr, *bi = nullptr;
  int rc;
  if ((rc = sc.get(&ai, tc_image_bytes(TB, K, 128)))) return rc;
  if ((rc = sc.get(&bi, tc_image_bytes(NP, K, 256)))) return rc;
  if ((rc = tc_pack_f32(s, m1, kCell, false, TB, kCell, 128, Kb, ai, 0, 0))) return rc;
  if ((rc = tc_pack_f32(s, ctx, D, false, TB, D, 128, Kb, ai, 0, kCell / 64))) return rc;
  if ((rc = tc_pack_f32(s, Wp, NP, true, NP, K, 256, Kb, bi, 0, 0))) return rc;
  return tc_gemm_images(s, ai, bi, TB, NP, K, out, NP, 0.f);
}

// rows [b0, b0 + bc) of a time-major byte tensor [outer][B][inner] -> contiguous [outer][bc][inner]
__global__ void gather_batch_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t outer, int B, int b0, int bc,
                                         int inner16) {
  const size_t n = outer * bc * inner16;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t c = i % inner16, ob = i / inner16;
    const size_t o = ob / bc, b = ob % bc;
    d4[i] = s4[(o * B + b0 + b) * inner16 + c];
  }
}

static int decoder_fwd_one(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws_, size_t ws_bytes, cudaStream_t s);

// B > 32 in bf16x3 mode: balanced row chunks through the one-tile path (decoder_layout.h: DecChunkPlan)
static int decoder_fwd_chunked(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws_, size_t ws_bytes, cudaStream_t s) {
  const int B = io->B, Te = io->Te, L = io->L, D = io->D, T = io->n_steps;
  const DecChunkPlan p = dec_chunk_plan(B, Te, L, D, T, io->mode);
  MSTTS_REQUIRE(ws_ && ws_bytes >= p.total, MSTTS_E_WORKSPACE, "decoder: workspace %zu < %zu", ws_bytes, p.total);
  MSTTS_REQUIRE(((uintptr_t)ws_ & 255) == 0, MSTTS_E_INVALID, "decoder: workspace must be 256-byte aligned");
  MSTTS_REQUIRE(((uintptr_t)io->prenet_mask & 15) == 0 && ((uintptr_t)io->zone_mask & 15) == 0, MSTTS_E_INVALID,
                "decoder: masks must be 16-byte aligned");
  char* ws = (char*)ws_;
  for (int c = 0; c < p.nchunks; ++c) {
    const int b0 = c * p.bc, bc = (B - b0 < p.bc) ? B - b0 : p.bc;
    uint8_t* pm = (uint8_t*)(ws + p.pm_off + (size_t)c * p.pm_bytes);
    uint8_t* zm = (uint8_t*)(ws + p.zm_off + (size_t)c * p.zm_bytes);
    gather_batch_rows_kernel<<<ew_grid((size_t)T * 2 * bc * kPrenet / 16), 256, 0, s>>>(io->prenet_mask, pm, (size_t)T * 2, B, b0, bc, kPrenet / 16);
    gather_batch_rows_kernel<<<ew_grid((size_t)T * 4 * bc * kCell / 16), 256, 0, s>>>(io->zone_mask, zm, (size_t)T * 4, B, b0, bc, kCell / 16);
    MsttsDecoderIO sub = *io;
    sub.B = bc;
    sub.memory = io->memory + (size_t)b0 * Te * D;
    sub.text_len = io->text_len + b0;
    sub.mel = io->mel + (size_t)b0 * L * kMel;
    sub.mel_len = io->mel_len + b0;
    sub.prenet_mask = pm;
    sub.zone_mask = zm;
    sub.linear = io->linear + (size_t)b0 * T * kMel;
    sub.stop = io->stop + (size_t)b0 * T;
    sub.align = io->align + (size_t)b0 * T * Te;
    int rc = decoder_fwd_one(w, &sub, ws + (size_t)c * p.chunk_ws, p.chunk_ws, s);
    if (rc) return rc;
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}

extern "C" int mstts_decoder_fwd(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws_, size_t ws_bytes,
                                 void* stream_) {
  int rc = check_io(w, io);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream_;
  if (!io->is_training) return decoder_fwd_free_running(w, io, ws_, ws_bytes, s);
  MSTTS_REQUIRE(io->mel && io->mel_len && io->zone_mask, MSTTS_E_INVALID, "decoder: training needs mel/mel_len/zone_mask");
  MSTTS_REQUIRE(io->n_steps <= io->L + 1, MSTTS_E_INVALID, "decoder: n_steps=%d > L+1=%d", io->n_steps, io->L + 1);
  if (dec_is_chunked(io->B, io->Te, io->mode)) return decoder_fwd_chunked(w, io, ws_, ws_bytes, s);
  return decoder_fwd_one(w, io, ws_, ws_bytes, s);
}

static int decoder_fwd_one(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws_, size_t ws_bytes, cudaStream_t s) {
  int rc;
  MSTTS_REQUIRE(io->mel && io->mel_len && io->zone_mask, MSTTS_E_INVALID, "decoder: training needs mel/mel_len/zone_mask");
  MSTTS_REQUIRE(io->n_steps <= io->L + 1, MSTTS_E_INVALID, "decoder: n_steps=%d > L+1=%d", io->n_steps, io->L + 1);
  const int B = io->B, Te = io->Te, L = io->L, D = io->D, T = io->n_steps;
  const DecLayout l = dec_layout(B, Te, L, D, T, io->mode);
  MSTTS_REQUIRE(ws_ && ws_bytes >= l.total, MSTTS_E_WORKSPACE, "decoder: workspace %zu < %zu", ws_bytes, l.total);
  MSTTS_REQUIRE(((uintptr_t)ws_ & 255) == 0, MSTTS_E_INVALID, "decoder: workspace must be 256-byte aligned");
  char* ws = (char*)ws_;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  const size_t TB = (size_t)T * B;

  // ---- operand preparation ----
  mask_memory_kernel<<<ew_grid((size_t)B * Te * D), 256, 0, s>>>(io->memory, io->text_len, F(l.values), B, Te, D);
  rc = gemm_rowmajor(s, B * Te, kAtt, D, F(l.values), D, w->memory_kernel, kAtt, F(l.keys), kAtt, 0.f);
  if (rc) return rc;
  fold_cell0_kernel<<<148 * 8, 256, 0, s>>>(w->cell0_kernel, F(l.W0r), D);
  compose_location_kernel<<<32, kAtt, 0, s>>>(w->loc_conv_kernel, w->loc_conv_bias, w->loc_dense_kernel, w->score_b,
                                              F(l.locF), F(l.locFb));
  // ---- hoisted prenet + prenet rows of cell 0 ----
  shift_frames_kernel<<<ew_grid(TB * kMel), 256, 0, s>>>(io->mel, F(l.frames), B, L, T);
  rc = gemm_rowmajor(s, (int)TB, kPrenet, kMel, F(l.frames), kMel, w->prenet0_kernel, kPrenet, F(l.pre_h), kPrenet, 0.f);
  if (rc) return rc;
  prenet_act_kernel<<<ew_grid(TB * kPrenet), 256, 0, s>>>(F(l.pre_h), w->prenet0_bias, io->prenet_mask, 0, B, TB * kPrenet);
  rc = gemm_rowmajor(s, (int)TB, kPrenet, kPrenet, F(l.pre_h), kPrenet, w->prenet1_kernel, kPrenet, F(l.pre), kPrenet, 0.f);
  if (rc) return rc;
  prenet_act_kernel<<<ew_grid(TB * kPrenet), 256, 0, s>>>(F(l.pre), w->prenet1_bias, io->prenet_mask, 1, B, TB * kPrenet);
  // prenet rows of cell 0's kernel over all steps (54 GFLOP at config 2)
  rc = gemm_rowmajor_ex(s, false, false, (int)TB, kGates, kPrenet, F(l.pre), kPrenet, w->cell0_kernel, kGates, F(l.g0pre), kGates, 0.f);
  if (rc) return rc;
  // ---- zero initial state (AttentionWrapper.zero_state, Modules.py:112) and the barrier counter ----
  const size_t BC = (size_t)B * kCell * sizeof(float);
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cz0, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.hz0, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cz1, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.hz1, 0, BC, s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.ctx, 0, (size_t)B * D * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.cum, 0, (size_t)B * Te * sizeof(float), s));
  MSTTS_CUDA(cudaMemsetAsync(ws + l.barrier, 0, 64, s));
  // ---- the loop ----
  rc = io->mode == MSTTS_MODE_BF16X3 ? dec_fwd_tc_entry(w, io, l, ws, s) : dec_fwd_persistent_entry(w, io, l, ws, s);
  if (rc) return rc;
  // ---- hoisted projection: [m1 | ctx] @ Wp + bp over all steps (Modules.py:292-294,309-321) ----
  rc = dec_projection(s, F(l.m1), F(l.ctx) + (size_t)B * D, w->proj_kernel, F(l.proj_tm), (int)TB, D);
  if (rc) return rc;
  finish_outputs_kernel<<<ew_grid(TB * (kMel + 1 + Te)), 256, 0, s>>>(F(l.proj_tm), w->proj_bias, F(l.align_tm), io->linear,
                                                                    io->stop, io->align, B, T, Te);
  if (io->steps_done) set_int_kernel<<<1, 1, 0, s>>>(io->steps_done, T);
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
