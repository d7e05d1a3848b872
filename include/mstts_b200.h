/*
 * mstts_b200.h -- C ABI of libmstts_b200.so, the sm_100a hot path of CODEJIN/multi_speaker_tts.
 *
 * The reference is pure Python/TF1 and has no FFI: the seam it offers is its Python module surface
 * (SURVEY.md 8b).  Each entry point below replaces the TF graph section named beside it; the Python
 * mirror of that surface (multi_speaker_tts_b200/Modules.py, Audio.py, WaveGlow/Modules.py) binds
 * these symbols through ctypes.
 *
 * Conventions
 *   - extern "C"; every function returns int: 0 = ok, <0 = MSTTS_E_* (mstts_last_error() gives a
 *     thread-local message).
 *   - all data pointers are DEVICE pointers owned by the caller, contiguous, fp32 unless noted;
 *     the library never allocates or frees caller memory.  Scratch ("workspace") is passed in with
 *     its size; query it with the matching *_workspace_bytes().
 *   - every call takes a cudaStream_t (as void*) and is asynchronous on it.
 *   - no torch types, no C++ types in signatures.
 */
#ifndef MSTTS_B200_H_
#define MSTTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSTTS_VERSION 100 /* 0.1.0 */

enum {
  MSTTS_OK = 0,
  MSTTS_E_INVALID = -1,     /* bad argument (null pointer, size out of range) */
  MSTTS_E_WORKSPACE = -2,   /* workspace too small */
  MSTTS_E_CUDA = -3,        /* a CUDA runtime call failed */
  MSTTS_E_UNSUPPORTED = -4, /* configuration not implemented (e.g. conv stride != 1) */
  MSTTS_E_DEVICE = -5       /* device is not sm_100 or cannot co-schedule the persistent grid */
};

/* precision modes of the recurrent GEMVs */
enum {
  MSTTS_MODE_FP32 = 0,  /* fp32 weights, fp32 FMA: the parity mode (L_inf vs oracle ~1e-5) */
  MSTTS_MODE_BF16X3 = 1,/* tcgen05, weights and activations split hi+lo bf16, 3 MMAs, fp32 accum */
  MSTTS_MODE_BF16 = 2   /* tcgen05, bf16 in / fp32 accum (fast mode; does not meet the 1e-3 gate) */
};

int mstts_version(void);
const char* mstts_last_error(void);
/* number of SMs / whether the device can run the persistent cluster grid; <0 on error */
int mstts_device_check(int device);
/* measurement hooks: with profiling on, CUDA events bracket every launch of the persistent kernels on the
 * caller's stream (no synchronisation); mstts_kernel_ms(which) waits for them and returns the summed
 * device time and the launch count (0 = decoder forward loop, 1 = decoder reverse loop) for this thread. */
int mstts_set_profiling(int on);                       /* also resets the recorded launches */
int mstts_kernel_ms(int which, float* sum_ms, int* count);

/* ------------------------------------------------------------------------------------------------
 * Tacotron2 decoder loop.
 * Replaces: Modules.Decoder_LSTM (Modules.py:76-119) = Decoder_Helper (:148-255) + ZoneoutLSTMCell x2
 * (ZoneoutLSTMCell.py:188-271) + Location_Sensitive_Attention (Location_Sensitive_Attention.py:43-85,
 * incl. the BahdanauAttention memory_layer built at :36-41) + Decoder_Decoder.projection
 * (Modules.py:309-321) + Decoder_Dynamic_Decode (:323-472).
 * Weights are in the TF variable layouts ([in, out] row-major; conv kernel [k, in, out]).
 * ---------------------------------------------------------------------------------------------- */
typedef struct MsttsDecoderWeights {
  const float* prenet0_kernel; /* [80,256]   decoder/decoder/prenet_0/dense/kernel */
  const float* prenet0_bias;   /* [256] */
  const float* prenet1_kernel; /* [256,256] */
  const float* prenet1_bias;   /* [256] */
  const float* cell0_kernel;   /* [256+2*D+1024, 4096]  rows: prenet | ctx | ctx | h ; cols: i j f o */
  const float* cell0_bias;     /* [4096] */
  const float* cell1_kernel;   /* [2048, 4096] */
  const float* cell1_bias;     /* [4096] */
  const float* memory_kernel;  /* [D,128]    attention/memory_layer/kernel (no bias) */
  const float* query_kernel;   /* [1024,128] query_layer/kernel (no bias) */
  const float* loc_conv_kernel;/* [31,1,32] */
  const float* loc_conv_bias;  /* [32] */
  const float* loc_dense_kernel;/* [32,128] (no bias) */
  const float* score_w;        /* [128] weight_w */
  const float* score_b;        /* [128] bias_b */
  const float* proj_kernel;    /* [1024+D, 81] */
  const float* proj_bias;      /* [81] */
} MsttsDecoderWeights;

/* Same fields, gradient outputs (fp32, caller-allocated, OVERWRITTEN not accumulated). */
typedef struct MsttsDecoderWeightGrads {
  float* prenet0_kernel; float* prenet0_bias; float* prenet1_kernel; float* prenet1_bias;
  float* cell0_kernel; float* cell0_bias; float* cell1_kernel; float* cell1_bias;
  float* memory_kernel; float* query_kernel;
  float* loc_conv_kernel; float* loc_conv_bias; float* loc_dense_kernel;
  float* score_w; float* score_b; float* proj_kernel; float* proj_bias;
} MsttsDecoderWeightGrads;

typedef struct MsttsDecoderIO {
  int B;            /* batch */
  int Te;           /* padded text length (memory time axis) */
  int L;            /* padded mel length (teacher-forcing frames) */
  int D;            /* memory depth (768 = 2*256 encoder + 256 speaker) */
  int n_steps;      /* training: max(mel_len)+1 (Modules.py:215,395).  inference: step cap (1000)+1 */
  int is_training;  /* 1: teacher forcing + zoneout masks; 0: free running (prenet dropout stays on) */
  int mode;         /* MSTTS_MODE_* */
  const float* memory;        /* [B,Te,D] encoder output ++ speaker embedding, un-masked */
  const int32_t* text_len;    /* [B] */
  const float* mel;           /* [B,L,80] */
  const int32_t* mel_len;     /* [B] */
  const uint8_t* prenet_mask; /* [n_steps,2,B,256] 0/1, step t consumes [t] (always applied) */
  const uint8_t* zone_mask;   /* [n_steps,2,2,B,1024] 0/1 ([t][cell][c|h]); NULL at inference */
  float* linear;              /* out [B,n_steps,80] */
  float* stop;                /* out [B,n_steps]    (logits) */
  float* align;               /* out [B,n_steps,Te] */
  int32_t* steps_done;        /* out [1] device: number of executed steps (== n_steps in training) */
} MsttsDecoderIO;

/* upstream gradients / data gradients of the decoder op (SURVEY A-11) */
typedef struct MsttsDecoderGrads {
  const float* d_linear;  /* in  [B,n_steps,80] */
  const float* d_stop;    /* in  [B,n_steps] */
  float* d_memory;        /* out [B,Te,D]  (through values and keys; zero beyond text_len) */
} MsttsDecoderGrads;

size_t mstts_decoder_workspace_bytes(int B, int Te, int L, int D, int n_steps, int mode);
/* byte offset of a named workspace region ("m1", "ctx", "dbg", ...; see decoder_layout.h) or (size_t)-1:
 * lets tests and profiling tools read saved activations / phase time stamps */
size_t mstts_decoder_ws_offset(const char* name, int B, int Te, int L, int D, int n_steps, int mode);
int mstts_decoder_fwd(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, void* ws, size_t ws_bytes,
                      void* stream);
/* must be called with the SAME io / ws that mstts_decoder_fwd filled (saved activations live in ws) */
int mstts_decoder_bwd(const MsttsDecoderWeights* w, const MsttsDecoderIO* io, const MsttsDecoderGrads* g,
                      const MsttsDecoderWeightGrads* dw, void* ws, size_t ws_bytes, void* stream);

/* decoder part of the loss, MSTTS_SV.py:127-144: linear MSE(+L1) on linear[:, :-1] vs mel (un-masked
 * mean) and stop BCE-with-logits vs 1-sequence_mask(mel_len, n_steps).  Writes loss[0]=linear,
 * loss[1]=stop and the gradients of (linear_loss + stop_loss) w.r.t. linear / stop. */
int mstts_decoder_loss(const float* linear, const float* stop, const float* mel, const int32_t* mel_len,
                       int B, int L, int n_steps, int use_l1, float* loss2, float* d_linear, float* d_stop,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Random masks: counter-based generator so dropout / zoneout bits never cross PCIe.
 * Replaces tf.layers.dropout's / ZoneoutLSTMCell.dropout_no_scale's floor(U + keep) (Modules.py:248-253,
 * ZoneoutLSTMCell.py:266-271).  out[i] = (hash(seed, i) < keep) ? 1 : 0.
 * ---------------------------------------------------------------------------------------------- */
int mstts_fill_mask(uint8_t* out, size_t n, float keep_prob, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tf.train.AdamOptimizer update, epsilon-hat form (MSTTS_SV.py:171-176): flat fp32 buffers.
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller.  The effective gradient is
 * g*grad_scale + l2*p: grad_scale = 1/world_size after the allreduce (or the clip factor for WaveGlow),
 * l2 = Weight_Regularization_Rate for the tensors in the reference's regularised set (MSTTS_SV.py:145-159).
 * ---------------------------------------------------------------------------------------------- */
int mstts_adam_tf(float* p, float* m, float* v, const float* g, size_t n, float lr_t, float b1, float b2,
                  float eps, float grad_scale, float l2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WaveGlow affine-coupling flows.
 * Replaces: WaveGlow.Modules.Glow_Train / Glow_Inference (WaveGlow/Modules.py:329-371) = 12 x Affine_Coupling_Layer
 * (:210-250) = Invertible 1x1 (WaveGlow/Inv1x1.py:9-32) + WaveNet (:252-327) with weight-normalised convs (:9-33).
 * Raw variables in TF layouts: weight-normed conv = (g [out], v [k,in,out], b [out]); end conv plain [1,512,c]; c = 8,6,4.
 * ---------------------------------------------------------------------------------------------- */
typedef struct MsttsWaveGlowWeights {
  const float* inv_w[12];     /* [c,c]; direction 1 expects the INVERSE kernel here */
  const float* start_g[12];   const float* start_v[12];   const float* start_b[12];     /* audio_initial_conv [1,c/2,512] */
  const float* in_g[12][8];   const float* in_v[12][8];   const float* in_b[12][8];     /* audio_in_i  [3,512,1024] */
  const float* cond_g[12][8]; const float* cond_v[12][8]; const float* cond_b[12][8];   /* mel_cond_i  [1,640,1024] */
  const float* res_g[12][8];  const float* res_v[12][8];  const float* res_b[12][8];    /* res_i [1,512,1024] (last: 512) */
  const float* end_w[12];     const float* end_b[12];                                   /* conv1d [1,512,c] (not weight-normed) */
} MsttsWaveGlowWeights;

size_t mstts_waveglow_workspace_bytes(int N, int T);
/* direction 0 (Glow_Train): audio_in [N,T,8] -> out z [N,T,8] (early outputs first, Modules.py:348-350),
 *   sums[0] = sum over flows of sum(log_s), sums[1] = sum(z^2)  (device doubles; log-det of the 1x1 kernels is host math).
 * direction 1 (Glow_Inference): audio_in = z [N,T,4], early_noise[0] / [1] = the [N,T,2] tensors concatenated after
 *   undoing flows 8 / 4 (already scaled by sigma) -> out [N,T,8].
 * mel_nt640: up-sampled, cropped and folded conditioning [N,T,640] (Restructure_*_Data, Modules.py:135-195). */
int mstts_waveglow_flows(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T,
                         int direction, const float* const* early_noise, float* out, double* sums, void* ws,
                         size_t ws_bytes, void* stream);
/* A/B switch for measurements: 1 = run the WN contractions of mstts_waveglow_flows through the row-major front-end of the
 * hand-written GEMM with separate gate / residual-skip kernels (the path the training forward uses) instead of the fused-epilogue
 * path over tile images (default 0).  Both are the same tcgen05 kernel; no vendor library is involved in either. */
int mstts_waveglow_set_path(int library_gemm);

/* Training (WaveGlow/WaveGlow.py:48-70): forward in the training direction that keeps every layer's operands in the workspace,
 * and the reverse pass for L = -sum(log_s)/n - sum(logdet W)/n + sum(z^2)/(2 sigma^2 n), n = N*T*8 (Modules.py:373-384).
 * Gradients are written (not accumulated) in the layouts of the raw variables; the logdet term of the twelve c x c kernels is
 * host math for the caller, as in the forward direction.  d_mel_nt640 (may be NULL): gradient w.r.t. the conditioning
 * [N,T,640], i.e. w.r.t. the up-sampled mel [N, 8T, 80].  fwd and bwd must use the same workspace, back to back. */
typedef struct MsttsWaveGlowGrads {
  float* inv_w[12];
  float* start_g[12];   float* start_v[12];   float* start_b[12];
  float* in_g[12][8];   float* in_v[12][8];   float* in_b[12][8];
  float* cond_g[12][8]; float* cond_v[12][8]; float* cond_b[12][8];
  float* res_g[12][8];  float* res_v[12][8];  float* res_b[12][8];
  float* end_w[12];     float* end_b[12];
} MsttsWaveGlowGrads;
size_t mstts_waveglow_train_workspace_bytes(int N, int T);
int mstts_waveglow_train_fwd(const MsttsWaveGlowWeights* w, const float* audio_in, const float* mel_nt640, int N, int T, float* z,
                             double* sums, void* ws, size_t ws_bytes, void* stream);
int mstts_waveglow_train_bwd(const MsttsWaveGlowWeights* w, const MsttsWaveGlowGrads* dw, const float* z, int N, int T, float sigma,
                             float* d_mel_nt640, void* ws, size_t ws_bytes, void* stream);
/* Upsample_Mel backward: d_up [N, keep, 80] -> d_kernel [1024, out, in], d_bias [80] */
size_t mstts_upsample_mel_bwd_workspace_bytes(int N, int Tm);
int mstts_upsample_mel_bwd(const float* mel, const float* d_up, int N, int Tm, int keep, float* d_kernel, float* d_bias, void* ws,
                           size_t ws_bytes, void* stream);

/* Upsample_Mel (Modules.py:198-208): ConvTranspose1d 80->80, k=1024, stride 256, VALID; kernel [1024, out, in];
 * writes the first `keep` of the (Tm-1)*256+1024 output frames: out [N, keep, 80]. */
size_t mstts_upsample_mel_workspace_bytes(int N, int Tm);
int mstts_upsample_mel(const float* mel, const float* kernel, const float* bias, int N, int Tm, int keep, float* out,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Zoneout LSTM over a whole sequence (H = 256): the recurrence of tf.nn.dynamic_rnn / stack_bidirectional_dynamic_rnn over
 * ZoneoutLSTMCell.call (ZoneoutLSTMCell.py:188-271) as used by the encoder BiLSTM (Modules.py:49-73) and the speaker-embedding
 * stack (Speaker_Embedding/Modules.py:12-37).  xk = x Kx + bias for all steps is the caller's GEMM; kh = the h rows of the cell
 * kernel [H,4H], gate order i j f o.  masks [T,2,B,H] u8 (c, h; indexed by loop step) or NULL (inference: no mask, the (1-rate)
 * factor stays).  reverse = 1 walks every row from its last valid frame down.  Outputs beyond lengths[b] are zero.
 * fwd saves acts [B,T,4H], c_prev, h_prev [B,T,H] when acts != NULL (pre-zeroed by the caller);
 * bwd returns dxk [B,T,4H] = gradient w.r.t. the gate pre-activations given dout [B,T,H] (w.r.t. the cell output m).
 * ---------------------------------------------------------------------------------------------- */
int mstts_zlstm_fwd(const float* xk, const float* kh, const int32_t* lengths, const uint8_t* masks, const float* x_res, int B, int T,
                    int H, int reverse, float keep, float* out, float* acts, float* c_prev, float* h_prev, void* stream);
int mstts_zlstm_bwd(const float* dout, const float* kh, const int32_t* lengths, const uint8_t* masks, const float* acts,
                    const float* c_prev, int B, int T, int H, int reverse, float keep, float* dxk, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tf.layers.conv1d(padding='same', stride 1) and its gradients as bf16x3 tensor-core GEMMs: encoder and postnet convolutions
 * (Modules.py:25-47,121-143).  x [B,T,Cin], kernel [k,Cin,Cout] (TF layout), y [B,T,Cout]; odd k <= 15, channels % 8 == 0.
 * bwd: dx (may be NULL) and dkernel are overwritten; the bias gradient is a column sum the caller owns.
 * ---------------------------------------------------------------------------------------------- */
size_t mstts_conv1d_workspace_bytes(int B, int T, int Cin, int Cout, int k);
int mstts_conv1d_fwd(const float* x, const float* kernel, const float* bias, int B, int T, int Cin, int Cout, int k, float* y, void* ws,
                     size_t ws_bytes, void* stream);
int mstts_conv1d_bwd(const float* x, const float* kernel, const float* dy, int B, int T, int Cin, int Cout, int k, float* dx,
                     float* dkernel, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * activation (0 relu / 1 tanh) -> tf.layers.batch_normalization -> tf.layers.dropout of the encoder / postnet conv stacks, fused
 * (Modules.py:29-45,125-141).  x, y [R, C] with R = B*T rows (padding included in the batch statistics, as in the reference);
 * training: biased batch statistics, moving <- moving * momentum + batch * (1 - momentum) updated in place, mask [R,C] u8 with
 * y scaled by 1/keep; inference: moving statistics, no dropout.  stats [2,C] (mean, rstd) and a_saved [R,C] (activation
 * output) feed the reverse pass, which overwrites dx, dgamma, dbeta.
 * ---------------------------------------------------------------------------------------------- */
size_t mstts_act_bn_dropout_workspace_bytes(int C);
int mstts_act_bn_dropout_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                             const uint8_t* mask, long long R, int C, int act, int training, float keep, float momentum, float eps, float* y,
                             float* a_saved, float* stats, void* ws, size_t ws_bytes, void* stream);
int mstts_act_bn_dropout_bwd(const float* dy, const float* a_saved, const float* stats, const float* gamma, const uint8_t* mask, long long R,
                             int C, int act, float keep, float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Audio features.  Replaces Audio.melspectrogram / spectrogram / spectrogram_and_mel (Audio.py:19-48,62-96):
 * pre-emphasis 0.97, librosa.stft (centre, reflect padding, periodic Hann(win) centred in n_fft), magnitude, optional
 * spectral subtraction, slaney mel filter bank (librosa.filters.mel defaults), 20 log10(max(1e-5,.)), clip to
 * [-max_abs, max_abs] (max_abs > 0) or [0,1].  wav [B,S] -> mel_out [B, 1+S/hop, n_mels] and/or spec_out
 * [B, 1+S/hop, n_fft/2+1] (either may be NULL).  One fused kernel, the complex spectrogram stays on chip.
 * ---------------------------------------------------------------------------------------------- */
size_t mstts_stft_mel_workspace_bytes(int B, int S, int n_fft, int hop, int n_mels, int spectral_subtract);
int mstts_stft_mel(const float* wav, int B, int S, int n_fft, int hop, int win, int n_mels, int sample_rate, float max_abs,
                   int spectral_subtract, float* mel_out, float* spec_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hand-written tcgen05 GEMM (csrc/tc_gemm.cu): C[M,N] fp32 = A[M,K] . B[K,N] as bf16x3 (A_hi B_hi + A_lo B_hi + A_hi B_lo, the
 * hi/lo tiles staged once in shared memory and multiplied three times) over operand images pre-tiled in the tensor core's
 * shared-memory layout: A [M/128][K/64][hi|lo][128 x 64], B [N/256][K/64][hi|lo][256 x 64], element (r,k) of a tile at byte
 * (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2.  N % 256 == 0, K % 64 == 0.
 * scratch (may be NULL; 9.7 MB = 74*256*128*4 + 1024 bytes): enables the split-K tail -- when the last round of the persistent
 * grid holds r <= 74 tiles they run as 2r half-K work items and are combined through this buffer.
 * mstts_tc_gemm_test tiles fp32 row-major A [M,K] and Bt [N,K] into the workspace first -- the parity hook; its workspace =
 * tiled A (4 bytes/element) + tiled B + 9.7 MB scratch + 2 KB.
 * ---------------------------------------------------------------------------------------------- */
int mstts_tc_gemm_tiled(const void* A_tiled, const void* B_tiled, int M, int N, int K, float* C, int ldc, void* scratch, void* stream);
int mstts_tc_gemm_test(const float* A, const float* Bt, int M, int N, int K, float* C, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * General dense product on that kernel (what every product outside the persistent loops goes through: the hoisted decoder
 * products and weight gradients, Modules.py:239-255,309-321; the WaveGlow reverse pass, WaveGlow/WaveGlow.py:54-74; the mel
 * up-sampling contraction, WaveGlow/Modules.py:198-208; the encoder / postnet convolutions).  The library links no vendor GEMM.
 *   C_b[M,N] = op(A_b) op(B_b) + beta C_b,  b in [0, batch)     row-major fp32, any leading dimensions
 *   op(A) is M x K (A stored K x M when transA), op(B) is K x N (B stored N x K when transB); stride* = elements between
 *   consecutive batches (strideB = 0 shares B).
 * A pack kernel splits the operands into bf16 hi + lo tile images (zero padded), the product runs as bf16x3 with fp32
 * accumulation in TMEM.  precise = 0: one accumulation chain over K (the TMEM accumulator truncates: ~1e-5 of max at K ~ 10^3);
 * 1: chains of <= 512 K-elements summed in double; 2: 3-way operand split, six partial products, chains of <= 256
 * (fp32-SGEMM accuracy at 4x the tensor work: the weight gradients of the narrow layers), and few-tile / long-K shapes are split over K with a deterministic second-pass reduction.  Scratch for
 * the images comes from a stream-ordered pool inside the library (cudaMallocAsync on the caller's stream, retained between
 * calls); mstts_release_scratch() synchronises the device and returns it to the driver.
 * ---------------------------------------------------------------------------------------------- */
int mstts_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, long long strideA, const float* B, int ldb,
                   long long strideB, float* C, int ldc, long long strideC, float beta, int batch, int precise, void* stream);
int mstts_release_scratch(void);

#ifdef __cplusplus
}
#endif
#endif /* MSTTS_B200_H_ */
