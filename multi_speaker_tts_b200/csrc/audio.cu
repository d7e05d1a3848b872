// Audio.py feature extraction on the GPU: pre-emphasis -> STFT (librosa.stft semantics) -> magnitude -> mel filter bank ->
// dB -> normalisation, fused into ONE kernel so the complex spectrogram never touches HBM (algorithmic traffic = read
// the waveform once + write [frames, 80]).  Replaces Audio.melspectrogram / spectrogram / spectrogram_and_mel
// (Audio.py:19-48,62-96).
//
// Persistent CTAs (8 per SM) loop over frames with the window / twiddle / filter tables staged once in shared memory: gather the reflect-padded, pre-emphasised, Hann-windowed frame into shared memory, real FFT of size
// n_fft as a complex Stockham radix-2 FFT of size n_fft/2 plus the split post-pass, |.|, sparse triangular filters
// (a warp per filter over its non-zero bin range), 20 log10(max(1e-5, .)), clip.
#include "common.cuh"

constexpr int kMaxFft = 4096;

struct StftParams {
  const float* wav;   // [B, S]
  int B, S, n_fft, hop, frames, n_mels, log2h;
  const float* window;  // [n_fft] periodic Hann(win) centred in n_fft
  const float2* tw;     // [n_fft/2] exp(-2 pi i j / n_fft)
  const float* fb;      // [n_mels, n_fft/2+1] dense filter bank
  const int* fb_range;  // [n_mels, 2] first / one-past-last non-zero bin
  const float* fbc;     // compact filter weights: filter m occupies fbc[fb_off[m] .. fb_off[m] + hi - lo)
  const int* fb_off;    // [n_mels + 1]
  float max_abs;        // > 0: symmetric normalisation to [-max_abs, max_abs]; <= 0: [0, 1]
  float* mel_out;       // [B, frames, n_mels] or NULL
  float* spec_out;      // [B, frames, n_fft/2+1] or NULL (Audio.spectrogram: ref level 20 dB, [0,1])
  float* mag_out;       // [B, frames, n_fft/2+1]: raw magnitudes for the spectral-subtraction path, or NULL
  const float* mag_mean;  // [B, n_fft/2+1] time means (spectral subtraction, second pass) or NULL
  const float* mag_in;    // second pass: magnitudes written by the first
};

__device__ __forceinline__ float amp_to_db(float x) { return 20.f * log10f(fmaxf(1e-5f, x)); }

// mel filter bank + dB + normalisation of one frame whose magnitudes sit in shared memory; the sparse (triangular)
// filter weights and their bin ranges are staged in shared memory once per CTA (fbc_s / rng_s, null: read global)
__device__ __forceinline__ void finish_frame(const StftParams& P, const float* mag_s, size_t frame_index, const float* fbc_s,
                                             const int* rng_s) {
  const int nb = P.n_fft / 2 + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (P.mel_out) {
    for (int m = warp; m < P.n_mels; m += nw) {
      int lo, hi, off;
      if (rng_s) {
        lo = rng_s[3 * m];
        hi = rng_s[3 * m + 1];
        off = rng_s[3 * m + 2];
      } else {
        lo = P.fb_range[2 * m];
        hi = P.fb_range[2 * m + 1];
        off = P.fb_off[m];
      }
      const float* wts = (fbc_s ? fbc_s : P.fbc) + off - lo;
      float s = 0.f;
      for (int k = lo + lane; k < hi; k += 32) s = fmaf(wts[k], mag_s[k], s);
      s = warp_sum(s);
      if (lane == 0) {
        const float db = amp_to_db(s);
        float v;
        if (P.max_abs > 0.f)
          v = fminf(fmaxf((2.f * P.max_abs) * ((db + 100.f) / 100.f) - P.max_abs, -P.max_abs), P.max_abs);
        else
          v = fminf(fmaxf((db + 100.f) / 100.f, 0.f), 1.f);
        P.mel_out[frame_index * P.n_mels + m] = v;
      }
    }
  }
  if (P.spec_out) {
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
      P.spec_out[frame_index * nb + k] = fminf(fmaxf((amp_to_db(mag_s[k]) - 20.f + 100.f) / 100.f, 0.f), 1.f);
  }
}

__global__ void __launch_bounds__(256) stft_mel_kernel(const StftParams P, size_t nframes) {
  extern __shared__ __align__(16) float smem_f[];
  const int N = P.n_fft, H = N / 2;
  float2* za = reinterpret_cast<float2*>(smem_f);  // [H]
  float2* zb = za + H;                             // [H]
  float2* tw_s = zb + H;                           // [H]
  float* win_s = reinterpret_cast<float*>(tw_s + H);  // [N]
  float* mag_s = win_s + N;                           // [H+1] (+pad)
  float* fbc_s = mag_s + H + 4;                       // [3*(H+1)] compact filter weights
  int* rng_s = reinterpret_cast<int*>(fbc_s + 3 * (H + 1));  // [n_mels][3]
  const int tid = threadIdx.x;
  // ---- per-CTA tables (a CTA then loops over many frames) ----
  for (int i = tid; i < H; i += blockDim.x) tw_s[i] = P.tw[i];
  for (int i = tid; i < N; i += blockDim.x) win_s[i] = P.window[i];
  if (P.mel_out) {
    const int nnz = P.fb_off[P.n_mels];
    for (int i = tid; i < nnz; i += blockDim.x) fbc_s[i] = P.fbc[i];
    for (int m = tid; m < P.n_mels; m += blockDim.x) {
      rng_s[3 * m] = P.fb_range[2 * m];
      rng_s[3 * m + 1] = P.fb_range[2 * m + 1];
      rng_s[3 * m + 2] = P.fb_off[m];
    }
  }
  __syncthreads();
  for (size_t fi = blockIdx.x; fi < nframes; fi += gridDim.x) {
    const int b = (int)(fi / P.frames), fr = (int)(fi % P.frames);
    const float* x = P.wav + (size_t)b * P.S;
    // ---- frame gather: reflect padding of the PRE-EMPHASISED signal (librosa pads after Audio.preemphasis ran) ----
#pragma unroll 4
    for (int i = tid; i < N; i += blockDim.x) {
      int j = fr * P.hop + i - H;
      if (j < 0) j = -j;
      if (j >= P.S) j = 2 * (P.S - 1) - j;
      j = min(max(j, 0), P.S - 1);
      const float x0 = x[j], x1 = x[max(j - 1, 0)];
      const float y = (j > 0) ? x0 - 0.97f * x1 : x0;
      reinterpret_cast<float*>(za)[i] = y * win_s[i];  // z[k] = (x[2k], x[2k+1])
    }
    __syncthreads();
    // ---- complex FFT of size H (Stockham autosort, radix 2): stage s has stride 1<<s ----
    float2* src = za;
    float2* dst = zb;
    for (int s = 0; s < P.log2h; ++s) {
      const int l = H >> (s + 1);
      const int m = 1 << s;
      for (int i = tid; i < H / 2; i += blockDim.x) {
        const int j = i >> s, k = i & (m - 1);  // i = j*m + k, j < l
        const float2 c0 = src[k + j * m];
        const float2 c1 = src[k + j * m + l * m];
        const float2 w = tw_s[j * 2 * m];  // exp(-2 pi i j / (H >> s)) = tw[j * N / (H >> s)]
        const float2 sum = make_float2(c0.x + c1.x, c0.y + c1.y);
        const float2 dif = make_float2(c0.x - c1.x, c0.y - c1.y);
        dst[k + 2 * j * m] = sum;
        dst[k + (2 * j + 1) * m] = make_float2(dif.x * w.x - dif.y * w.y, dif.x * w.y + dif.y * w.x);
      }
      __syncthreads();
      float2* t = src;
      src = dst;
      dst = t;
    }
    // ---- split post-pass: X[k] = (Z[k] + conj(Z[H-k]))/2 - i e^{-2 pi i k/N} (Z[k] - conj(Z[H-k]))/2 ----
    for (int k = tid; k <= H; k += blockDim.x) {
      const float2 zk = src[k % H];
      const float2 zc = src[(H - k) % H];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
      const float2 w = (k < H) ? tw_s[k] : make_float2(-1.f, 0.f);
      const float2 wo = make_float2(w.x * o.x - w.y * o.y, w.x * o.y + w.y * o.x);
      const float re = e.x + wo.y, im = e.y - wo.x;
      mag_s[k] = sqrtf(re * re + im * im);
    }
    __syncthreads();
    if (P.mag_out) {
      for (int k = tid; k <= H; k += blockDim.x) P.mag_out[fi * (H + 1) + k] = mag_s[k];
    } else {
      finish_frame(P, mag_s, fi, fbc_s, rng_s);
    }
    __syncthreads();  // mag_s / za are rewritten by the next frame
  }
}

// spectral subtraction, second pass: M' = max(M - mean_t(M)/10, 0) (Audio.py:45-46), then the usual tail
__global__ void __launch_bounds__(256) subtract_finish_kernel(const StftParams P) {
  extern __shared__ __align__(16) float smem_f[];
  float* mag_s = smem_f;
  const int nb = P.n_fft / 2 + 1;
  const size_t fi = blockIdx.x;
  const int b = (int)(fi / P.frames);
  for (int k = threadIdx.x; k < nb; k += blockDim.x)
    mag_s[k] = fmaxf(P.mag_in[fi * nb + k] - P.mag_mean[(size_t)b * nb + k] / 10.f, 0.f);
  __syncthreads();
  finish_frame(P, mag_s, fi, nullptr, nullptr);
}

// mean over frames of the magnitudes, one thread per (utterance, bin), fixed order
__global__ void mag_mean_kernel(const float* __restrict__ mag, int B, int frames, int nb, float* __restrict__ mean) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nb) return;
  const int b = i / nb, k = i % nb;
  float s = 0.f;
  for (int f = 0; f < frames; ++f) s += mag[((size_t)b * frames + f) * nb + k];
  mean[i] = s / (float)frames;
}

// tables: periodic Hann(win) centred in n_fft, twiddles, slaney mel filter bank (librosa.filters.mel defaults)
// The tables depend only on (n_fft, win, n_mels, sr): a tag in the workspace lets repeated calls with the same
// workspace skip the rebuild (the reference rebuilds the mel basis on every call, Audio.py:78).
__global__ void stft_tag_kernel(int4* tag, int n_fft, int win, int n_mels, int sr) { *tag = make_int4(n_fft, win, n_mels, sr); }

__global__ void stft_tables_kernel(const int4* tag, int n_fft, int win, int n_mels, int sr, float* window, float2* tw, float* fb, int* fb_range) {
  {
    const int4 t = *tag;
    if (t.x == n_fft && t.y == win && t.z == n_mels && t.w == sr) return;
  }
  const int nb = n_fft / 2 + 1;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int off = (n_fft - win) / 2;
  for (int i = tid; i < n_fft; i += nth) {
    const int j = i - off;
    window[i] = (j >= 0 && j < win) ? (float)(0.5 - 0.5 * cospi(2.0 * (double)j / (double)win)) : 0.f;
  }
  for (int i = tid; i < n_fft / 2; i += nth) {
    double s, c;
    sincospi(-2.0 * (double)i / (double)n_fft, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  // slaney mel scale (htk=False): linear below 1 kHz (200/3 Hz per mel), log above (step log(6.4)/27)
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  const double f_nyq = 0.5 * (double)sr;
  const double mel_max = f_nyq >= min_log_hz ? min_log_mel + log(f_nyq / min_log_hz) / logstep : f_nyq / f_sp;
  auto mel_to_hz = [&](double m) { return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m; };
  for (int m = tid; m < n_mels; m += nth) {
    const double f0 = mel_to_hz(mel_max * (double)m / (double)(n_mels + 1));
    const double f1 = mel_to_hz(mel_max * (double)(m + 1) / (double)(n_mels + 1));
    const double f2 = mel_to_hz(mel_max * (double)(m + 2) / (double)(n_mels + 1));
    const double enorm = 2.0 / (f2 - f0);
    int lo = nb, hi = 0;
    for (int k = 0; k < nb; ++k) {
      const double fk = f_nyq * (double)k / (double)(nb - 1);
      const double lower = (fk - f0) / (f1 - f0), upper = (f2 - fk) / (f2 - f1);
      const double wv = fmax(0.0, fmin(lower, upper)) * enorm;
      fb[(size_t)m * nb + k] = (float)wv;
      if (wv > 0.0) {
        lo = min(lo, k);
        hi = max(hi, k + 1);
      }
    }
    if (hi <= lo) lo = hi = 0;
    fb_range[2 * m] = lo;
    fb_range[2 * m + 1] = hi;
  }
}

// compact (non-zero) filter weights: off[m] = prefix sum of the range lengths, fbc[off[m] + k - lo] = fb[m][k]
__global__ void stft_compact_kernel(const int4* tag, int n_fft, int win, int n_mels, int sr, const float* fb, const int* fb_range, int* fb_off,
                                    float* fbc) {
  {
    const int4 t = *tag;
    if (t.x == n_fft && t.y == win && t.z == n_mels && t.w == sr) return;
  }
  __shared__ int off_s[257];
  const int nb = n_fft / 2 + 1;
  if (threadIdx.x == 0) {
    int o = 0;
    for (int m = 0; m < n_mels; ++m) {
      off_s[m] = o;
      o += fb_range[2 * m + 1] - fb_range[2 * m];
    }
    off_s[n_mels] = o;
  }
  __syncthreads();
  for (int m = threadIdx.x; m <= n_mels; m += blockDim.x) fb_off[m] = off_s[m];
  for (int m = 0; m < n_mels; ++m) {
    const int lo = fb_range[2 * m], hi = fb_range[2 * m + 1];
    for (int k = lo + threadIdx.x; k < hi; k += blockDim.x) fbc[off_s[m] + k - lo] = fb[(size_t)m * nb + k];
  }
}

static size_t stft_ws_layout(int B, int frames, int n_fft, int n_mels, int subtract, size_t* o_tag, size_t* o_win, size_t* o_tw, size_t* o_fb,
                             size_t* o_rng, size_t* o_mag, size_t* o_mean, size_t* o_fbc, size_t* o_off) {
  const size_t nb = n_fft / 2 + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  *o_tag = take(16);
  *o_win = take((size_t)n_fft * 4);
  *o_tw = take((size_t)n_fft / 2 * 8);
  *o_fb = take((size_t)n_mels * nb * 4);
  *o_rng = take((size_t)n_mels * 2 * 4);
  *o_fbc = take((size_t)3 * nb * 4);
  *o_off = take((size_t)(n_mels + 1) * 4);
  *o_mag = take(subtract ? (size_t)B * frames * nb * 4 : 0);
  *o_mean = take(subtract ? (size_t)B * nb * 4 : 0);
  return off;
}

extern "C" size_t mstts_stft_mel_workspace_bytes(int B, int S, int n_fft, int hop, int n_mels, int spectral_subtract) {
  if (B <= 0 || S <= 0 || n_fft <= 0 || hop <= 0) return 0;
  size_t t, a, b2, c, d, e, f, g2, h2;
  return stft_ws_layout(B, 1 + S / hop, n_fft, n_mels > 0 ? n_mels : 1, spectral_subtract, &t, &a, &b2, &c, &d, &e, &f, &g2, &h2);
}

extern "C" int mstts_stft_mel(const float* wav, int B, int S, int n_fft, int hop, int win, int n_mels, int sample_rate,
                              float max_abs, int spectral_subtract, float* mel_out, float* spec_out, void* ws_, size_t ws_bytes,
                              void* stream) {
  MSTTS_REQUIRE(wav && (mel_out || spec_out) && ws_, MSTTS_E_INVALID, "stft_mel: null pointer");
  MSTTS_REQUIRE(n_fft >= 64 && n_fft <= kMaxFft && (n_fft & (n_fft - 1)) == 0, MSTTS_E_UNSUPPORTED,
                "stft_mel: n_fft=%d must be a power of two in [64,%d]", n_fft, kMaxFft);
  MSTTS_REQUIRE(win >= 1 && win <= n_fft && hop >= 1, MSTTS_E_INVALID, "stft_mel: win=%d hop=%d n_fft=%d", win, hop, n_fft);
  MSTTS_REQUIRE(S > n_fft / 2, MSTTS_E_INVALID, "stft_mel: signal shorter than n_fft/2 cannot be reflect-padded (S=%d)", S);
  MSTTS_REQUIRE(!mel_out || (n_mels >= 1 && n_mels <= 256 && sample_rate > 0), MSTTS_E_INVALID, "stft_mel: n_mels=%d sr=%d", n_mels,
                sample_rate);
  cudaStream_t s = (cudaStream_t)stream;
  const int frames = 1 + S / hop;
  size_t o_tag, o_win, o_tw, o_fb, o_rng, o_mag, o_mean, o_fbc, o_off;
  const size_t need = stft_ws_layout(B, frames, n_fft, n_mels > 0 ? n_mels : 1, spectral_subtract, &o_tag, &o_win, &o_tw, &o_fb, &o_rng, &o_mag, &o_mean, &o_fbc, &o_off);
  MSTTS_REQUIRE(ws_bytes >= need, MSTTS_E_WORKSPACE, "stft_mel: workspace %zu < %zu", ws_bytes, need);
  char* ws = (char*)ws_;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.wav = wav; P.B = B; P.S = S; P.n_fft = n_fft; P.hop = hop; P.frames = frames; P.n_mels = n_mels;
  int lg = 0;
  while ((1 << lg) < n_fft / 2) ++lg;
  P.log2h = lg;
  P.window = (float*)(ws + o_win); P.tw = (float2*)(ws + o_tw); P.fb = (float*)(ws + o_fb); P.fb_range = (int*)(ws + o_rng);
  P.fbc = (float*)(ws + o_fbc); P.fb_off = (int*)(ws + o_off);
  P.max_abs = max_abs; P.mel_out = mel_out; P.spec_out = spec_out;
  {
    const int nm = n_mels > 0 ? n_mels : 1, sr = sample_rate > 0 ? sample_rate : 1;
    stft_tables_kernel<<<16, 128, 0, s>>>((const int4*)(ws + o_tag), n_fft, win, nm, sr, (float*)(ws + o_win), (float2*)(ws + o_tw),
                                          (float*)(ws + o_fb), (int*)(ws + o_rng));
    stft_compact_kernel<<<1, 256, 0, s>>>((const int4*)(ws + o_tag), n_fft, win, nm, sr, (float*)(ws + o_fb), (int*)(ws + o_rng), (int*)(ws + o_off),
                                          (float*)(ws + o_fbc));
    stft_tag_kernel<<<1, 1, 0, s>>>((int4*)(ws + o_tag), n_fft, win, nm, sr);
  }
  const size_t smem = (size_t)(3 * (n_fft / 2)) * sizeof(float2) + (size_t)n_fft * 4 + (size_t)(n_fft / 2 + 4) * 4 +
                      (size_t)3 * (n_fft / 2 + 1) * 4 + (size_t)3 * 256 * 4;
  MSTTS_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t nframes = (size_t)B * frames;
  MSTTS_REQUIRE(nframes < (1ull << 31), MSTTS_E_INVALID, "stft_mel: too many frames");
  const unsigned grid = (unsigned)(nframes < 148u * 8u ? nframes : 148u * 8u);  // persistent: 8 CTAs per SM
  if (!spectral_subtract) {
    stft_mel_kernel<<<grid, 256, smem, s>>>(P, nframes);
  } else {
    StftParams Q = P;
    Q.mag_out = (float*)(ws + o_mag);
    stft_mel_kernel<<<grid, 256, smem, s>>>(Q, nframes);
    const int nb = n_fft / 2 + 1;
    mag_mean_kernel<<<(B * nb + 127) / 128, 128, 0, s>>>((float*)(ws + o_mag), B, frames, nb, (float*)(ws + o_mean));
    P.mag_in = (float*)(ws + o_mag);
    P.mag_mean = (float*)(ws + o_mean);
    subtract_finish_kernel<<<(unsigned)nframes, 256, (size_t)nb * sizeof(float), s>>>(P);
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
