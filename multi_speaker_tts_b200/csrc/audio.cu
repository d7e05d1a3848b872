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


// =====================================================================================================================
// Team kernel (the fast path): a "team" of H/8 threads owns one frame; a CTA of 256 threads runs 256/(H/8) consecutive
// frames of one utterance at a time and loops persistently over such groups.
//   * the raw segment the group needs (n_fft + (teams-1)*hop samples) is staged ONCE, coalesced, reflect-indexed and
//     pre-emphasised, in shared memory: the 4x overlap of neighbouring frames never goes back to L2;
//   * the n_fft/2-point complex FFT is a Stockham autosort with radix-8 butterflies held in registers (one radix-2/4 stage
//     first when log2 is not a multiple of 3): 3 shared-memory round trips for n_fft = 1024 instead of 9, synchronised
//     with named barriers per team, never the whole CTA;
//   * shared arrays are XOR-swizzled so both the stride-8 scatter of the first stage and every consecutive access are conflict free;
//   * mel filters: one thread per filter over its non-zero bins (the triangles are 3..45 bins wide), then dB, clip, store.
// =====================================================================================================================
// XOR swizzle of the complex work arrays: slot = i ^ ((i >> 4) & 15).  Consecutive accesses stay a permutation inside an
// aligned group of 16 (ideal wavefronts, no padding), and the stride-8 scatter of the first radix-8 stage (i = 8j + t) lands
// in 16 distinct 8-byte banks per half-warp.  (Padding i + i/8 fixed the scatter but doubled the wavefronts of every
// consecutive access: 32 float2 then span 36 slots.)
__device__ __forceinline__ int padi(int i) { return i ^ ((i >> 4) & 15); }
// sqrt.approx.f32: max relative error 2^-23 (PTX ISA) -- the magnitudes feed a dB scale compared at 1e-3; the IEEE sqrtf
// sequence was 9 % of the kernel's instructions
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  const float2 t = a;
  a = make_float2(t.x + b.x, t.y + b.y);
  b = make_float2(t.x - b.x, t.y - b.y);
}
// forward DFTs (kernel e^{-2 pi i nk/R}), result in natural order in o[]
template <int R>
__device__ __forceinline__ void dft_regs(float2 (&v)[R], float2 (&o)[R]);
template <>
__device__ __forceinline__ void dft_regs<2>(float2 (&v)[2], float2 (&o)[2]) {
  bfly(v[0], v[1]);
  o[0] = v[0];
  o[1] = v[1];
}
__device__ __forceinline__ void dft4_inplace(float2& a, float2& b, float2& c, float2& d) {  // -> X0, X1, X2, X3 in a, b, c, d
  bfly(a, c);
  bfly(b, d);
  d = make_float2(d.y, -d.x);  // * -i
  bfly(a, b);                  // a = X0, b = X2
  bfly(c, d);                  // c = X1, d = X3
  const float2 t = b;
  b = c;
  c = t;
}
template <>
__device__ __forceinline__ void dft_regs<4>(float2 (&v)[4], float2 (&o)[4]) {
  dft4_inplace(v[0], v[1], v[2], v[3]);
#pragma unroll
  for (int r = 0; r < 4; ++r) o[r] = v[r];
}
template <>
__device__ __forceinline__ void dft_regs<8>(float2 (&v)[8], float2 (&o)[8]) {
  constexpr float kS = 0.70710678118654752f;
  bfly(v[0], v[4]);
  bfly(v[1], v[5]);
  bfly(v[2], v[6]);
  bfly(v[3], v[7]);
  v[5] = make_float2((v[5].x + v[5].y) * kS, (v[5].y - v[5].x) * kS);    // * w8^1
  v[6] = make_float2(v[6].y, -v[6].x);                                   // * w8^2 = -i
  v[7] = make_float2((v[7].y - v[7].x) * kS, (-v[7].x - v[7].y) * kS);   // * w8^3
  dft4_inplace(v[0], v[1], v[2], v[3]);  // X0 X2 X4 X6
  dft4_inplace(v[4], v[5], v[6], v[7]);  // X1 X3 X5 X7
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    o[2 * r] = v[r];
    o[2 * r + 1] = v[4 + r];
  }
}

template <int TEAM>
__device__ __forceinline__ void team_sync(int team) {
  if (TEAM == 32)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TEAM) : "memory");
}

// padi is linear over GF(2): padi(a + b) = padi(a) ^ padi(b) whenever a and b occupy disjoint bit fields, which every index of
// the power-of-two schedule does (thread part | unrolled-loop part).  The thread parts are swizzled once per stage, the loop
// parts are compile-time constants: one XOR per access instead of shift + and + xor + add.
__host__ __device__ constexpr int padc(int i) { return i ^ ((i >> 4) & 15); }
template <int V>
__host__ __device__ constexpr int ilog2c() { return V <= 1 ? 0 : 1 + ilog2c<V / 2>(); }

// one Stockham stage of radix R over H points: butterfly j handles inputs j + r H/R, outputs (j/Ns) Ns R + j%Ns + r Ns
// (NS = compile-time Ns; NS <= TEAM, so j%Ns and j/Ns split into a thread part and an iteration part)
template <int R, int H, int TEAM, int NS>
__device__ __forceinline__ void fft_stage(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ twst,
                                          const int ltid) {
  static_assert(NS <= TEAM, "j % Ns must be a function of the thread index alone");
  constexpr int LR = ilog2c<R>(), LNS = ilog2c<NS>();
  const int k = ltid & (NS - 1);
  const int pl = padi(ltid);                                        // loads: index = ltid + (it TEAM + r H/R)
  const int ps = padi(((ltid >> LNS) << (LNS + LR)) | k);           // stores: index = (ltid/NS) NS R + k + (it TEAM R + r NS)
#pragma unroll
  for (int it = 0; it < (H / R + TEAM - 1) / TEAM; ++it) {
    if (H / R < TEAM && ltid >= H / R) break;
    float2 v[R], o[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[pl ^ padc(it * TEAM + r * (H / R))];
    if (NS > 1) {
      // per-stage table twst[(r-1) Ns + k] = exp(-2 pi i k r / (Ns R)): lanes of a warp have consecutive k, so the reads are
      // conflict free (indexing the n_fft-th roots table directly is an 8-way conflict: stride 128 B between lanes)
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmulf(v[r], twst[(r - 1) * NS + k]);
    }
    dft_regs<R>(v, o);
#pragma unroll
    for (int r = 0; r < R; ++r) dst[ps ^ padc(it * TEAM * R + r * NS)] = o[r];
  }
}

// the FIRST stage (Ns = 1, no twiddles) reads the staged raw segment directly and applies the window from registers: the
// windowed frame is never written to shared memory.  z[n] = (x[2n] w[2n], x[2n+1] w[2n+1]); wreg holds this thread's H/TEAM
// window pairs in the order the loop consumes them (the same for every frame).
template <int R, int H, int TEAM>
__device__ __forceinline__ void fft_stage_first(const float* __restrict__ rw, const bool even, const float2 (&wreg)[H / TEAM],
                                                float2* __restrict__ dst, const int ltid) {
  constexpr int LR = ilog2c<R>();
  const int ps = padi(ltid << LR);  // stores: index = ltid R + (it TEAM R + r)
  int idx = 0;
#pragma unroll
  for (int it = 0; it < (H / R + TEAM - 1) / TEAM; ++it) {
    if (H / R < TEAM && ltid >= H / R) break;
    float2 v[R], o[R];
#pragma unroll
    for (int r = 0; r < R; ++r, ++idx) {
      const int n = ltid + it * TEAM + r * (H / R);
      const float2 x = even ? *reinterpret_cast<const float2*>(rw + 2 * n) : make_float2(rw[2 * n], rw[2 * n + 1]);
      v[r] = make_float2(x.x * wreg[idx].x, x.y * wreg[idx].y);
    }
    dft_regs<R>(v, o);
#pragma unroll
    for (int r = 0; r < R; ++r) dst[ps ^ padc(it * TEAM * R + r)] = o[r];
  }
}

// the radix-8 stages after the first one (Ns = NS, 8 NS, ... < H), ping-ponging src / dst; TWOFF = offset of the stage's twiddles
template <int H, int TEAM, int NS, int TWOFF>
__device__ __forceinline__ void fft_rest(float2*& src, float2*& dst, const float2* __restrict__ twst, const int ltid, const int team) {
  if constexpr (NS < H) {
    fft_stage<8, H, TEAM, NS>(src, dst, twst + TWOFF, ltid);
    float2* t = src; src = dst; dst = t;
    team_sync<TEAM>(team);
    fft_rest<H, TEAM, NS * 8, TWOFF + 7 * NS>(src, dst, twst, ltid, team);
  }
}

template <int LOG2H>
struct StftTeamCfg {
  static constexpr int H = 1 << LOG2H, N = 2 * H;
  static constexpr int TEAM = (H / 8 < 32) ? 32 : (H / 8 > 256 ? 256 : H / 8);
  static constexpr int R1 = (LOG2H % 3 == 1) ? 2 : (LOG2H % 3 == 2) ? 4 : 8;  // radix of the first stage
  static constexpr int kTabFloats = (2 * H + 3 * (H + 1) + 3 * 256 + 1) & ~1;  // tw_s, fbc_s, rng_s (complex arrays 8-byte aligned)
};

template <int LOG2H, int NT>
__global__ void __launch_bounds__(NT) stft_team_kernel(const StftParams P, int groups_per_utt, int ngroups) {
  using Cfg = StftTeamCfg<LOG2H>;
  constexpr int H = Cfg::H, N = Cfg::N, TEAM = Cfg::TEAM, R1 = Cfg::R1;
  constexpr int TPC = NT / TEAM;             // frames per CTA iteration
  constexpr int ZP = H;                      // complex array length (XOR swizzle, no padding)
  extern __shared__ __align__(16) float smem_f[];
  float2* tw_s = reinterpret_cast<float2*>(smem_f);              // [H]
  float* fbc_s = reinterpret_cast<float*>(tw_s + H);             // [3*(H+1)] compact filter weights
  int* rng_s = reinterpret_cast<int*>(fbc_s + 3 * (H + 1));      // [256][3]
  float2* z_all = reinterpret_cast<float2*>(smem_f + Cfg::kTabFloats);
  float* mag_all = reinterpret_cast<float*>(z_all + (size_t)TPC * 2 * ZP);  // [TPC][H+4]
  float2* twst_s = reinterpret_cast<float2*>(mag_all + TPC * (H + 4));  // [H] per-stage twiddle tables (radix-8 stages)
  float* raw_s = reinterpret_cast<float*>(twst_s + H);           // [N + (TPC-1)*hop]
  const int tid = threadIdx.x, team = tid / TEAM, ltid = tid % TEAM;
  float2* za = z_all + (size_t)team * 2 * ZP;
  float2* zb = za + ZP;
  float* mag_s = mag_all + team * (H + 4);

  for (int i = tid; i < H; i += NT) tw_s[i] = P.tw[i];
  if (P.mel_out) {
    const int nnz = P.fb_off[P.n_mels];
    for (int i = tid; i < nnz; i += NT) fbc_s[i] = P.fbc[i];
    for (int m = tid; m < P.n_mels; m += NT) {
      rng_s[3 * m] = P.fb_range[2 * m];
      rng_s[3 * m + 1] = P.fb_range[2 * m + 1];
      rng_s[3 * m + 2] = P.fb_off[m];
    }
  }
  {
    int Ns = (R1 == 8) ? 1 : R1, off = 0;
#pragma unroll
    for (int it = 0; it < LOG2H / 3; ++it) {
      if (Ns > 1) {
        for (int idx = tid; idx < 7 * Ns; idx += NT) {
          const int r = idx / Ns + 1, k = idx % Ns;
          const int i = r * k * (2 * H / (Ns * 8));  // P.tw[i] = exp(-2 pi i * i / (2H)), i < H; the other half is its negative
          float2 w = P.tw[i & (H - 1)];
          if (i & H) w = make_float2(-w.x, -w.y);
          twst_s[off + idx] = w;
        }
        off += 7 * Ns;
      }
      Ns *= 8;
    }
  }
  // this thread's window pairs, in the order fft_stage_first consumes them
  float2 wreg[H / TEAM];
  {
    int idx = 0;
#pragma unroll
    for (int j = ltid; j < H / R1; j += TEAM) {
#pragma unroll
      for (int r = 0; r < R1; ++r, ++idx) {
        const int n = j + r * (H / R1);
        wreg[idx] = make_float2(P.window[2 * n], P.window[2 * n + 1]);
      }
    }
  }
  __syncthreads();
  const int rawlen = N + (TPC - 1) * P.hop;
  const float inv_gpu = 1.f / (float)groups_per_utt;
  const int MG = TEAM / 4;  // mel filters in flight per team: 4 lanes each
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    int b = __float2int_rz(((float)g + 0.5f) * inv_gpu);  // floor(g / groups_per_utt) without the integer-division sequence
    if (b * groups_per_utt > g) --b;                       // (the reciprocal product can be off by one at most)
    if ((b + 1) * groups_per_utt <= g) ++b;
    const int fr0 = (g - b * groups_per_utt) * TPC;
    const float* x = P.wav + (size_t)b * P.S;
    // ---- stage the pre-emphasised, reflect-padded segment (librosa pads AFTER Audio.preemphasis ran) ----
    const int j0 = fr0 * P.hop - H;
    if (j0 >= 1 && j0 + rawlen <= P.S) {  // interior group: no reflection, x[j-1] always exists
      const float* xb = x + j0;
      for (int i = tid; i < rawlen; i += NT) raw_s[i] = __ldg(xb + i) - 0.97f * __ldg(xb + i - 1);
    } else {
      for (int i = tid; i < rawlen; i += NT) {
        int j = j0 + i;
        if (j < 0) j = -j;
        if (j >= P.S) j = 2 * (P.S - 1) - j;
        j = min(max(j, 0), P.S - 1);
        const float x0 = __ldg(x + j), x1 = __ldg(x + max(j - 1, 0));
        raw_s[i] = (j > 0) ? x0 - 0.97f * x1 : x0;
      }
    }
    __syncthreads();
    const int fr = fr0 + team;
    const bool live = fr < P.frames;
    float2* src = za;
    float2* dst = zb;
    if (live) {  // first stage straight from the staged segment (window in registers)
      fft_stage_first<R1, H, TEAM>(raw_s + team * P.hop, ((team * P.hop) & 1) == 0, wreg, zb, ltid);
    }
    __syncthreads();  // raw_s may be overwritten by the next group from here on; teams run independently below
    if (!live) continue;
    // ---- remaining stages of the complex FFT of size H (the first one wrote zb) ----
    src = zb;
    dst = za;
    fft_rest<H, TEAM, R1, 0>(src, dst, twst_s, ltid, team);
    // ---- split post-pass: X[k] = (Z[k] + conj(Z[H-k]))/2 - i e^{-2 pi i k/N} (Z[k] - conj(Z[H-k]))/2, magnitude.  The pair
    //      (k, H-k) shares its two loads and its twiddle (w_{H-k} = -conj(w_k)): X[H-k] = (e.x - wo.y, -e.y - wo.x) ----
    for (int k = ltid; k <= H / 2; k += TEAM) {
      const float2 zk = src[padi(k)];
      const float2 zc = src[padi((H - k) & (H - 1))];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
      const float2 wo = cmulf(tw_s[k], o);
      const float re = e.x + wo.y, im = e.y - wo.x;
      const float re2 = e.x - wo.y, im2 = e.y + wo.x;
      mag_s[k] = sqrt_approx(re * re + im * im);
      mag_s[H - k] = sqrt_approx(re2 * re2 + im2 * im2);
    }
    team_sync<TEAM>(team);
    const size_t fi = (size_t)b * P.frames + fr;
    if (P.mag_out) {
      for (int k = ltid; k <= H; k += TEAM) P.mag_out[fi * (H + 1) + k] = mag_s[k];
    } else {
      if (P.mel_out) {
        // 4 lanes per triangular filter (3 .. 45 bins wide): neighbouring filters have similar widths, so a round costs its
        // widest filter / 4 iterations (thread-per-filter: the widest filter of each warp pass, with every lane on its own bank
        // pattern); the sums go through shared memory so that dB + clip + store run once per filter, not once per round
        float* msum = reinterpret_cast<float*>(dst);  // the work array the last FFT stage did not write
        const int gl = ltid & 3, grp = ltid >> 2;
        for (int m0 = 0; m0 < P.n_mels; m0 += MG) {
          const int m = m0 + grp;
          float sum = 0.f;
          if (m < P.n_mels) {
            const int lo = rng_s[3 * m], hi = rng_s[3 * m + 1];
            const float* wts = fbc_s + rng_s[3 * m + 2] - lo;
            // pointer pair + 4-deep unroll: 2 LDS + 1 FFMA per element instead of recomputing both addresses
            const float* wp = wts + lo + gl;
            const float* mp = mag_s + lo + gl;
            int n = hi - lo - gl;  // this lane owns offsets 0, 4, 8, ... < n
#pragma unroll 1
            for (; n > 12; n -= 16, wp += 16, mp += 16) {
              sum = fmaf(wp[0], mp[0], sum);
              sum = fmaf(wp[4], mp[4], sum);
              sum = fmaf(wp[8], mp[8], sum);
              sum = fmaf(wp[12], mp[12], sum);
            }
            if (n > 0) sum = fmaf(wp[0], mp[0], sum);
            if (n > 4) sum = fmaf(wp[4], mp[4], sum);
            if (n > 8) sum = fmaf(wp[8], mp[8], sum);
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          if (gl == 0 && m < P.n_mels) msum[m] = sum;
        }
        team_sync<TEAM>(team);
        for (int m = ltid; m < P.n_mels; m += TEAM) {
          const float db = amp_to_db(msum[m]);
          float v;
          if (P.max_abs > 0.f)
            v = fminf(fmaxf((2.f * P.max_abs) * ((db + 100.f) * 0.01f) - P.max_abs, -P.max_abs), P.max_abs);
          else
            v = fminf(fmaxf((db + 100.f) * 0.01f, 0.f), 1.f);
          P.mel_out[fi * P.n_mels + m] = v;
        }
      }
      if (P.spec_out) {
        for (int k = ltid; k <= H; k += TEAM)
          P.spec_out[fi * (H + 1) + k] = fminf(fmaxf((amp_to_db(mag_s[k]) - 20.f + 100.f) / 100.f, 0.f), 1.f);
      }
    }
    team_sync<TEAM>(team);  // mag_s / za are rewritten by this team's next frame
  }
}

template <int LOG2H, int NT>
static size_t stft_team_smem(int hop) {
  using Cfg = StftTeamCfg<LOG2H>;
  constexpr int H = Cfg::H, N = Cfg::N, TPC = NT / Cfg::TEAM, ZP = H;
  return (size_t)Cfg::kTabFloats * 4 + (size_t)TPC * 2 * ZP * 8 + (size_t)TPC * (H + 4) * 4 + (size_t)H * 8 +
         (size_t)(N + (TPC - 1) * hop) * 4 + 16;
}

// returns 1 if launched, 0 if this shape is left to the generic kernel, <0 on error
template <int LOG2H, int NT>
static int stft_team_launch(const StftParams& P, cudaStream_t s) {
  constexpr int TPC = NT / StftTeamCfg<LOG2H>::TEAM;
  const size_t smem = stft_team_smem<LOG2H, NT>(P.hop);
  int dev = 0, max_optin = 0;
  MSTTS_CUDA(cudaGetDevice(&dev));
  MSTTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem > (size_t)max_optin || P.n_mels > 256) return 0;
  MSTTS_CUDA(cudaFuncSetAttribute(stft_team_kernel<LOG2H, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int gpu = (P.frames + TPC - 1) / TPC;
  const long long ngroups = (long long)P.B * gpu;
  int per_sm = 1;
  MSTTS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_team_kernel<LOG2H, NT>, NT, smem));
  if (per_sm < 1) per_sm = 1;
  const long long cap = 148LL * per_sm;
  stft_team_kernel<LOG2H, NT><<<(unsigned)(ngroups < cap ? ngroups : cap), NT, smem, s>>>(P, gpu, (int)ngroups);
  MSTTS_CUDA(cudaGetLastError());
  return 1;
}

// threads per CTA of the team kernel: 512 (two resident CTAs = 16 frames in flight per SM at n_fft = 1024) unless
// MSTTS_STFT_NT=256 asks for the smaller CTA (three resident CTAs = 12 frames)
static int stft_nt() {
  static const int nt = [] {
    const char* e = getenv("MSTTS_STFT_NT");
    return (e && atoi(e) == 256) ? 256 : 512;
  }();
  return nt;
}
template <int LOG2H>
static int stft_team_launch_nt(const StftParams& P, cudaStream_t s) {
  if (stft_nt() == 512) {
    const int r = stft_team_launch<LOG2H, 512>(P, s);
    if (r != 0) return r;  // launched or failed; 0 = does not fit: try the smaller CTA
  }
  return stft_team_launch<LOG2H, 256>(P, s);
}

static int stft_launch_main(const StftParams& P, cudaStream_t s, size_t nframes, size_t smem_generic) {
  int done = 0;
  if (P.hop <= P.n_fft) {
    switch (P.log2h) {
      case 7: done = stft_team_launch_nt<7>(P, s); break;
      case 8: done = stft_team_launch_nt<8>(P, s); break;
      case 9: done = stft_team_launch_nt<9>(P, s); break;
      case 10: done = stft_team_launch_nt<10>(P, s); break;
      case 11: done = stft_team_launch_nt<11>(P, s); break;
      default: break;
    }
  }
  if (done < 0) return done;
  if (!done) {  // generic radix-2 kernel: any power-of-two n_fft in [64, 4096], any hop
    const unsigned grid = (unsigned)(nframes < 148u * 8u ? nframes : 148u * 8u);
    stft_mel_kernel<<<grid, 256, smem_generic, s>>>(P, nframes);
    MSTTS_CUDA(cudaGetLastError());
  }
  return MSTTS_OK;
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
// ONE single-block launch: returns at once when the tag matches (the steady state costs one ~2 us launch); otherwise builds
// window / twiddles / filter bank, compacts the non-zero filter weights and writes the tag last.
__global__ void __launch_bounds__(1024) stft_tables_kernel(int4* tag, int n_fft, int win, int n_mels, int sr, float* window, float2* tw,
                                                           float* fb, int* fb_range, int* fb_off, float* fbc) {
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
  __syncthreads();
  // compact (non-zero) filter weights: off[m] = prefix sum of the range lengths, fbc[off[m] + k - lo] = fb[m][k]
  __shared__ int off_s[257];
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
  __syncthreads();
  if (threadIdx.x == 0) *tag = make_int4(n_fft, win, n_mels, sr);
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
    stft_tables_kernel<<<1, 1024, 0, s>>>((int4*)(ws + o_tag), n_fft, win, nm, sr, (float*)(ws + o_win), (float2*)(ws + o_tw),
                                          (float*)(ws + o_fb), (int*)(ws + o_rng), (int*)(ws + o_off), (float*)(ws + o_fbc));
  }
  const size_t smem = (size_t)(3 * (n_fft / 2)) * sizeof(float2) + (size_t)n_fft * 4 + (size_t)(n_fft / 2 + 4) * 4 +
                      (size_t)3 * (n_fft / 2 + 1) * 4 + (size_t)3 * 256 * 4;
  MSTTS_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t nframes = (size_t)B * frames;
  MSTTS_REQUIRE(nframes < (1ull << 31), MSTTS_E_INVALID, "stft_mel: too many frames");
  int rc;
  if (!spectral_subtract) {
    if ((rc = stft_launch_main(P, s, nframes, smem))) return rc;
  } else {
    StftParams Q = P;
    Q.mag_out = (float*)(ws + o_mag);
    if ((rc = stft_launch_main(Q, s, nframes, smem))) return rc;
    const int nb = n_fft / 2 + 1;
    mag_mean_kernel<<<(B * nb + 127) / 128, 128, 0, s>>>((float*)(ws + o_mag), B, frames, nb, (float*)(ws + o_mean));
    P.mag_in = (float*)(ws + o_mag);
    P.mag_mean = (float*)(ws + o_mean);
    subtract_finish_kernel<<<(unsigned)nframes, 256, (size_t)nb * sizeof(float), s>>>(P);
  }
  MSTTS_CUDA(cudaGetLastError());
  return MSTTS_OK;
}
