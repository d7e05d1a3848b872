// Skinny GEMV building block shared by the forward and reverse persistent kernels (fp32 parity mode).
//   out[b][col] = sum_{k in [k0,k1)} x[b][k] * W[k][col]      x = [xa (Ka) | xb (Kb)] per batch row
// One lane owns one output column (col < 0: idle lane) and NB batch accumulators; the K range is split
// over warps (and half-warps) by the caller, which also owns the cross-slice reduction.
#pragma once
#include "common.cuh"

constexpr int kRedStride = 40;  // floats per (slice, batch) row of the K-split reduction buffer

template <int NB>
__device__ __forceinline__ void gemv_acc(const float* __restrict__ W, int ldw, int col, int k0, int k1, const float* xa,
                                         int Ka, const float* xb, int Kb, int nb, float (&acc)[NB]) {
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = 0.f;
  const bool active = col >= 0;
  const float* wcol = W + (active ? col : 0);
#pragma unroll 2
  for (int k = k0; k < k1; k += 4) {
    const float* xp;
    int xs;
    if (k < Ka) {
      xp = xa + k;
      xs = Ka;
    } else {
      xp = xb + (k - Ka);
      xs = Kb;
    }
    const float* wp = wcol + (size_t)k * ldw;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
    if (active) {
      w0 = ld_nc_na(wp);
      w1 = ld_nc_na(wp + ldw);
      w2 = ld_nc_na(wp + 2 * (size_t)ldw);
      w3 = ld_nc_na(wp + 3 * (size_t)ldw);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (b < nb) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + (size_t)b * xs);
        acc[b] = fmaf(xv.x, w0, acc[b]);
        acc[b] = fmaf(xv.y, w1, acc[b]);
        acc[b] = fmaf(xv.z, w2, acc[b]);
        acc[b] = fmaf(xv.w, w3, acc[b]);
      }
    }
  }
}

// forward cells: 32 lanes = this CTA's 32 gate columns (4 gates x 8 units), K split over the 8 warps.
// red layout: [8 slices][NB][kRedStride], column index = gate*8 + unit
template <int NB>
__device__ __forceinline__ void lstm_gemv(const float* __restrict__ W, int K, const float* xa, int Ka, const float* xb,
                                          int Kb, int nb, float* red, int unit0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = (lane >> 3) * kCell + unit0 + (lane & 7);
  const int kslice = K >> 3;
  float acc[NB];
  gemv_acc<NB>(W, kGates, col, warp * kslice, (warp + 1) * kslice, xa, Ka, xb, Kb, nb, acc);
#pragma unroll
  for (int b = 0; b < NB; ++b) red[(warp * NB + b) * kRedStride + lane] = acc[b];
}
