// Stream-ordered scratch for the dense-product front-ends (operand tile images, split-K slabs).  Backed by the device's
// default CUDA memory pool with the release threshold lifted, so after the first step an allocation is a pointer bump on
// the stream: no synchronisation, safe under several streams and under CUDA-graph capture.  mstts_release_scratch() trims it.
#pragma once
#include "common.cuh"

int scratch_alloc(void** p, size_t bytes, cudaStream_t s);
void scratch_free(void* p, cudaStream_t s);

// a set of allocations freed together (on every return path)
struct ScratchScope {
  cudaStream_t s;
  void* ptrs[8];
  int n = 0;
  explicit ScratchScope(cudaStream_t st) : s(st) {}
  ~ScratchScope() {
    for (int i = n - 1; i >= 0; --i) scratch_free(ptrs[i], s);
  }
  int get(void** p, size_t bytes) {
    if (n >= 8) {
      mstts_set_error("ScratchScope: too many allocations");
      return MSTTS_E_INVALID;
    }
    int rc = scratch_alloc(p, bytes, s);
    if (rc == MSTTS_OK) ptrs[n++] = *p;
    return rc;
  }
};
