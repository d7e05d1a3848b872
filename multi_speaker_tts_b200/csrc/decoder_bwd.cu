// Reverse pass of the decoder loop (placeholder until the reverse-time kernel lands).
#include "common.cuh"
#include "gemm.h"

size_t dec_bwd_extra_bytes(int, int, int, int) { return 0; }

extern "C" int mstts_decoder_bwd(const MsttsDecoderWeights*, const MsttsDecoderIO*, const MsttsDecoderGrads*,
                                 const MsttsDecoderWeightGrads*, void*, size_t, void*) {
  mstts_set_error("decoder_bwd: not implemented yet");
  return MSTTS_E_UNSUPPORTED;
}
