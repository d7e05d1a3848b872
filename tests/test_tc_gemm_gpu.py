"""GPU parity of the hand-written tcgen05 GEMM (csrc/tc_gemm.cu, through the C ABI): C = A . B as bf16x3 with the hi/lo split on
chip, against the fp64 product of the same fp32 operands.  Shapes cover a partial last m-tile, a single tile, several rounds of
the persistent grid and the split-K tail (last round with <= 74 tiles)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [
    (100, 256, 64),        # one partial tile, one k-block
    (1000, 512, 320),      # 16 tiles, odd number of k-blocks
    (5000, 1024, 512),     # 160 tiles: 148 whole + 12 tiles in the split-K tail
    (19000, 1024, 256),    # 596 tiles: 4 rounds + tail of 4
    (4736, 1024, 128),     # exactly 148 tiles (no tail)
])
def test_tc_gemm_matches_fp64(cuda_dev, M, N, K):
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device=cuda_dev).manual_seed(M + N + K)
    A = torch.randn(M, K, device=cuda_dev, generator=g)
    Bt = torch.randn(N, K, device=cuda_dev, generator=g)
    out = torch.full((M, N), float('nan'), device=cuda_dev)
    Mt, Nt = (M + 127) // 128, N // 256
    ws = torch.empty(Mt * 128 * K * 4 + Nt * 256 * K * 4 + 74 * 256 * 128 * 4 + 8192, device=cuda_dev, dtype=torch.uint8)
    rc = lib.mstts_tc_gemm_test(_lib.ptr(A), _lib.ptr(Bt), M, N, K, _lib.ptr(out), C.c_void_p(ws.data_ptr()), ws.numel(),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mstts_tc_gemm_test")
    ref = (A.double() @ Bt.double().t())
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert torch.isfinite(out).all()
    assert err < 5e-5, err       # bf16x3: ~16 mantissa bits per operand, fp32 accumulation


def test_tc_gemm_refuses_bad_shapes(cuda_dev):
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    x = torch.zeros(256, 64, device=cuda_dev)
    ws = torch.empty(1 << 26, device=cuda_dev, dtype=torch.uint8)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.mstts_tc_gemm_test(_lib.ptr(x), _lib.ptr(x), 256, 200, 64, _lib.ptr(x), C.c_void_p(ws.data_ptr()), ws.numel(), st) == -1  # N % 256
    assert lib.mstts_tc_gemm_test(_lib.ptr(x), _lib.ptr(x), 256, 256, 40, _lib.ptr(x), C.c_void_p(ws.data_ptr()), ws.numel(), st) == -1  # K % 64
    assert lib.mstts_tc_gemm_test(_lib.ptr(x), _lib.ptr(x), 256, 256, 64, _lib.ptr(x), C.c_void_p(ws.data_ptr()), 1024, st) == -2        # workspace
