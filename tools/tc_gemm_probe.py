#!/usr/bin/env python
"""Parity + throughput probe of the hand-written tcgen05 GEMM (csrc/tc_gemm.cu) against torch / cuBLAS bf16.
python tools/tc_gemm_probe.py [M N K]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_speaker_tts_b200 import _lib

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (16000, 1024, 6528)
dev = torch.device("cuda:0")
lib = _lib.lib()
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(M, K, device=dev, generator=g)
Bt = torch.randn(N, K, device=dev, generator=g)
Cm = torch.zeros(M, N, device=dev)
Mt, Nt = (M + 127) // 128, N // 256
nbytes = Mt * 128 * K * 4 + Nt * 256 * K * 4 + 4096 + 74 * 256 * 128 * 4 + 1024  # hi + lo tiles of both operands
ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.mstts_tc_gemm_test(_lib.ptr(A), _lib.ptr(Bt), M, N, K, _lib.ptr(Cm), C.c_void_p(ws.data_ptr()), ws.numel(), st), "tc_gemm_test")
torch.cuda.synchronize()
Ab, Bb = A.bfloat16(), Bt.bfloat16()
ref = (A.double() @ Bt.double().t()).float()
err = (Cm - ref).abs().max().item() / ref.abs().max().item()
print("M=%d N=%d K=%d  bf16x3 (on-chip hi/lo) max rel err vs the fp64 product of the fp32 operands: %.3e" % (M, N, K, err))
base = (ws.data_ptr() + 1023) & ~1023
a_t, b_t = C.c_void_p(base), C.c_void_p(base + Mt * 128 * K * 4)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


scr = C.c_void_p(base + Mt * 128 * K * 4 + Nt * 256 * K * 4 + 256)
ms0 = timeit(lambda: lib.mstts_tc_gemm_tiled(a_t, b_t, M, N, K, _lib.ptr(Cm), N, None, st))
ms = timeit(lambda: lib.mstts_tc_gemm_tiled(a_t, b_t, M, N, K, _lib.ptr(Cm), N, scr, st))
err2 = (Cm - ref).abs().max().item() / ref.abs().max().item()
print('split-K tail: %.3f ms (whole tiles only: %.3f ms), max rel err %.3e' % (ms, ms0, err2))
Bn = Bb.t().contiguous()
out = torch.empty(M, N, device=dev)
ms_ref = timeit(lambda: torch.matmul(Ab, Bn))
fl = 2.0 * M * N * K
print("tcgen05 bf16x3 kernel %.3f ms = %.0f TFLOP/s executed (3 MMAs per product; %.0f algorithmic)   |   one plain bf16 matmul "
      "(cuBLAS, bf16 out) %.3f ms = %.0f TFLOP/s" % (ms, 3 * fl / ms / 1e9, fl / ms / 1e9, ms_ref, fl / ms_ref / 1e9))
