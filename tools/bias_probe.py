#!/usr/bin/env python
"""Is the error of the tcgen05 product biased?  mean(out - ref) against the fp64 product, per precision level, for zero-mean and
for positive operands, next to the fp32 library matmul (round-to-nearest accumulation)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_speaker_tts_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for K in (81, 256, 1024, 4096):
    for kind in ("zero-mean", "positive"):
        M, N = 2048, 512
        A = torch.randn(M, K, device=dev, generator=g)
        B = torch.randn(K, N, device=dev, generator=g)
        if kind == "positive":
            A, B = A.abs(), B.abs()
        ref = A.double() @ B.double()
        scale = (A.double().abs() @ B.double().abs()).mean().item()      # mean of sum |a||b|
        line = "K=%5d %-9s" % (K, kind)
        for prec in (0, 1, 2):
            out = torch.empty(M, N, device=dev)
            rc = lib.mstts_gemm_f32(0, 0, M, N, K, _lib.ptr(A), K, 0, _lib.ptr(B), N, 0, _lib.ptr(out), N, 0, 0.0, 1, prec, st)
            _lib.check(rc, "gemm")
            e = out.double() - ref
            line += " | p%d mean %+.2e rms %.2e" % (prec, e.mean().item() / scale, e.pow(2).mean().sqrt().item() / scale)
        e = (A @ B).double() - ref
        line += " | fp32 mean %+.2e rms %.2e" % (e.mean().item() / scale, e.pow(2).mean().sqrt().item() / scale)
        print(line)
