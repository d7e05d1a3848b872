"""GPU parity of the general dense-product front-end (mstts_gemm_f32 -> pack kernels -> tc_gemm_kernel<3> -> split-K reduce) against
the fp64 product of the same fp32 operands.  The shapes are the ones the hot path issues: the hoisted decoder products
(Modules.py:239-255,309-321: 80 / 81-wide sides, K = all decoder steps), the weight gradients (transposed A, long K, few tiles
-> split-K), the per-utterance attention products (batched, shared / strided operands), the convolution-as-one-product trick
(overlapping rows: leading dimension smaller than the row length) and accumulation into an existing C (beta = 1)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(lib, _lib, tA, tB, M, N, K, A, lda, sA, B, ldb, sB, Cm, ldc, sC, beta, batch, precise=0):
    rc = lib.mstts_gemm_f32(int(tA), int(tB), M, N, K, _lib.ptr(A), lda, sA, _lib.ptr(B), ldb, sB, _lib.ptr(Cm), ldc, sC, beta, batch, precise,
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mstts_gemm_f32")


@pytest.mark.parametrize("tA,tB,M,N,K,beta", [
    (0, 0, 100, 81, 1792, 0.0),      # projection: ragged N, ldc = 81 (scalar stores)
    (0, 0, 2000, 256, 80, 0.0),      # prenet layer 0: K = 80 (padded k-block)
    (0, 0, 3000, 4096, 256, 0.0),    # prenet rows of cell 0: many tiles, short K
    (1, 0, 1024, 4096, 3000, 0.0),   # weight gradient X^T dG: transposed A, long K
    (1, 0, 80, 256, 5000, 0.0),      # prenet0 weight gradient: one tile, split-K over ~78 slices
    (1, 0, 1792, 81, 2500, 0.0),     # projection weight gradient: ragged N + split-K
    (0, 1, 1500, 1024, 81, 0.0),     # d m1 = dproj . Wp^T: transposed B, K = 81
    (0, 1, 700, 768, 128, 1.0),      # dvalues += dkeys . Wm^T: beta = 1
    (1, 1, 300, 130, 70, 1.0),       # both transposed, nothing aligned, beta = 1
    (0, 0, 1, 1, 1, 0.0),            # degenerate
])
@pytest.mark.parametrize("precise", [0, 1, 2])
def test_gemm_f32_matches_fp64(cuda_dev, tA, tB, M, N, K, beta, precise):
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device=cuda_dev).manual_seed(1000 * M + 10 * N + K)
    A = torch.randn((K, M) if tA else (M, K), device=cuda_dev, generator=g)
    B = torch.randn((N, K) if tB else (K, N), device=cuda_dev, generator=g)
    C0 = torch.randn(M, N, device=cuda_dev, generator=g)
    out = C0.clone()
    _run(lib, _lib, tA, tB, M, N, K, A, A.shape[1], 0, B, B.shape[1], 0, out, N, 0, beta, 1, precise)
    opA = A.double().t() if tA else A.double()
    opB = B.double().t() if tB else B.double()
    ref = opA @ opB + beta * C0.double()
    assert torch.isfinite(out).all()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    # level 0: bf16x3, one accumulation chain (the truncating TMEM accumulator adds ~K * 1e-8); level 1: chains <= 512;
    # level 2 (3-way split, six products, chains <= 256, double reduce): fp32-SGEMM accuracy -- compared with the error of
    # the fp32 product itself
    sgemm_err = ((A.t() if tA else A) @ (B.t() if tB else B) + beta * C0 - ref).abs().max().item() / ref.abs().max().item()
    print("M=%d N=%d K=%d precise=%d: rel err %.2e (fp32 matmul: %.2e)" % (M, N, K, precise, err, sgemm_err))
    assert err < {0: 5e-5, 1: 1.5e-5, 2: max(1e-6, 2 * sgemm_err)}[precise], err


def test_gemm_f32_is_deterministic_with_split_k(cuda_dev):
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device=cuda_dev).manual_seed(7)
    A = torch.randn(6000, 200, device=cuda_dev, generator=g)   # transposed: M = 200, K = 6000
    B = torch.randn(6000, 300, device=cuda_dev, generator=g)
    outs = []
    for _ in range(3):
        out = torch.empty(200, 300, device=cuda_dev)
        _run(lib, _lib, 1, 0, 200, 300, 6000, A, 200, 0, B, 300, 0, out, 300, 0, 0.0, 1)
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("share_b", [False, True])
def test_gemm_f32_batched(cuda_dev, share_b):
    """per-utterance products: dvalues[b] = align_b^T dctx_b (batch strides inside time-major buffers), shared kernels"""
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device=cuda_dev).manual_seed(11)
    Bn, T, Te, D = 5, 90, 37, 256
    align = torch.randn(T, Bn, Te, device=cuda_dev, generator=g)     # A_b stored [K = T][M = Te], ld = Bn * Te, stride Te
    if share_b:
        dctx = torch.randn(T, D, device=cuda_dev, generator=g)
        ldb, sB = D, 0
    else:
        dctx = torch.randn(T, Bn, D, device=cuda_dev, generator=g)   # B_b [K = T][N = D], ld = Bn * D, stride D
        ldb, sB = Bn * D, D
    out = torch.full((Bn, Te, D), float('nan'), device=cuda_dev)
    _run(lib, _lib, 1, 0, Te, D, T, align, Bn * Te, Te, dctx, ldb, sB, out, D, Te * D, 0.0, Bn)
    for b in range(Bn):
        rhs = dctx.double() if share_b else dctx[:, b].double()
        ref = align[:, b].double().t() @ rhs
        err = (out[b].double() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 5e-5, (b, err)


def test_gemm_f32_overlapping_rows_is_a_convolution(cuda_dev):
    """rows of k*C values with leading dimension C over a zero-padded buffer = the im2col matrix of a 'same' convolution"""
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device=cuda_dev).manual_seed(13)
    Bn, T, Cin, Cout, k = 3, 50, 16, 24, 5
    p = k // 2
    x = torch.randn(Bn, T, Cin, device=cuda_dev, generator=g)
    w = torch.randn(k, Cin, Cout, device=cuda_dev, generator=g)
    xp = torch.zeros(Bn, T + 2 * p, Cin, device=cuda_dev)
    xp[:, p:p + T] = x
    out = torch.empty(Bn, T, Cout, device=cuda_dev)
    _run(lib, _lib, 0, 0, T, Cout, k * Cin, xp, Cin, (T + 2 * p) * Cin, w, Cout, 0, out, Cout, T * Cout, 0.0, Bn)
    ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double().permute(2, 1, 0), padding=p).transpose(1, 2)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-5, err


def test_gemm_f32_argument_errors(cuda_dev):
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    x = torch.zeros(64, 64, device=cuda_dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.mstts_gemm_f32(0, 0, 64, 64, 64, None, 64, 0, _lib.ptr(x), 64, 0, _lib.ptr(x), 64, 0, 0.0, 1, 0, st) == -1
    assert lib.mstts_gemm_f32(0, 0, 64, 64, 0, _lib.ptr(x), 64, 0, _lib.ptr(x), 64, 0, _lib.ptr(x), 64, 0, 0.0, 1, 0, st) == -1
    assert lib.mstts_gemm_f32(0, 0, 0, 64, 64, _lib.ptr(x), 64, 0, _lib.ptr(x), 64, 0, _lib.ptr(x), 64, 0, 0.0, 1, 0, st) == 0   # empty product
