"""CPU tests of the host side: config tree, C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hyper_parameters_tree():
    from multi_speaker_tts_b200 import Hyper_Parameters as hp
    assert hp.Sound.Mel_Dim == 80 and hp.Sound.Sample_Rate == 16000 and hp.Sound.Max_Abs_Mel == 4
    assert hp.Decoder.LSTM.Cell_Size == 1024 and hp.Decoder.LSTM.Nums == 2 and hp.Decoder.LSTM.Zoneout_Rate == 0.1
    assert hp.Decoder.LSTM.Max_Inference_Length == 1000
    assert hp.Decoder.PreNet.Size == 256 and hp.Decoder.PreNet.Dropout_Rate == 0.5 and hp.Decoder.PreNet.Use_Dropout
    assert hp.Attention.Memory_Size == 128 and hp.Attention.Conv.Kernel_Size == 31 and hp.Attention.Conv.Channel == 32
    assert hp.Encoder.BiLSTM.Cell_Size == 256 and hp.Speaker_Embedding.Embedding_Size == 256
    assert hp.Train.ADAM.Epsilon == 1e-6 and hp.Train.Learning_Rate.Decay_Step == 10000
    assert hp.WaveGlow.Flows == 12 and hp.WaveGlow.WaveNet.Channels == 512 and hp.WaveGlow.Train.Max_Signal_Length == 8000
    assert hp.Use_Vocoder in ('WaveGlow', 'Taco1_Mel_to_Spect')


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mstts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mstts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multi_speaker_tts_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libmstts_b200.so is not built: run `python __graft_entry__.py build`")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(names) == set(_lib.EXPORTS), (set(names) ^ set(_lib.EXPORTS))
    lib.mstts_version.restype = ctypes.c_int
    assert lib.mstts_version() == 100


def test_workspace_query_and_argument_errors_without_gpu():
    from multi_speaker_tts_b200 import _lib
    lib = _lib.lib()
    assert lib.mstts_decoder_workspace_bytes(32, 128, 800, 768, 801, 0) > 2 * 10 ** 9
    assert lib.mstts_decoder_workspace_bytes(0, 128, 800, 768, 801, 0) == 0
    # null arguments are rejected before any CUDA call
    assert lib.mstts_decoder_fwd(None, None, None, 0, None) == -1
    assert b"null" in lib.mstts_last_error()
    assert lib.mstts_adam_tf(None, None, None, None, 16, 0.1, 0.9, 0.999, 1e-6, 1.0, 0.0, None) == -1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "multi_speaker_tts_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_library_links_no_vendor_gemm():
    """Every dense product runs on the hand-written tcgen05 kernel: the shared library must not depend on cuBLAS / cuBLASLt /
    cuDNN / cuFFT (checked on the ELF's DT_NEEDED entries, no GPU required)."""
    import subprocess
    from multi_speaker_tts_b200 import _lib
    out = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("readelf not available")
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out.stdout)
    assert needed, out.stdout
    for n in needed:
        assert not re.search(r"cublas|cudnn|cufft|cutlass|nvjet", n, flags=re.I), needed
    src = os.path.join(ROOT, "multi_speaker_tts_b200", "csrc")
    for f in os.listdir(src):
        txt = open(os.path.join(src, f)).read()
        assert "cublas_v2.h" not in txt and "cublasLt" not in txt and "cudnn.h" not in txt, f
