"""GPU parity of the fused STFT+mel kernel (C ABI) vs the numpy oracle: features L_inf < 1e-3 in normalised units."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _wav(S, seed, B=1):
    return np.random.default_rng(seed).uniform(-0.99, 0.99, (B, S)).astype(np.float32)


@pytest.mark.parametrize("num_freq,shift,length,sr,S,sub", [
    (1025, 12.5, 50, 16000, 16000, False),          # reference defaults (Hyper_Parameters.py:5-10)
    (513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 22050, 22050, False),   # BASELINE config 4 parameters
    (1025, 12.5, 50, 16000, 9000, True),            # spectral subtraction (Pattern_Generate.py:52)
    (129, 4, 12.5, 16000, 3000, False),
    (257, 8, 25, 16000, 5000, False),               # n_fft 512: radix-4 first stage, one frame per warp
    (2049, 12.5, 50, 16000, 20000, False),          # n_fft 4096: 256-thread teams
])
def test_melspectrogram_parity(cuda_dev, num_freq, shift, length, sr, S, sub):
    from oracle import audio_oracle as A
    from multi_speaker_tts_b200 import Audio as G
    x = _wav(S, 3)[0]
    ref = A.melspectrogram(x, num_freq, shift, length, 80, sr, max_abs_value=4, spectral_subtract=sub)
    got = G.melspectrogram(x, num_freq, shift, length, 80, sr, max_abs_value=4, spectral_subtract=sub)
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    print("mel Linf %.3e" % err)
    assert err < TOL
    sref, mref = A.spectrogram_and_mel(x, num_freq, shift, length, sr, num_mels=80, max_abs_mels=None, spectral_subtract=sub)
    sgot, mgot = G.spectrogram_and_mel(x, num_freq, shift, length, sr, num_mels=80, max_abs_mels=None, spectral_subtract=sub)
    assert np.abs(sgot - sref).max() < TOL and np.abs(mgot - mref).max() < TOL
    assert np.abs(G.spectrogram(x, num_freq, shift, length, sr, spectral_subtract=sub) - sref).max() < TOL


def test_batched_device_input(cuda_dev):
    from oracle import audio_oracle as A
    from multi_speaker_tts_b200 import Audio as G
    x = _wav(8000, 5, B=3)
    got = G.melspectrogram(torch.from_numpy(x).to(cuda_dev), 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4)
    assert got.is_cuda and got.shape == (3, 80, 1 + 8000 // 256)
    for b in range(3):
        ref = A.melspectrogram(x[b], 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4)
        assert np.abs(got[b].cpu().numpy() - ref).max() < TOL


def test_bad_arguments_are_refused(cuda_dev):
    from multi_speaker_tts_b200 import Audio as G
    from multi_speaker_tts_b200._lib import MsttsError
    with pytest.raises(MsttsError):
        G.stft_features(_wav(4000, 1), 1000, 250, 1000, 16000, 80)      # n_fft not a power of two
    with pytest.raises(MsttsError):
        G.stft_features(_wav(100, 1), 1024, 256, 1024, 16000, 80)       # too short to reflect-pad


def test_against_committed_golden(cuda_dev):
    import os
    from multi_speaker_tts_b200 import Audio as G
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_mel.npz"))
    x = np.random.default_rng(5).uniform(-0.9, 0.9, 6400).astype(np.float32)
    assert np.abs(G.melspectrogram(x, 1025, 12.5, 50, 80, 16000, max_abs_value=4) - g["mel_ref_defaults"]).max() < TOL
    assert np.abs(G.melspectrogram(x, 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4) - g["mel_config4"]).max() < TOL
    assert abs(float(G.spectrogram(x, 1025, 12.5, 50, 16000).sum()) - float(g["spec_ref_defaults_checksum"])) < 1e-2
