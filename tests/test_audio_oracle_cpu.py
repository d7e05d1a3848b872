"""CPU tests of the audio oracle: librosa / scipy semantics cross-checked against torch.stft and torchaudio."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as A


def test_preemphasis_is_lfilter():
    x = np.random.default_rng(0).standard_normal(50)
    y = A.preemphasis(x)
    assert y[0] == x[0] and np.allclose(y[1:], x[1:] - 0.97 * x[:-1])


@pytest.mark.parametrize("n_fft,hop,win,S", [(1024, 256, 1024, 5000), (2048, 200, 800, 7777), (256, 64, 200, 1000)])
def test_stft_matches_torch_stft(n_fft, hop, win, S):
    x = np.random.default_rng(1).standard_normal(S)
    D = A.stft(x, n_fft, hop, win)
    t = torch.stft(torch.from_numpy(x), n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win, periodic=True, dtype=torch.float64),
                   center=True, pad_mode='reflect', return_complex=True)
    assert D.shape == (n_fft // 2 + 1, 1 + S // hop)
    assert np.allclose(D, t.numpy()[:, :D.shape[1]], atol=1e-4)


@pytest.mark.parametrize("sr,n_fft", [(16000, 2048), (22050, 1024)])
def test_mel_basis_matches_torchaudio_slaney(sr, n_fft):
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2, 80, sr, norm='slaney', mel_scale='slaney').T.numpy()
    ours = A.mel_basis(sr, n_fft, 80)
    assert ours.shape == fb.shape
    assert np.allclose(ours, fb, atol=2e-6)
    assert (ours >= 0).all() and (ours.sum(1) > 0).all()


def test_melspectrogram_ranges_and_shapes():
    x = np.random.default_rng(2).uniform(-0.9, 0.9, 16000)
    m = A.melspectrogram(x, 1025, 12.5, 50, 80, 16000, max_abs_value=4)
    assert m.shape == (80, 1 + 16000 // 200) and m.min() >= -4 and m.max() <= 4
    s, m2 = A.spectrogram_and_mel(x, 1025, 12.5, 50, 16000, num_mels=80, max_abs_mels=4)
    assert s.shape == (1025, 81) and 0 <= s.min() and s.max() <= 1 and np.allclose(m, m2)
    m3 = A.melspectrogram(x, 1025, 12.5, 50, 80, 16000, max_abs_value=4, spectral_subtract=True)
    assert (m3 <= m + 1e-9).all()


def test_golden_fixture_reproduces():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_mel.npz"))
    x = np.random.default_rng(5).uniform(-0.9, 0.9, 6400).astype(np.float32)
    assert np.abs(A.melspectrogram(x, 1025, 12.5, 50, 80, 16000, max_abs_value=4) - g["mel_ref_defaults"]).max() < 1e-5
    assert np.abs(A.melspectrogram(x, 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4) - g["mel_config4"]).max() < 1e-5
