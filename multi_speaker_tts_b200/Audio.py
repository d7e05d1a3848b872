"""Audio.py surface of the reference on the GPU: ``melspectrogram``, ``spectrogram``, ``spectrogram_and_mel`` keep the
reference signatures (Audio.py:19-40) and return arrays shaped like the reference ([dim, frames] per waveform).  Inputs may
be a 1-D waveform (numpy or torch) or a batch [B, S] of equal-length waveforms (torch CUDA tensor: stays on device).
Inverse transforms / Griffin-Lim (Audio.py:15-16,24-27,50-68,89-99) are out of the hot-path scope.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate):
    n_fft = (num_freq - 1) * 2
    hop_length = int(frame_shift_ms / 1000 * sample_rate)
    win_length = int(frame_length_ms / 1000 * sample_rate)
    return n_fft, hop_length, win_length


_WS_CACHE = {}  # (device, n_fft, win, n_mels, sr) -> workspace tensor whose tables are already built


def stft_features(wav, n_fft, hop, win, sample_rate, num_mels=None, max_abs_value=None, spectral_subtract=False,
                  want_mel=True, want_spec=False, device=None):
    """wav [B,S] or [S] -> (mel [B,frames,num_mels] | None, spec [B,frames,n_fft/2+1] | None) on the GPU"""
    lib = _lib.lib()
    if isinstance(wav, np.ndarray):
        wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
    if not wav.is_cuda:
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("multi_speaker_tts_b200.Audio runs on CUDA only (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        wav = wav.to(device)
    wav = wav.float().contiguous()
    if wav.dim() == 1:
        wav = wav[None]
    B, S = wav.shape
    frames = 1 + S // hop
    dev = wav.device
    n_mels = int(num_mels) if want_mel else 0
    mel = torch.empty(B, frames, n_mels, device=dev) if want_mel else None
    spec = torch.empty(B, frames, n_fft // 2 + 1, device=dev) if want_spec else None
    nbytes = lib.mstts_stft_mel_workspace_bytes(B, S, n_fft, hop, max(n_mels, 1), int(bool(spectral_subtract)))
    key = (dev, n_fft, win, n_mels, int(sample_rate))
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, device=dev, dtype=torch.uint8)  # zero tag: tables get built on first use
        _WS_CACHE[key] = ws
    with torch.cuda.device(dev):
        rc = lib.mstts_stft_mel(_lib.ptr(wav), B, S, n_fft, hop, win, n_mels, int(sample_rate),
                                float(max_abs_value) if max_abs_value is not None else 0.0, int(bool(spectral_subtract)),
                                _lib.ptr(mel), _lib.ptr(spec), C.c_void_p(ws.data_ptr()), ws.numel(),
                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "mstts_stft_mel")
    return mel, spec


def _like_reference(t, was_batch):
    t = t.transpose(1, 2)  # [B, dim, frames]: the reference returns [dim, frames]
    return t if was_batch else t[0].cpu().numpy()


def melspectrogram(y, num_freq, frame_shift_ms, frame_length_ms, num_mels, sample_rate, max_abs_value=None, spectral_subtract=False):
    n_fft, hop, win = _stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    was_batch = hasattr(y, 'dim') and y.dim() == 2
    mel, _ = stft_features(y, n_fft, hop, win, sample_rate, num_mels, max_abs_value, spectral_subtract)
    return _like_reference(mel, was_batch)


def spectrogram(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, ref_level_db=20, spectral_subtract=False):
    if ref_level_db != 20:
        raise NotImplementedError("ref_level_db is fixed at the reference default (20)")
    n_fft, hop, win = _stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    was_batch = hasattr(y, 'dim') and y.dim() == 2
    _, spec = stft_features(y, n_fft, hop, win, sample_rate, None, None, spectral_subtract, want_mel=False, want_spec=True)
    return _like_reference(spec, was_batch)


def spectrogram_and_mel(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spect_ref_level_db=20, num_mels=80,
                        max_abs_mels=None, spectral_subtract=False):
    if spect_ref_level_db != 20:
        raise NotImplementedError("spect_ref_level_db is fixed at the reference default (20)")
    n_fft, hop, win = _stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    was_batch = hasattr(y, 'dim') and y.dim() == 2
    mel, spec = stft_features(y, n_fft, hop, win, sample_rate, num_mels, max_abs_mels, spectral_subtract, want_spec=True)
    return _like_reference(spec, was_batch), _like_reference(mel, was_batch)
