"""CPU ORACLE (test infrastructure, not product code) -- Audio.py feature extraction.

numpy restatement of Audio.melspectrogram / spectrogram / spectrogram_and_mel (Audio.py:19-48,62-96) including what
librosa / scipy do on the reference's behalf (they are not installed here):
  * scipy.signal.lfilter([1,-0.97],[1],x)                      Audio.py:12-13
  * librosa.stft(y, n_fft, hop_length, win_length): center=True, reflect padding of n_fft/2, periodic Hann(win_length)
    zero-padded (centred) to n_fft, frames = 1 + len(y)//hop, complex64 result      Audio.py:62-64
  * librosa.filters.mel(sr, n_fft, n_mels): slaney scale (htk=False), fmin 0, fmax sr/2, slaney area norm, float32
PARITY UNPINNED (no reference tests, librosa absent); pinned by cross-checks against torch.stft and
torchaudio.functional.melscale_fbanks in tests/test_audio_oracle_cpu.py.
"""
import numpy as np


def preemphasis(x, coef=0.97):
    y = np.array(x, dtype=np.float64, copy=True)
    y[1:] -= coef * np.asarray(x, dtype=np.float64)[:-1]
    return y


def stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate):
    n_fft = (num_freq - 1) * 2
    hop = int(frame_shift_ms / 1000 * sample_rate)
    win = int(frame_length_ms / 1000 * sample_rate)
    return n_fft, hop, win


def hann_periodic_centered(win, n_fft):
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win) / win)
    out = np.zeros(n_fft)
    off = (n_fft - win) // 2
    out[off:off + win] = w
    return out


def stft(y, n_fft, hop, win):
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, n_fft // 2, mode='reflect')
    frames = 1 + len(y) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(frames)[:, None]
    fr = yp[idx] * hann_periodic_centered(win, n_fft)[None, :]
    return np.fft.rfft(fr, axis=1).astype(np.complex64).T  # [bins, frames] like librosa


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sample_rate, n_fft, n_mels):
    fftfreqs = np.linspace(0, sample_rate / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sample_rate / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def _magnitude(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spectral_subtract=False):
    n_fft, hop, win = stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    M = np.abs(stft(preemphasis(y), n_fft, hop, win))
    if spectral_subtract:
        M = np.clip(M - np.mean(M, axis=1, keepdims=True) / 10, a_min=0.0, a_max=np.inf)
    return M


def _amp_to_db(x):
    return 20 * np.log10(np.maximum(1e-5, x))


def _normalize(S, min_level_db=-100):
    return np.clip((S - min_level_db) / -min_level_db, 0, 1)


def _symmetric_normalize(S, min_level_db=-100, max_abs_value=4):
    return np.clip((2 * max_abs_value) * ((S - min_level_db) / (-min_level_db)) - max_abs_value, -max_abs_value, max_abs_value)


def spectrogram(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, ref_level_db=20, spectral_subtract=False):
    M = _magnitude(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spectral_subtract)
    return _normalize(_amp_to_db(M) - ref_level_db)


def melspectrogram(y, num_freq, frame_shift_ms, frame_length_ms, num_mels, sample_rate, max_abs_value=None, spectral_subtract=False):
    M = _magnitude(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spectral_subtract)
    n_fft = (num_freq - 1) * 2
    S = _amp_to_db(np.dot(mel_basis(sample_rate, n_fft, num_mels), M))
    return _normalize(S) if max_abs_value is None else _symmetric_normalize(S, max_abs_value=max_abs_value)


def spectrogram_and_mel(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spect_ref_level_db=20, num_mels=80, max_abs_mels=None,
                        spectral_subtract=False):
    M = _magnitude(y, num_freq, frame_shift_ms, frame_length_ms, sample_rate, spectral_subtract)
    n_fft = (num_freq - 1) * 2
    spect_S = _normalize(_amp_to_db(M) - spect_ref_level_db)
    mel_S = _amp_to_db(np.dot(mel_basis(sample_rate, n_fft, num_mels), M))
    mel_S = _normalize(mel_S) if max_abs_mels is None else _symmetric_normalize(mel_S, max_abs_value=max_abs_mels)
    return spect_S, mel_S
