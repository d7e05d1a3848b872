"""Feeder surface of the reference (Feeder.py): ``placeholder_Dict`` keys / dtypes, ``Get_Train_Pattern`` and
``Get_Inference_Pattern`` returning feed dicts keyed by the placeholder objects.

Data sources, in order: the reference's pickled dataset under ``hp.Train.Pattern_Path`` (METADATA.PICKLE + one pickle
per utterance with 'Token' and 'Mel', Pattern_Generate.py:66-76) when it exists; otherwise seeded synthetic patterns of
the shapes in SURVEY 8d (there is no dataset and no network in the build image).  The feed values are host numpy arrays,
exactly like the reference's; ``MSTTS_SV.Tacotron2`` moves them to the device through pinned memory.
"""
import os
import pickle
import time
import wave
import zlib
from collections import deque
from random import shuffle
from threading import Thread

import numpy as np

from . import Hyper_Parameters as hp


def _current_cuda_device():
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else None
    except (ImportError, RuntimeError):
        return None


def _pinned(a):
    """The same array in page-locked host memory when a CUDA device is present (plain ndarray otherwise): the collation runs in the
    feeder's background thread, so the train step's host -> device copies become DMA transfers with no staging copy inside the step."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    except (ImportError, RuntimeError):
        pass
    return a

# Token_Index_Dict.json of the reference: <S>, <E>, then the printable characters in ASCII order (42 symbols)
TOKEN_INDEX_DICT = {t: i for i, t in enumerate(['<S>', '<E>'] + list(' !"\'(),-.:;?ABCDEFGHIJKLMNOPQRSTUVWXYZ[]'))}


class Placeholder(object):
    """hashable stand-in for a tf.placeholder: name, dtype, static shape"""

    def __init__(self, name, dtype, shape):
        self.name, self.dtype, self.shape = name, dtype, shape

    def __repr__(self):
        return "Placeholder(%s, %s, %s)" % (self.name, np.dtype(self.dtype).name, self.shape)


class Feeder(object):
    def __init__(self, is_Training=False, synthetic=None, seed=1234, rank=0, synthetic_shape=None):
        self.is_Training = is_Training
        self.Placeholder_Generate()
        self._rng = np.random.default_rng(seed + rank)
        self._synthetic_shape = synthetic_shape  # (B, Te, L) or None: hp.Train.Batch_Size and ragged lengths
        self._cuda_device = _current_cuda_device()  # the generator thread pins its arrays in this device's context
        self._pattern_path = hp.Train.Pattern_Path  # captured: the generator thread outlives later changes of hp
        self._batch_size = hp.Train.Batch_Size
        meta = os.path.join(self._pattern_path, hp.Train.Metadata_File.upper()).replace("\\", "/")
        self.synthetic = (not os.path.exists(meta)) if synthetic is None else synthetic
        self.Metadata_Load()
        if self.is_Training and not self.synthetic:
            if hp.Train.Use_Pre_in_Main_Train:
                self.pre_Pattern_Queue = deque()
                t = Thread(target=self.Train_Pattern_Generate, args=[True])
                t.daemon = True
                t.start()
            self.pattern_Queue = deque()
            t = Thread(target=self.Train_Pattern_Generate, args=[False])
            t.daemon = True
            t.start()

    def Placeholder_Generate(self):
        """Feeder.py:34-41"""
        m = hp.Sound.Mel_Dim
        self.placeholder_Dict = {
            "Is_Training": Placeholder("is_training_placeholder", np.bool_, ()),
            "Token": Placeholder("token_placeholder", np.int32, (None, None)),
            "Token_Length": Placeholder("token_length_placeholder", np.int32, (None,)),
            "Mel": Placeholder("mel_placeholder", np.float32, (None, None, m)),
            "Mel_Length": Placeholder("mel_length_placeholder", np.int32, (None,)),
            "Speaker_Embedding_Mel": Placeholder("speaker_embedding_mel_placeholder", np.float32, (None, None, m)),
        }

    def Metadata_Load(self):
        """Feeder.py:43-62"""
        if self.is_Training and not self.synthetic:
            with open(os.path.join(self._pattern_path, hp.Train.Metadata_File.upper()).replace("\\", "/"), 'rb') as f:
                self.metadata_Dict = pickle.load(f)
            if not all([
                    len(self.metadata_Dict['Token_Index_Dict']) == hp.Encoder.Embedding.Token_Size,
                    self.metadata_Dict['Spectrogram_Dim'] == hp.Sound.Spectrogram_Dim,
                    self.metadata_Dict['Mel_Dim'] == hp.Sound.Mel_Dim,
                    self.metadata_Dict['Frame_Shift'] == hp.Sound.Frame_Shift,
                    self.metadata_Dict['Frame_Length'] == hp.Sound.Frame_Length,
                    self.metadata_Dict['Sample_Rate'] == hp.Sound.Sample_Rate]):
                raise ValueError('The metadata information and hyper parameter setting are not consistent.')
        else:
            self.metadata_Dict = {'Token_Index_Dict': dict(TOKEN_INDEX_DICT)}

    def Speaker_Embedding_Mel(self, mel_List):
        """Feeder.py:64-87: Sample_Nums windows of Mel_Frame frames, hop Overlap_Frame, centred in the utterance;
        a too-short utterance is copied (zero padded) into every window.  -> [len * Sample_Nums, Mel_Frame, Mel_Dim]"""
        inf = hp.Speaker_Embedding.Inference
        required = inf.Sample_Nums * (inf.Mel_Frame - inf.Overlap_Frame) + inf.Overlap_Frame
        out = np.zeros((len(mel_List), inf.Sample_Nums, inf.Mel_Frame, hp.Sound.Mel_Dim), dtype=np.float32)
        for index, mel in enumerate(mel_List):
            if mel.shape[0] < required:
                sample = mel[:inf.Mel_Frame]
                out[index, :, :sample.shape[0]] = sample
            else:
                for s in range(inf.Sample_Nums):
                    start = int((mel.shape[0] - required) / 2) + s * inf.Overlap_Frame
                    out[index, s] = mel[start:start + inf.Mel_Frame]
        return np.reshape(out, (-1, inf.Mel_Frame, hp.Sound.Mel_Dim))

    def _collate(self, token_List, mel_List):
        """Feeder.py:146-172: tokens padded with <E>, mels with zeros"""
        n = len(token_List)
        tok = np.zeros((n, max(t.shape[0] for t in token_List)), dtype=np.int32) + self.metadata_Dict['Token_Index_Dict']['<E>']
        mel = np.zeros((n, max(m.shape[0] for m in mel_List), hp.Sound.Mel_Dim), dtype=np.float32)
        for i, (t, m) in enumerate(zip(token_List, mel_List)):
            tok[i, :t.shape[0]] = t
            mel[i, :m.shape[0]] = m
        p = self.placeholder_Dict
        return {
            p["Is_Training"]: True,
            p["Token"]: _pinned(tok),
            p["Token_Length"]: _pinned(np.array([t.shape[0] for t in token_List]).astype(np.int32)),
            p["Mel"]: _pinned(mel),
            p["Mel_Length"]: _pinned(np.array([m.shape[0] for m in mel_List]).astype(np.int32)),
            p["Speaker_Embedding_Mel"]: _pinned(self.Speaker_Embedding_Mel(mel_List)),
        }

    def Train_Pattern_Generate(self, is_Pre_Train=False):
        """Feeder.py:89-174 (background thread over the pickled dataset)"""
        if self._cuda_device is not None:  # a new thread starts on device 0: page-locking there would open a context on GPU 0
            import torch                   # from every rank of a data-parallel job
            torch.cuda.set_device(self._cuda_device)
        md = self.metadata_Dict
        wanted = hp.Train.Pre_Train_Dataset_List if is_Pre_Train else hp.Train.Main_Train_Dataset_List
        queue = self.pre_Pattern_Queue if is_Pre_Train else self.pattern_Queue
        file_List = [p for p in md['File_List'] if md['Dataset_Dict'][p] in wanted]
        lo = hp.Train.Use_Wav_Length_Range[0] / hp.Sound.Frame_Shift
        hi = hp.Train.Use_Wav_Length_Range[1] / hp.Sound.Frame_Shift
        path_List = [(p, md['Mel_Length_Dict'][p]) for p in file_List if lo <= md['Mel_Length_Dict'][p] <= hi]
        if hp.Train.Pattern_Sorting_by_Mel_Length:
            path_List = [p for p, _ in sorted(path_List, key=lambda x: x[1])]
        else:
            path_List = [p for p, _ in path_List]
        S, E = md['Token_Index_Dict']['<S>'], md['Token_Index_Dict']['<E>']
        while True:
            if not hp.Train.Pattern_Sorting_by_Mel_Length:
                shuffle(path_List)
            batches = [path_List[x:x + self._batch_size] for x in range(0, len(path_List), self._batch_size)]
            shuffle(batches)
            i = 0
            while i < len(batches):
                if len(queue) >= hp.Train.Max_Pattern_Queue:
                    time.sleep(0.1)
                    continue
                token_List, mel_List = [], []
                for file_Path in batches[i]:
                    with open(os.path.join(self._pattern_path, file_Path).replace("\\", "/"), "rb") as f:
                        pattern_Dict = pickle.load(f)
                    token_List.append(np.hstack([S, pattern_Dict['Token'], E]).astype(np.int32))
                    mel_List.append(pattern_Dict['Mel'])
                queue.append(self._collate(token_List, mel_List))
                i += 1

    def _synthetic_pattern(self):
        """tokens ~ U{2..41} between <S> and <E>, mel ~ clip(N(0,1.5), +-4); ragged unless a shape is fixed"""
        rng = self._rng
        if self._synthetic_shape is not None:
            B, Te, L = self._synthetic_shape
            tl, ml = np.full(B, Te), np.full(B, L)
        else:
            B = hp.Train.Batch_Size
            tl, ml = rng.integers(16, 129, size=B), rng.integers(200, 801, size=B)
        S, E = self.metadata_Dict['Token_Index_Dict']['<S>'], self.metadata_Dict['Token_Index_Dict']['<E>']
        token_List = [np.hstack([S, rng.integers(2, hp.Encoder.Embedding.Token_Size, size=int(n) - 2), E]).astype(np.int32)
                      for n in tl]
        mel_List = [np.clip(rng.standard_normal((int(n), hp.Sound.Mel_Dim)) * 1.5, -hp.Sound.Max_Abs_Mel,
                            hp.Sound.Max_Abs_Mel).astype(np.float32) for n in ml]
        return self._collate(token_List, mel_List)

    def Get_Train_Pattern(self, is_Pre_Train=False):
        """Feeder.py:176-184"""
        if self.synthetic:
            return self._synthetic_pattern()
        queue = self.pre_Pattern_Queue if is_Pre_Train else self.pattern_Queue
        while len(queue) == 0:
            time.sleep(0.01)
        return queue.popleft()

    def _speaker_mel(self, path):
        """mel of a speaker wav for the embedding network (Feeder.py:211-223).  Reads 16-bit PCM wav files with the
        standard library; the features run on the GPU (Audio.melspectrogram).  A missing file (the reference would
        raise) degrades to a deterministic pseudo-speaker seeded by the path, so inference stays runnable offline."""
        from . import Audio
        if os.path.exists(path):
            with wave.open(path, 'rb') as f:
                sr, n, ch = f.getframerate(), f.getnframes(), f.getnchannels()
                x = np.frombuffer(f.readframes(n), dtype=np.int16).astype(np.float32) / 32768.0
            if ch > 1:
                x = x.reshape(-1, ch).mean(axis=1)
            if sr != hp.Sound.Sample_Rate:
                from scipy.signal import resample_poly
                g = np.gcd(sr, hp.Sound.Sample_Rate)
                x = resample_poly(x, hp.Sound.Sample_Rate // g, sr // g).astype(np.float32)
            x = _trim(x, top_db=15, frame_length=32, hop_length=16) * 0.99
            mel = Audio.melspectrogram(y=x, num_freq=hp.Sound.Spectrogram_Dim, frame_shift_ms=hp.Sound.Frame_Shift,
                                       frame_length_ms=hp.Sound.Frame_Length, num_mels=hp.Sound.Mel_Dim,
                                       sample_rate=hp.Sound.Sample_Rate, max_abs_value=hp.Sound.Max_Abs_Mel)
            return np.transpose(np.asarray(mel.cpu() if hasattr(mel, 'cpu') else mel)).astype(np.float32)
        rng = np.random.default_rng(zlib.crc32(path.encode()))
        return np.clip(rng.standard_normal((400, hp.Sound.Mel_Dim)) * 1.5, -4, 4).astype(np.float32)

    def Get_Inference_Pattern(self, speaker_Wav_Path_List, text_List):
        """Feeder.py:186-232: Mel = zeros [B,1,80], Mel_Length = 0"""
        tid = self.metadata_Dict['Token_Index_Dict']
        token_List = [np.array([tid['<S>']] + [tid[c] for c in text.upper()] + [tid['<E>']]).astype(np.int32)
                      for text in text_List]
        n = len(text_List)
        tok = np.zeros((n, max(t.shape[0] for t in token_List)), dtype=np.int32) + tid['<E>']
        for i, t in enumerate(token_List):
            tok[i, :t.shape[0]] = t
        p = self.placeholder_Dict
        return {
            p["Is_Training"]: False,
            p["Token"]: tok,
            p["Token_Length"]: np.array([t.shape[0] for t in token_List]).astype(np.int32),
            p["Mel"]: np.zeros((n, 1, hp.Sound.Mel_Dim), dtype=np.float32),
            p["Mel_Length"]: np.array([0 for _ in text_List]).astype(np.int32),
            p["Speaker_Embedding_Mel"]: self.Speaker_Embedding_Mel([self._speaker_mel(x) for x in speaker_Wav_Path_List]),
        }


def _trim(y, top_db=15, frame_length=32, hop_length=16):
    """librosa.effects.trim restated: frames (centred, reflect-padded) whose RMS is within top_db of the maximum"""
    if y.shape[0] < frame_length:
        return y
    yp = np.pad(y, frame_length // 2, mode='reflect')
    nfr = 1 + (yp.shape[0] - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(nfr)[:, None]
    rms = np.sqrt(np.mean(yp[idx] ** 2, axis=1))
    db = 20.0 * np.log10(np.maximum(1e-5 * 1.0, rms)) - 20.0 * np.log10(max(1e-5, rms.max()))  # power_to_db(ref=max)
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return y[:0]
    start, end = int(nz[0]) * hop_length, min(y.shape[0], (int(nz[-1]) + 1) * hop_length)
    return y[start:end]
