"""WaveGlow.WaveGlow surface of the reference (WaveGlow/WaveGlow.py:17-175) without TensorFlow: the vocoder trainer.

``Run_Train_Step(feed_dict)`` stands in for ``session.run(train_Tensor_Dict, feed_dict)``: Restructure_Train_Data ->
training-direction flows with saved activations -> Glow_Loss -> reverse pass -> (one all-reduce of the flat gradient
buffer) -> tf.clip_by_global_norm(., 0.1) -> TF Adam (epsilon 1e-8) with the exponential-decay learning rate
(WaveGlow/WaveGlow.py:54-74).  All compute goes through ``libmstts_b200.so``.
"""
import math
import os
import time

import numpy as np
import torch

from .. import Hyper_Parameters as hp
from .. import checkpoint
from ..Feeder import Placeholder
from ..decoder import adam_tf
from . import Modules

CLIP_NORM = 0.1  # WaveGlow/WaveGlow.py:70


class Feeder(object):
    """WaveGlow/Feeder.py surface: placeholders 'Audio' [N,S] and 'Mel' [N,Tm,80].  No dataset exists offline, so training
    patterns are seeded synthetic audio with the mel computed by the GPU feature kernel (the reference's own pairing:
    Audio.melspectrogram of the signal, WaveGlow/Feeder.py:81-93) or, with ``random_mel``, random mels of the right length."""

    def __init__(self, seed=1234, rank=0, batch_size=None, signal_length=None, random_mel=True):
        self.placeholder_Dict = {'Mel': Placeholder('mel_placeholder', np.float32, (None, None, hp.Sound.Mel_Dim)),
                                 'Audio': Placeholder('audio_placeholder', np.float32, (None, None))}
        self._rng = np.random.default_rng(seed + rank)
        self.batch_size = batch_size or hp.WaveGlow.Train.Batch_Size
        self.signal_length = signal_length or hp.WaveGlow.Train.Max_Signal_Length
        self.random_mel = random_mel

    def Get_Train_Pattern(self):
        N, S = self.batch_size, self.signal_length
        audio = np.clip(self._rng.standard_normal((N, S)) * 0.3, -0.99, 0.99).astype(np.float32)
        # enough mel frames for the crop in Restructure_Train_Data: (Tm - 1) * 256 + 1024 >= S
        Tm = max(1, math.ceil((S - hp.WaveGlow.Upsample.Kernel_Size) / hp.WaveGlow.Upsample.Strides) + 1)
        mel = np.clip(self._rng.standard_normal((N, Tm, hp.Sound.Mel_Dim)) * 1.5, -4, 4).astype(np.float32)
        return {self.placeholder_Dict['Audio']: audio, self.placeholder_Dict['Mel']: mel}


def learning_rate(global_step):
    """WaveGlow/WaveGlow.py:54-60: exponential decay from step 0, floored at Min (no upper clip)"""
    lr = hp.WaveGlow.Train.Learning_Rate
    return max(lr.Min, lr.Initial * lr.Decay_Rate ** (global_step / lr.Decay_Step))


class WaveGlow(object):
    def __init__(self, device=None, seed=0, process_group=None, feeder=None, raws=None, up_kernel=None, up_bias=None):
        if not torch.cuda.is_available():
            raise RuntimeError("multi_speaker_tts_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.feeder = feeder if feeder is not None else Feeder()
        self.Tensor_Generate(seed, raws, up_kernel, up_bias)

    def Tensor_Generate(self, seed, raws, up_kernel, up_bias):
        """every variable (weight-norm g / v / bias, end convs, invertible 1x1 kernels, upsampling kernel) as a view of one flat
        fp32 buffer, with matching flat gradient / Adam-moment buffers"""
        if raws is None:
            raws, up_kernel, up_bias = _reference_init(seed)
        tensors = []

        def walk(r):
            for grp in ('start',):
                for k in ('g', 'v', 'b'):
                    tensors.append(r[grp][k])
            for grp in ('in', 'cond', 'res'):
                for x in r[grp]:
                    for k in ('g', 'v', 'b'):
                        tensors.append(x[k])
            tensors.extend([r['end_w'], r['end_b'], r['inv_w']])
        for r in raws:
            walk(r)
        tensors.extend([up_kernel, up_bias])
        offs, off = [], 0
        for t in tensors:
            offs.append(off)
            off += (t.numel() + 3) // 4 * 4
        dev = self.device
        self.flat_p = torch.zeros(off, device=dev)
        self.flat_g = torch.zeros(off, device=dev)
        self.flat_m = torch.zeros(off, device=dev)
        self.flat_v = torch.zeros(off, device=dev)
        it = iter(zip(tensors, offs))

        def views(flat, fill):
            out = []
            local = iter(zip(tensors, offs))

            def nxt():
                t, o = next(local)
                v = flat[o:o + t.numel()].view(t.shape)
                if fill:
                    v.copy_(t)
                return v
            for _ in raws:
                d = {'start': {k: nxt() for k in ('g', 'v', 'b')}}
                for grp in ('in', 'cond', 'res'):
                    d[grp] = [{k: nxt() for k in ('g', 'v', 'b')} for _ in range(Modules.LAYERS)]
                d['end_w'], d['end_b'], d['inv_w'] = nxt(), nxt(), nxt()
                out.append(d)
            return out, nxt(), nxt()
        del it
        praws, pk, pb = views(self.flat_p, True)
        graws, gk, gb = views(self.flat_g, False)
        self.params = Modules.WaveGlowParams.__new__(Modules.WaveGlowParams)
        self.params.raws, self.params.up_kernel, self.params.up_bias, self.params.device = praws, pk, pb, dev
        self.grads, self.g_up_kernel, self.g_up_bias = graws, gk, gb
        self.n_params = sum(t.numel() for t in tensors)
        self.global_Step = 0
        self.train_Tensor_Dict = {k: k for k in ['Global_Step', 'Learning_Rate', 'Log_S_Loss', 'Log_Det_W_Loss', 'Audio_Loss', 'Train_OP']}
        self.inference_Tensor_Dict = {k: k for k in ['Global_Step', 'Audio']}

    def Run_Train_Step(self, feed_dict):
        p = self.feeder.placeholder_Dict
        dev = self.device

        def up(a):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            return t if t.is_cuda else t.pin_memory().to(dev, non_blocking=True)
        audio, mel = up(feed_dict[p['Audio']]), up(feed_dict[p['Mel']])
        a, m = Modules.Restructure_Train_Data(audio, mel, self.params)
        z, losses, _, d_mel = Modules.Glow_Train_Backward(a, m, self.params, grads=self.grads)
        N, T, _ = a.shape
        dk, db = Modules.Upsample_Mel_Backward(mel, d_mel.reshape(N, T * hp.WaveGlow.Groups, hp.Sound.Mel_Dim), self.params)
        self.g_up_kernel.copy_(dk)
        self.g_up_bias.copy_(db)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)  # the single gradient all-reduce of the step
        # tf.clip_by_global_norm(gradients, 0.1) on the (mean) gradient; one host read per step (losses travel with it)
        stats = torch.stack([torch.linalg.vector_norm(self.flat_g, dtype=torch.float64) / self.world] + [x.double() for x in losses]).cpu().tolist()
        gnorm = stats[0]
        scale = (CLIP_NORM / max(gnorm, CLIP_NORM)) / self.world
        step = self.global_Step
        lr = learning_rate(step)
        t = step + 1
        ad = hp.WaveGlow.Train.ADAM
        lr_t = lr * math.sqrt(1.0 - ad.Beta2 ** t) / (1.0 - ad.Beta1 ** t)
        adam_tf(self.flat_p, self.flat_m, self.flat_v, self.flat_g, lr_t, ad.Beta1, ad.Beta2, ad.Epsilon, scale, 0.0)
        self.global_Step += 1
        return {'Global_Step': step, 'Learning_Rate': lr, 'Log_S_Loss': stats[1], 'Log_Det_W_Loss': stats[2], 'Audio_Loss': stats[3],
                'Global_Norm': gnorm, 'Train_OP': None}

    def Train(self, max_Steps=None):
        """WaveGlow/WaveGlow.py:110-140 (the reference loops forever; the periodic wav export is outside this build)"""
        done = 0
        while max_Steps is None or done < max_Steps:
            start_Time = time.time()
            r = self.Run_Train_Step(self.feeder.Get_Train_Pattern())
            print('\t\t'.join(['Time: {:0.3f}'.format(time.time() - start_Time), 'Global step: {}'.format(r['Global_Step']),
                               'Learning rate: {:0.5f}'.format(r['Learning_Rate']), 'Log S Loss: {:0.5f}'.format(r['Log_S_Loss']),
                               'Log Det W Loss: {:0.5f}'.format(r['Log_Det_W_Loss']), 'Audio Loss: {:0.5f}'.format(r['Audio_Loss'])]))
            if (r['Global_Step'] + 1) % hp.WaveGlow.Train.Checkpoint_Save_Timing == 0:
                self.Save()
            done += 1

    def Run_Inference(self, mels, sigma=1.0, generator=None):
        """session.run(inference_Tensor_Dict, {Mel: mels}): Restructure_Inference_Data -> Glow_Inference -> [N, samples]"""
        mels = mels if torch.is_tensor(mels) else torch.from_numpy(np.ascontiguousarray(mels))
        a, m = Modules.Restructure_Inference_Data(mels.to(self.device), self.params, generator=generator)
        return {'Global_Step': self.global_Step, 'Audio': Modules.Glow_Inference(a, m, self.params, sigma, generator=generator)}

    def Save(self):
        """tf.train.Saver(max_to_keep=5).save(..., 'CHECKPOINT', global_step): variables + the Adam slots (the reference's Saver
        checkpoints the optimizer's m / v with the variables); rank 0 writes, CHECKPOINT-<step>.pt, newest five kept"""
        cpu = lambda d: {k: (cpu(v) if isinstance(v, dict) else [cpu(x) for x in v] if isinstance(v, list) else v.detach().cpu())
                         for k, v in d.items()}
        blob = {'raws': [cpu(r) for r in self.params.raws], 'up_kernel': self.params.up_kernel.cpu(),
                'up_bias': self.params.up_bias.cpu(), 'global_step': self.global_Step,
                'flat_m': self.flat_m.cpu(), 'flat_v': self.flat_v.cpu()}
        return checkpoint.save(hp.WaveGlow.Checkpoint_Path, blob, self.global_Step, max_to_keep=5, process_group=self.pg)

    def Restore(self):
        path = checkpoint.latest_checkpoint(hp.WaveGlow.Checkpoint_Path)
        if path is None:
            print('There is no checkpoint.')
            return
        blob = torch.load(path, map_location='cpu')

        def put(dst, src):  # copy a saved tree into the views of the flat parameter buffer
            if isinstance(dst, dict):
                for k in dst:
                    put(dst[k], src[k])
            elif isinstance(dst, list):
                for d, x in zip(dst, src):
                    put(d, x)
            else:
                dst.copy_(src.to(self.device))
        put(self.params.raws, blob['raws'])
        self.params.up_kernel.copy_(blob['up_kernel'].to(self.device))
        self.params.up_bias.copy_(blob['up_bias'].to(self.device))
        for name in ('flat_m', 'flat_v'):
            if name in blob and blob[name].numel() == getattr(self, name).numel():
                getattr(self, name).copy_(blob[name].to(self.device))
        self.global_Step = int(blob.get('global_step', 0))
        print('Checkpoint \'{}\' is loaded.'.format(path))


def _reference_init(seed):
    """variables as the reference creates them (SURVEY A-6): glorot g and v, zero end conv, N(0,1) 1x1 kernels with the
    determinant made positive, U(0, 0.02) upsampling kernel"""
    g = torch.Generator().manual_seed(seed)

    def glorot(shape):
        rf = 1
        for d in shape[:-2]:
            rf *= d
        lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).float()

    def wn(shape):
        lim = math.sqrt(6.0 / (2 * shape[-1]))
        return {'g': ((torch.rand(shape[-1], generator=g, dtype=torch.float64) * 2 - 1) * lim).float(), 'v': glorot(shape),
                'b': torch.zeros(shape[-1])}
    raws = []
    C, L, mel = hp.WaveGlow.WaveNet.Channels, hp.WaveGlow.WaveNet.Layers, hp.WaveGlow.Groups * hp.Sound.Mel_Dim
    for f in range(hp.WaveGlow.Flows):
        c = Modules.flow_channels(f)
        r = {'start': wn((1, c // 2, C)), 'in': [wn((hp.WaveGlow.WaveNet.Kernel_Size, C, 2 * C)) for _ in range(L)],
             'cond': [wn((1, mel, 2 * C)) for _ in range(L)],
             'res': [wn((1, C, 2 * C if i < L - 1 else C)) for i in range(L)],
             'end_w': torch.zeros(1, C, c), 'end_b': torch.zeros(c)}
        W = torch.randn((c, c), generator=g, dtype=torch.float64)
        if torch.linalg.det(W) < 0:
            W[:, 0] *= -1
        r['inv_w'] = W.float()
        raws.append(r)
    up_k = (torch.rand((hp.WaveGlow.Upsample.Kernel_Size, hp.Sound.Mel_Dim, hp.Sound.Mel_Dim), generator=g, dtype=torch.float64) * 0.02).float()
    return raws, up_k, torch.zeros(hp.Sound.Mel_Dim)
