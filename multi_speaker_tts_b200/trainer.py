"""Decoder train step on one GPU rank: masks -> forward loop -> loss -> reverse loop -> (allreduce) -> TF Adam.

This is the data-parallel hot path of MSTTS_SV.py:127-192 restricted to the decoder variables: one flat fp32
parameter buffer, one flat gradient buffer (a single NCCL all-reduce per step when world_size > 1), TF-style
Adam with the reference's learning-rate schedule and weight-regularisation set.
"""
import math

import torch

from . import synthetic
from . import Hyper_Parameters as hp
from .decoder import decoder_forward, decoder_backward, decoder_loss, fill_mask, adam_tf

# Variables inside the reference's regularised set (name filter at MSTTS_SV.py:145-159): no 'bias', 'lstm',
# 'rnn', 'weight_w', 'projection' in the TF variable name.
L2_KEYS = ['prenet_0/kernel', 'prenet_1/kernel', 'memory_layer/kernel', 'query_layer/kernel',
           'location/conv1d/kernel', 'location/dense/kernel']


def learning_rate(global_step):
    lr_hp = hp.Train.Learning_Rate
    lr = lr_hp.Initial * lr_hp.Decay_Rate ** ((global_step - lr_hp.Decay_Start_Step) / lr_hp.Decay_Step)
    return min(max(lr, lr_hp.Min), lr_hp.Initial)


class DecoderTrainer(object):
    def __init__(self, device, mem_dim=768, mode="fp32", seed=0, process_group=None, weights=None):
        self.device = device
        self.mode = mode
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        shapes = synthetic.decoder_weight_shapes(mem_dim)
        order = L2_KEYS + [k for k in shapes if k not in L2_KEYS]
        offs, off = {}, 0
        for k in order:
            n = 1
            for d in shapes[k]:
                n *= d
            offs[k] = (off, n)
            off += (n + 3) // 4 * 4  # keep every tensor 16-byte aligned
            if k == L2_KEYS[-1]:
                self.n_l2 = off
        self.n_total = off
        self.flat_p = torch.zeros(off, device=device)
        self.flat_g = torch.zeros(off, device=device)
        self.flat_m = torch.zeros(off, device=device)
        self.flat_v = torch.zeros(off, device=device)
        self.w = {k: self.flat_p[o:o + n].view(shapes[k]) for k, (o, n) in offs.items()}
        self.g = {k: self.flat_g[o:o + n].view(shapes[k]) for k, (o, n) in offs.items()}
        init = weights if weights is not None else synthetic.init_decoder_weights(seed, mem_dim)
        for k, v in init.items():
            self.w[k].copy_(v)
        self.global_step = 0
        self.seed = seed
        self.workspace = None
        self._masks = {}

    def _mask_buffers(self, T, B):
        key = (T, B)
        if key not in self._masks:
            self._masks = {key: (torch.empty(T, 2, B, 256, device=self.device, dtype=torch.uint8),
                                 torch.empty(T, 2, 2, B, 1024, device=self.device, dtype=torch.uint8))}
        return self._masks[key]

    def train_step(self, memory, text_len, mel, mel_len, n_steps, masks=None):
        """One optimiser step.  Returns the device tensor [linear_loss, stop_loss] (no host sync)."""
        B = memory.shape[0]
        if masks is None:
            pm, zm = self._mask_buffers(n_steps, B)
            base = (self.seed * 1000003 + self.global_step) * 4
            fill_mask(pm, 1.0 - hp.Decoder.PreNet.Dropout_Rate, base + 1)
            fill_mask(zm, 1.0 - hp.Decoder.LSTM.Zoneout_Rate, base + 2)
        else:
            pm, zm = masks
        lin, stop, align, st = decoder_forward(self.w, memory, text_len, mel, mel_len, pm, zm, True, n_steps, self.mode,
                                               workspace=self.workspace)
        self.workspace = st.ws
        loss2, dlin, dstop = decoder_loss(lin, stop, mel, mel_len, hp.Train.Use_L1_Loss)
        _, self.d_memory = decoder_backward(st, self.w, dlin, dstop, grad_out=self.g)
        if self.world > 1:
            ev = getattr(self, 'allreduce_events', None)
            if ev is not None:  # bench: per-rank wait + wire time of the collective (CUDA events, no synchronisation)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            torch.distributed.all_reduce(self.flat_g, group=self.pg)  # the single gradient all-reduce
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        self.global_step += 1
        t = self.global_step
        a = hp.Train.ADAM
        lr_t = learning_rate(t - 1) * math.sqrt(1.0 - a.Beta2 ** t) / (1.0 - a.Beta1 ** t)
        gs = 1.0 / self.world
        nl = self.n_l2
        adam_tf(self.flat_p[:nl], self.flat_m[:nl], self.flat_v[:nl], self.flat_g[:nl], lr_t, a.Beta1, a.Beta2, a.Epsilon,
                gs, hp.Train.Weight_Regularization_Rate)
        adam_tf(self.flat_p[nl:], self.flat_m[nl:], self.flat_v[nl:], self.flat_g[nl:], lr_t, a.Beta1, a.Beta2, a.Epsilon,
                gs, 0.0)
        return loss2

    def weight_regularization_loss(self):
        return hp.Train.Weight_Regularization_Rate * 0.5 * self.flat_p[:self.n_l2].square().sum()
