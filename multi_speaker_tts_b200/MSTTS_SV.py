"""MSTTS_SV.Tacotron2 surface of the reference (MSTTS_SV.py:20-520) without TensorFlow.

``Tacotron2(is_Training)`` keeps the reference's attributes and methods (``feeder``, ``train_Tensor_Dict`` /
``inference_Tensor_Dict`` key sets, ``Restore``, ``Train``, ``Inference``); a ``tf.Session.run(fetches, feed_dict)`` is
replaced by ``Run_Train_Step(feed_dict)`` / ``Run_Inference(feed_dict)`` which take the Feeder's feed dicts unchanged and
return the same result dicts.  The decoder loop -- the hot path -- runs in the persistent sm_100a kernels
(``Modules.Decoder_LSTM``); encoder, postnet and the frozen speaker-embedding net are library ops.

Differences the survey asks for (8b "constructor side effects"): a missing speaker-embedding / vocoder checkpoint does not
raise -- the frozen sub-nets keep their seeded random initialisation and a notice is printed; ``Train`` skips the
periodic inference + export when ``Inference_Sentence_in_Train.txt`` is absent and accepts ``max_Steps``.
Training is data parallel when a process group is given: one flat fp32 gradient buffer, one all-reduce per step.
"""
import math
import os
import time

import numpy as np
import torch

from . import Hyper_Parameters as hp
from . import Feeder, Modules, checkpoint
from .Location_Sensitive_Attention import Location_Sensitive_Attention
from .Speaker_Embedding import Modules as Speaker_Embedding_Modules
from .decoder import adam_tf
from .synthetic import TF_VARIABLE_NAMES, decoder_weight_shapes

_FROZEN_SCOPES = ('speaker_embedding', 'mel_to_spectrogram', 'waveglow')
_WR_EXCLUDE = ('bias', 'embedding', 'lstm', 'rnn', 'weight_w', 'projection')


def variable_shapes():
    """TF variable name -> (shape, kind) for every Tacotron2-side variable (SURVEY A-8).
    kind: 'kernel' (glorot-uniform), 'zeros', 'ones' (trainable) or 'moving_mean' / 'moving_variance' (not trainable)."""
    e, d = hp.Encoder, hp.Decoder
    emb = hp.Speaker_Embedding.Embedding_Size
    out = {'encoder/embedding_variable': ((e.Embedding.Token_Size, e.Embedding.Embedding_Size), 'kernel')}

    def conv_stack(scope, n, k, cin, couts):
        for i in range(n):
            p = '%s/conv_%d' % (scope, i)
            out[p + '/conv1d/kernel'] = ((k, cin, couts[i]), 'kernel')
            out[p + '/conv1d/bias'] = ((couts[i],), 'zeros')
            out[p + '/batch_normalization/gamma'] = ((couts[i],), 'ones')
            out[p + '/batch_normalization/beta'] = ((couts[i],), 'zeros')
            out[p + '/batch_normalization/moving_mean'] = ((couts[i],), 'moving_mean')
            out[p + '/batch_normalization/moving_variance'] = ((couts[i],), 'moving_variance')
            cin = couts[i]

    conv_stack('encoder', e.Conv.Nums, e.Conv.Kernel_Size, e.Embedding.Embedding_Size, [e.Conv.Channel] * e.Conv.Nums)
    cin = e.Conv.Channel
    for n in range(e.BiLSTM.Nums):
        for dr in ('fw', 'bw'):
            p = 'encoder/bilstm/stack_bidirectional_rnn/cell_%d/bidirectional_rnn/%s/zoneout_lstm_cell' % (n, dr)
            out[p + '/kernel'] = ((cin + e.BiLSTM.Cell_Size, 4 * e.BiLSTM.Cell_Size), 'kernel')
            out[p + '/bias'] = ((4 * e.BiLSTM.Cell_Size,), 'zeros')
        cin = 2 * e.BiLSTM.Cell_Size
    mem_dim = 2 * e.BiLSTM.Cell_Size + emb
    for short, shape in decoder_weight_shapes(mem_dim).items():
        out[TF_VARIABLE_NAMES[short]] = (shape, 'zeros' if short.endswith('bias') or short.endswith('bias_b') else 'kernel')
    conv_stack('decoder', d.Conv.Nums, d.Conv.Kernel_Size, hp.Sound.Mel_Dim,
               [d.Conv.Channel] * (d.Conv.Nums - 1) + [hp.Sound.Mel_Dim])
    # frozen speaker-embedding net (Speaker_Embedding/Modules.py:6-37)
    s = hp.Speaker_Embedding
    out['speaker_embedding/dense/kernel'] = ((hp.Sound.Mel_Dim, emb), 'kernel')
    out['speaker_embedding/dense/bias'] = ((emb,), 'zeros')
    for i in range(s.LSTM.Nums):
        p = 'speaker_embedding/lstm/rnn/multi_rnn_cell/cell_%d/lstmcell_%d' % (i, i)
        out[p + '/kernel'] = ((emb + s.LSTM.Cell_Size, 4 * s.LSTM.Cell_Size), 'kernel')
        out[p + '/bias'] = ((4 * s.LSTM.Cell_Size,), 'zeros')
    return out


def is_trainable(name, kind):
    return not kind.startswith('moving') and not name.startswith(_FROZEN_SCOPES)


def in_weight_regularization(name):
    """the name filter of MSTTS_SV.py:145-159"""
    low = name.lower()
    return not any(x in low for x in _WR_EXCLUDE) and not name.startswith(_FROZEN_SCOPES)


def _init(shape, kind, gen):
    if kind in ('zeros', 'moving_mean'):
        return torch.zeros(shape)
    if kind in ('ones', 'moving_variance'):
        return torch.ones(shape)
    if len(shape) == 1:
        fi = fo = shape[0]
    else:
        rf = 1
        for x in shape[:-2]:
            rf *= x
        fi, fo = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(6.0 / (fi + fo))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).float()


class _FusedLosses(torch.autograd.Function):
    """linear / postnet / stop losses of MSTTS_SV.py:127-144 with their gradients from mstts_decoder_loss (csrc/capi.cu): the
    kernel computes mean((x - mel)^2) (+ mean|x - mel|) over [B, L, 80] and the stop-token BCE over [B, T] together with d x and
    d stop; it runs once on the linear output (+ stop) and once on the postnet output"""

    @staticmethod
    def forward(ctx, linear, stop, post, mel, mel_len, use_l1):
        from .decoder import decoder_loss
        l_a, d_lin, d_stop = decoder_loss(linear.contiguous(), stop.contiguous(), mel, mel_len, use_l1)
        l_b, d_post, _ = decoder_loss(post.contiguous(), stop.contiguous(), mel, mel_len, use_l1)
        ctx.save_for_backward(d_lin, d_stop, d_post)
        return l_a[0].clone(), l_b[0].clone(), l_a[1].clone()

    @staticmethod
    def backward(ctx, g_lin, g_post, g_stop):
        d_lin, d_stop, d_post = ctx.saved_tensors
        return d_lin * g_lin, d_stop * g_stop, d_post * g_post, None, None, None


class Tacotron2(object):
    def __init__(self, is_Training=False, device=None, seed=0, process_group=None, feeder=None, mode=None):
        if not torch.cuda.is_available():
            raise RuntimeError("multi_speaker_tts_b200 needs a CUDA device (no CPU fallback)")
        self.is_Training = is_Training
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.mode = mode
        self.seed = seed
        self.feeder = feeder if feeder is not None else Feeder.Feeder(is_Training=is_Training)
        self.Tensor_Generate()
        self.Speaker_Embedding_Load()
        self.Vocoder_Load()

    # ---- variables ------------------------------------------------------------------------------------------------
    def Tensor_Generate(self):
        """Variables in one flat fp32 buffer (regularised set first, then the other trainables), matching flat gradient
        and Adam-moment buffers; non-trainable variables (moving statistics, frozen sub-nets) beside it."""
        gen = torch.Generator().manual_seed(self.seed)
        shapes = variable_shapes()
        init = {k: _init(s, kind, gen) for k, (s, kind) in shapes.items()}
        train = [k for k, (s, kind) in shapes.items() if is_trainable(k, kind)]
        order = [k for k in train if in_weight_regularization(k)] + [k for k in train if not in_weight_regularization(k)]
        offs, off = {}, 0
        for k in order:
            n = init[k].numel()
            offs[k] = (off, n)
            off += (n + 3) // 4 * 4
            if in_weight_regularization(k):
                self.n_l2 = off
        dev = self.device
        self.flat_p = torch.zeros(off, device=dev)
        self.flat_g = torch.zeros(off, device=dev)
        self.flat_m = torch.zeros(off, device=dev)
        self.flat_v = torch.zeros(off, device=dev)
        self.variables = {}
        self._grad_views = {}
        for k, (o, n) in offs.items():
            self.variables[k] = self.flat_p[o:o + n].view(shapes[k][0])
            self.variables[k].copy_(init[k])
            self._grad_views[k] = self.flat_g[o:o + n].view(shapes[k][0])
        for k in shapes:
            if k not in offs:
                self.variables[k] = init[k].to(dev)
        self.trainable = order
        self.global_Step = 0
        self._short = {v: k for k, v in TF_VARIABLE_NAMES.items()}
        keys = ['Global_Step', 'Learning_Rate', 'Loss', 'Linear_Loss', 'Postnet_Loss', 'Stop_Loss',
                'Weight_Regularization_Loss', 'Train_OP']
        self.train_Tensor_Dict = {k: k for k in keys} if self.is_Training else None
        self.inference_Tensor_Dict = {k: k for k in ['Global_Step', 'Linear', 'Mel', 'Stop', 'Attention_History']}
        self.waveglow_params = None

    def _decoder_variables(self, v):
        return {short: v[name] for short, name in TF_VARIABLE_NAMES.items()}

    def Speaker_Embedding_Load(self):
        path = checkpoint.latest_checkpoint(hp.Speaker_Embedding.Checkpoint_Path)
        if path is not None:
            self._load(path, lambda k: k.startswith('speaker_embedding'))
            print('Speaker embedding checkpoint \'{}\' is loaded.'.format(path))
        else:
            print('There is no speaker embedding checkpoint: the frozen speaker network keeps its seeded random initialisation.')

    def Vocoder_Load(self):
        if hp.Use_Vocoder.upper() != 'WaveGlow'.upper():
            print('Vocoder \'{}\' is not part of this build; Inference returns mels without waveforms.'.format(hp.Use_Vocoder))
            return
        path = checkpoint.latest_checkpoint(hp.WaveGlow.Checkpoint_Path)
        if path is not None:
            from .WaveGlow import Modules as WaveGlow_Modules
            blob = torch.load(path, map_location='cpu')
            self.waveglow_params = WaveGlow_Modules.WaveGlowParams(blob['raws'], blob['up_kernel'], blob['up_bias'], self.device)
            print('Vocoder checkpoint \'{}\' is loaded.'.format(path))
        else:
            print('There is no vocoder checkpoint: Inference returns mels without waveforms.')

    def _load(self, path, select=lambda k: True):
        blob = torch.load(path, map_location='cpu')
        for k, t in blob['variables'].items():
            if k in self.variables and select(k):
                self.variables[k].copy_(t.to(self.device))
        return blob

    def Restore(self):
        """MSTTS_SV.py:244-251: tf.train.latest_checkpoint(hp.Checkpoint_Path) -> Saver.restore"""
        path = checkpoint.latest_checkpoint(hp.Checkpoint_Path)
        if path is None:
            print('There is no checkpoint.')
            return
        blob = self._load(path, lambda k: not k.startswith(_FROZEN_SCOPES))
        self.global_Step = int(blob.get('global_step', 0))
        for name in ('flat_m', 'flat_v'):
            if name in blob:
                getattr(self, name).copy_(blob[name].to(self.device))
        print('Checkpoint \'{}\' is loaded.'.format(path))

    def Save(self):
        """MSTTS_SV.py:30-40,287-289: Saver(max_to_keep=5).save(..., 'CHECKPOINT', global_step) -> CHECKPOINT-<step> (+ the
        `checkpoint` state file); variables keyed by their TF names (SURVEY A-8) + the Adam slots, like tf.train.Saver.
        Under data parallel the batch-norm moving statistics (updated from per-rank batches) are averaged first so that
        every rank would write the same file, and rank 0 writes it."""
        if self.world > 1:
            for k in sorted(self.variables):
                if k.endswith(('/moving_mean', '/moving_variance')) and not k.startswith(_FROZEN_SCOPES):
                    torch.distributed.all_reduce(self.variables[k], group=self.pg)
                    self.variables[k].div_(self.world)
        blob = {'variables': {k: v.detach().cpu() for k, v in self.variables.items()}, 'global_step': self.global_Step,
                'flat_m': self.flat_m.cpu(), 'flat_v': self.flat_v.cpu()}
        return checkpoint.save(hp.Checkpoint_Path, blob, self.global_Step, max_to_keep=5, process_group=self.pg)

    # ---- graph ----------------------------------------------------------------------------------------------------
    def _to_device(self, feed_dict):
        p = self.feeder.placeholder_Dict
        out = {}
        for key in ('Token', 'Token_Length', 'Mel', 'Mel_Length', 'Speaker_Embedding_Mel'):
            a = feed_dict[p[key]]
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            if not t.is_cuda:
                if key == 'Mel_Length':  # known on the host: the decoder's step count needs no device read-back
                    out['Mel_Length_Max'] = int(t.max())
                if not t.is_pinned():  # the Feeder hands out page-locked arrays; anything else is staged here
                    t = t.pin_memory()
                t = t.to(self.device, non_blocking=True)
            out[key] = t
        out['Is_Training'] = bool(feed_dict[p['Is_Training']])
        return out

    def _forward(self, v, feed, masks=None):
        """MSTTS_SV.py:45-98.  v: name -> tensor (leaves for autograd in training)."""
        training = feed['Is_Training']
        masks = masks or {}
        branch = Modules.side_stream(feed['Token'].device, 'speaker')
        with branch, torch.no_grad():  # frozen (MSTTS_SV.py:183-190); independent of the encoder: runs beside it
            e = Speaker_Embedding_Modules.Restructure(feed['Speaker_Embedding_Mel'], self.variables)
            lengths = torch.full((e.shape[0],), hp.Speaker_Embedding.Inference.Mel_Frame, device=e.device)
            e = Speaker_Embedding_Modules.Stack_LSTM(e, lengths, training, self.variables)
            e = Speaker_Embedding_Modules.Inference(e)                               # [B, 256]
        x = Modules.Encoder_Embedding(feed['Token'], v)
        x = Modules.Encoder_Conv(x, training, v, masks.get('encoder_conv'))
        x = Modules.Encoder_BiLSTM(x, feed['Token_Length'], training, v, masks.get('encoder_bilstm'))
        e = branch.join(e)
        memory = torch.cat([x, e[:, None, :].expand(-1, x.shape[1], -1)], dim=-1).contiguous()
        dv = self._decoder_variables(v)
        att = Location_Sensitive_Attention(
            num_units=hp.Attention.Memory_Size, memory=memory, memory_length=feed['Token_Length'],
            conv_kernel_size=hp.Attention.Conv.Kernel_Size, conv_stride_size=hp.Attention.Conv.Stride,
            conv_channel=hp.Attention.Conv.Channel, dropout_rate=hp.Attention.Conv.Dropout_Rate, is_training=training,
            variables=dv)
        out, state = Modules.Decoder_LSTM(feed['Mel'], feed['Mel_Length'], att, training, variables=dv,
                                          masks=masks.get('decoder'), mode=self.mode,
                                          seed=self.seed * 1000003 + self.global_Step,
                                          max_length=feed.get('Mel_Length_Max') if training else None)
        post = Modules.Decoder_Conv(out.linear, training, v, masks.get('postnet'))
        post = out.linear + post
        return out, post, state.alignment_history.stack().permute(1, 2, 0)            # Attention_History [B,Te,T]

    def _losses(self, v, out, post, feed):
        """MSTTS_SV.py:127-161"""
        mel, mel_len = feed['Mel'], feed['Mel_Length']
        T = out.linear.shape[1]
        if out.linear.is_cuda and out.linear.dtype == torch.float32 and mel.shape[1] == T - 1:
            # the three loss terms and their gradients in two launches of the library's loss kernel (the op-by-op form below
            # is ~100 tiny launches forward + backward); same unmasked means
            linear_Loss, postnet_Loss, stop_Loss = _FusedLosses.apply(out.linear, out.stop.squeeze(2), post, mel, mel_len,
                                                                      bool(hp.Train.Use_L1_Loss))
            wr = (0.5 * hp.Train.Weight_Regularization_Rate) * torch.linalg.vector_norm(self.flat_p[:self.n_l2]) ** 2
            return linear_Loss, postnet_Loss, stop_Loss, wr
        stop_target = (torch.arange(T, device=mel.device)[None, :] >= mel_len[:, None]).float()
        lin, pst = out.linear[:, :-1], post[:, :-1]
        linear_Loss = torch.mean((lin - mel) ** 2)
        postnet_Loss = torch.mean((pst - mel) ** 2)
        if hp.Train.Use_L1_Loss:
            linear_Loss = linear_Loss + torch.mean(torch.abs(lin - mel))
            postnet_Loss = postnet_Loss + torch.mean(torch.abs(pst - mel))
        stop_Loss = torch.nn.functional.binary_cross_entropy_with_logits(out.stop.squeeze(2), stop_target)
        # the regularised variables are the first n_l2 floats of the flat parameter buffer (padding is zero): one reduction
        # instead of two tiny kernels per variable.  Reported only -- its gradient is the l2 * p term inside the Adam kernel
        wr = (0.5 * hp.Train.Weight_Regularization_Rate) * torch.linalg.vector_norm(self.flat_p[:self.n_l2]) ** 2
        return linear_Loss, postnet_Loss, stop_Loss, wr

    def learning_rate(self, global_step):
        """MSTTS_SV.py:163-169"""
        lr = hp.Train.Learning_Rate
        x = lr.Initial * lr.Decay_Rate ** ((global_step - lr.Decay_Start_Step) / lr.Decay_Step)
        return min(max(x, lr.Min), lr.Initial)

    def Run_Train_Step(self, feed_dict, masks=None):
        """session.run(train_Tensor_Dict, feed_dict): forward, losses, backward, (all-reduce), TF Adam.  The weight
        regularisation enters the update as l2 * p inside the Adam kernel (its gradient), not through autograd."""
        feed = self._to_device(feed_dict)
        # autograd leaves over the flat parameter buffer: views that share its storage, so they stay current across the in-place
        # Adam updates and are built once (rebuilt if somebody rebinds an entry of self.variables)
        lc = getattr(self, '_leaf_cache', None)
        if lc is None or any(lc[k][0] is not self.variables[k] for k in self.trainable):
            lc = {k: (self.variables[k], self.variables[k].detach().requires_grad_(True)) for k in self.trainable}
            self._leaf_cache = lc
        v = dict(self.variables)
        for k in self.trainable:
            v[k] = lc[k][1]
        out, post, _ = self._forward(v, feed, masks)
        linear_Loss, postnet_Loss, stop_Loss, wr = self._losses(v, out, post, feed)
        leaves = [v[k] for k in self.trainable]
        grads = torch.autograd.grad(linear_Loss + postnet_Loss + stop_Loss, leaves, allow_unused=True)
        dst, src = [], []
        for k, g in zip(self.trainable, grads):
            if g is None:
                self._grad_views[k].zero_()
            else:
                dst.append(self._grad_views[k])
                src.append(g)
        torch._foreach_copy_(dst, src)  # into the flat gradient buffer: a few fused launches instead of one copy per variable
        if self.world > 1:
            ev = getattr(self, 'allreduce_events', None)
            if ev is not None:  # bench: per-rank wait + wire time of the collective (CUDA events, no synchronisation)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            torch.distributed.all_reduce(self.flat_g, group=self.pg)  # the single gradient all-reduce of the step
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        step = self.global_Step
        lr = self.learning_rate(step)
        t = step + 1
        a = hp.Train.ADAM
        lr_t = lr * math.sqrt(1.0 - a.Beta2 ** t) / (1.0 - a.Beta1 ** t)
        nl, gs = self.n_l2, 1.0 / self.world
        adam_tf(self.flat_p[:nl], self.flat_m[:nl], self.flat_v[:nl], self.flat_g[:nl], lr_t, a.Beta1, a.Beta2, a.Epsilon, gs,
                hp.Train.Weight_Regularization_Rate)
        adam_tf(self.flat_p[nl:], self.flat_m[nl:], self.flat_v[nl:], self.flat_g[nl:], lr_t, a.Beta1, a.Beta2, a.Epsilon, gs,
                0.0)
        self.global_Step += 1
        losses = torch.stack([linear_Loss, postnet_Loss, stop_Loss, wr]).detach().cpu().tolist()  # the step's only D2H read
        return {'Global_Step': step, 'Learning_Rate': lr, 'Loss': sum(losses), 'Linear_Loss': losses[0],
                'Postnet_Loss': losses[1], 'Stop_Loss': losses[2], 'Weight_Regularization_Loss': losses[3], 'Train_OP': None}

    def Run_Inference(self, feed_dict, masks=None):
        """session.run(inference_Tensor_Dict, feed_dict)"""
        feed = self._to_device(feed_dict)
        with torch.no_grad():
            out, post, att = self._forward(self.variables, feed, masks)
        return {'Global_Step': self.global_Step, 'Linear': out.linear.cpu().numpy(), 'Mel': post.cpu().numpy(),
                'Stop': torch.sigmoid(out.stop.squeeze(2)).cpu().numpy(), 'Attention_History': att.cpu().numpy()}

    # ---- loops ----------------------------------------------------------------------------------------------------
    def Train(self, max_Steps=None):
        """MSTTS_SV.py:253-293 (the reference loops forever; max_Steps bounds it)"""
        def Run_Inference():
            if not os.path.exists('Inference_Sentence_in_Train.txt'):
                return
            paths, sentences = [], []
            with open('Inference_Sentence_in_Train.txt', 'r') as f:
                for line in f.readlines():
                    p, s = line.strip().split('\t')
                    paths.append(p)
                    sentences.append(s)
            self.Inference(paths, sentences)

        Run_Inference()
        done = 0
        while max_Steps is None or done < max_Steps:
            start_Time = time.time()
            pre = hp.Train.Use_Pre_in_Main_Train and self.global_Step < hp.Train.Pre_Step
            result_Dict = self.Run_Train_Step(self.feeder.Get_Train_Pattern(is_Pre_Train=pre))
            print('\t\t'.join([
                'Time: {:0.3f}'.format(time.time() - start_Time),
                'Global step: {}'.format(result_Dict['Global_Step']),
                'Mode: {}'.format('Pre-train' if result_Dict['Global_Step'] < hp.Train.Pre_Step else 'Main'),
                'Learning rate: {:0.5f}'.format(result_Dict['Learning_Rate']),
                'Linear loss: {:0.5f}'.format(result_Dict['Linear_Loss']),
                'Postnet loss: {:0.5f}'.format(result_Dict['Postnet_Loss']),
                'Stop loss: {:0.5f}'.format(result_Dict['Stop_Loss']),
                'WR loss: {:0.5f}'.format(result_Dict['Weight_Regularization_Loss']),
            ]))
            if (result_Dict['Global_Step'] + 1) % hp.Train.Checkpoint_Save_Timing == 0:
                self.Save()
            if (result_Dict['Global_Step'] + 1) % hp.Train.Inference_Timing == 0:
                Run_Inference()
            done += 1

    def Inference(self, path_List, text_List, file_Prefix=None):
        """MSTTS_SV.py:295-389: free-running decode (+ WaveGlow vocoding in Mel_Split_Length chunks when a vocoder is
        loaded).  Returns the result dict; plot / wav export (matplotlib, librosa) is outside this build."""
        result_Dict = self.Run_Inference(self.feeder.Get_Inference_Pattern(path_List, text_List))
        if self.waveglow_params is not None:
            from .WaveGlow import Modules as WaveGlow_Modules
            L = hp.WaveGlow.Inference.Mel_Split_Length
            wavs = []
            for mel in result_Dict['Mel']:
                chunks = [mel[x:x + L] for x in range(0, mel.shape[0], L)]
                pat = np.zeros((len(chunks), max(c.shape[0] for c in chunks), hp.Sound.Mel_Dim), dtype=np.float32)
                for i, c in enumerate(chunks):
                    pat[i, :c.shape[0]] = c
                parts = []
                for s in range(0, len(chunks), hp.WaveGlow.Inference.Batch_Size):
                    m = torch.from_numpy(pat[s:s + hp.WaveGlow.Inference.Batch_Size]).to(self.device)
                    a, mm = WaveGlow_Modules.Restructure_Inference_Data(m, self.waveglow_params)
                    parts.append(WaveGlow_Modules.Glow_Inference(a, mm, self.waveglow_params).cpu().numpy())
                wavs.append(np.concatenate([p.reshape(-1) for p in parts]))
            result_Dict['Wav'] = wavs
        return result_Dict
