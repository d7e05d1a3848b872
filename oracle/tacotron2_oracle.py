"""CPU ORACLE (test infrastructure, not product code) -- full Tacotron2 graph of MSTTS_SV.py:45-161 around the decoder
oracle: frozen speaker-embedding net, encoder (embedding, 3x conv+ReLU+BN+dropout, zoneout BiLSTM), speaker concat,
decoder loop (decoder_oracle), postnet (5x conv+tanh+BN+dropout), losses.  Row-by-row / step-by-step loops, no batching
tricks, so it is an independent restatement of what ``multi_speaker_tts_b200.Modules`` computes with batched ops.
Only tests/, smoke() and the bench cpu_baseline legs may import this.

PARITY UNPINNED (see decoder_oracle.py): TF1 cannot run here and the reference has no fixtures.  [TF-internal] facts
used: SURVEY Appendix A-4 (tf.layers defaults, BN axis/momentum/eps, conv SAME cross-correlation), A-5 (losses).

Reference lines: MSTTS_SV.py:45-98 (wiring), :127-161 (losses); Modules.py:15-73 (encoder), :121-143 (postnet);
Speaker_Embedding/Modules.py:6-37,127-137; ZoneoutLSTMCell.py:228-271; Feeder.py:155 (token padding).
All randomness enters as explicit masks.
"""
import torch
import torch.nn.functional as F

from . import decoder_oracle as D

BN_EPS, BN_MOMENTUM = 1e-3, 0.99


def conv_bn_stack(x, v, scope, n, act, training, dropout_rate, masks):
    for i in range(n):
        p = '%s/conv_%d' % (scope, i)
        w = v[p + '/conv1d/kernel']                       # [k, in, out]
        k = w.shape[0]
        y = F.conv1d(F.pad(x.transpose(1, 2), (k // 2, k // 2)), w.permute(2, 1, 0), v[p + '/conv1d/bias'])
        y = act(y.transpose(1, 2))
        if training:
            flat = y.reshape(-1, y.shape[-1])
            mean = flat.mean(dim=0)
            var = ((flat - mean) ** 2).mean(dim=0)        # biased, over (B,T) incl. padding
        else:
            mean, var = v[p + '/batch_normalization/moving_mean'], v[p + '/batch_normalization/moving_variance']
        y = (y - mean) / torch.sqrt(var + BN_EPS) * v[p + '/batch_normalization/gamma'] + v[p + '/batch_normalization/beta']
        if training:
            y = y / (1.0 - dropout_rate) * masks[i]
        x = y
    return x


def dynamic_rnn(x, lengths, kernel, bias, training, masks, zoneout=0.1, residual=False):
    """tf.nn.dynamic_rnn(ZoneoutLSTMCell, sequence_length): per row, steps beyond the length emit zeros and keep state"""
    B, T, _ = x.shape
    H = kernel.shape[1] // 4
    out = torch.zeros(B, T, H, dtype=x.dtype)
    for b in range(B):
        c = torch.zeros(1, H, dtype=x.dtype)
        h = torch.zeros(1, H, dtype=x.dtype)
        for t in range(int(lengths[b])):
            mc = mh = None
            if training:
                mc, mh = masks[t, 0, b:b + 1].to(x.dtype), masks[t, 1, b:b + 1].to(x.dtype)
            m, c, h = D.lstm_cell(x[b:b + 1, t], c, h, kernel, bias, mc, mh, zoneout)
            out[b, t] = (m + x[b:b + 1, t] if residual else m)[0]
    return out


def reverse_rows(x, lengths):
    y = x.clone()
    for b in range(x.shape[0]):
        n = int(lengths[b])
        y[b, :n] = torch.flip(x[b, :n], dims=[0])
    return y


def speaker_embedding(v, se_mel, training=False, sample_nums=5):
    x = se_mel @ v['speaker_embedding/dense/kernel'] + v['speaker_embedding/dense/bias']
    lengths = [x.shape[1]] * x.shape[0]
    for i in range(3):
        p = 'speaker_embedding/lstm/rnn/multi_rnn_cell/cell_%d/lstmcell_%d' % (i, i)
        assert not training, "the speaker network is frozen; the oracle runs it without zoneout masks"
        x = dynamic_rnn(x, lengths, v[p + '/kernel'], v[p + '/bias'], False, None, residual=(i < 2))
    e = x[:, -1, :].reshape(x.shape[0] // sample_nums, sample_nums, -1).mean(dim=1)
    return e / torch.sqrt(torch.clamp((e * e).sum(), min=1e-12))     # l2_normalize over the whole tensor (quirk B-4)


def forward(v, feed, masks, training=True, max_steps=D.MAX_INFERENCE_LENGTH, speaker_training=False):
    """v: TF-name -> tensor; feed: Token, Token_Length, Mel, Mel_Length, Speaker_Embedding_Mel (torch CPU tensors);
    masks: encoder_conv [3][B,Te,512], encoder_bilstm [(fw, bw)] each [Te,2,B,256], decoder (prenet, zone),
    postnet [5][B,T,C].  Returns linear, stop, align [B,T,Te], postnet output, memory."""
    from multi_speaker_tts_b200.synthetic import TF_VARIABLE_NAMES
    e = speaker_embedding(v, feed['Speaker_Embedding_Mel'], speaker_training)
    x = v['encoder/embedding_variable'][feed['Token'].long()]
    x = conv_bn_stack(x, v, 'encoder', 3, torch.relu, training, 0.5, masks.get('encoder_conv'))
    tl = feed['Token_Length']
    p = 'encoder/bilstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn'
    mf, mb = masks['encoder_bilstm'][0] if training else (None, None)
    fw = dynamic_rnn(x, tl, v[p + '/fw/zoneout_lstm_cell/kernel'], v[p + '/fw/zoneout_lstm_cell/bias'], training, mf)
    bw = dynamic_rnn(reverse_rows(x, tl), tl, v[p + '/bw/zoneout_lstm_cell/kernel'], v[p + '/bw/zoneout_lstm_cell/bias'],
                     training, mb)
    x = torch.cat([fw, reverse_rows(bw, tl)], dim=-1)
    memory = torch.cat([x, e[:, None, :].expand(-1, x.shape[1], -1)], dim=-1)
    w = {short: v[name] for short, name in TF_VARIABLE_NAMES.items()}
    pm, zm = masks['decoder']
    lin, stop, align = D.decoder_forward(w, memory, tl, feed['Mel'], feed['Mel_Length'], pm, zm if training else None,
                                         is_training=training, max_steps=max_steps)
    post = conv_bn_stack(lin, v, 'decoder', 5, torch.tanh, training, 0.5, masks.get('postnet'))
    return lin, stop, align, lin + post, memory


def losses(v, lin, stop, post, feed, wr_keys, use_l1=True, wr_rate=1e-6):
    mel, mel_len = feed['Mel'], feed['Mel_Length']
    l1, sl = D.decoder_loss(lin, stop, mel, mel_len, use_l1)
    d = post[:, :-1] - mel
    pl = (d * d).mean() + (d.abs().mean() if use_l1 else 0.0)
    wr = wr_rate * sum(0.5 * (v[k] ** 2).sum() for k in wr_keys)
    return l1, pl, sl, wr
