"""Host-side surface (MSTTS_SV / Feeder / Modules callers either side of the decoder) on the CPU: the batched library-op
implementations in multi_speaker_tts_b200.Modules against the row-by-row oracle, the variable inventory, the
weight-regularisation name filter and the Feeder contracts.  (Decoder_LSTM itself needs the GPU: tests/test_tacotron2_gpu.py.)"""
import math

import numpy as np
import pytest
import torch

from multi_speaker_tts_b200 import Feeder, Hyper_Parameters as hp


def _variables(seed=0, dtype=torch.float64):
    from multi_speaker_tts_b200 import MSTTS_SV as M
    gen = torch.Generator().manual_seed(seed)
    v = {}
    for k, (s, kind) in M.variable_shapes().items():
        t = M._init(s, kind, gen).to(dtype)
        if kind in ('zeros', 'ones', 'moving_mean', 'moving_variance'):  # make every bias / BN term matter
            t = t + 0.1 * torch.randn(s, generator=gen, dtype=dtype) if kind != 'moving_variance' else \
                t + 0.3 * torch.rand(s, generator=gen, dtype=dtype)
        v[k] = t
    return v


def test_variable_inventory_and_regularised_set():
    from multi_speaker_tts_b200 import MSTTS_SV as M
    s = M.variable_shapes()
    n = sum(math.prod(v[0]) for k, v in s.items() if M.is_trainable(k, v[1]))
    assert n == 30278977  # "~30.3 M params" in the Tacotron2 trainable set
    wr = [k for k, v in s.items() if M.is_trainable(k, v[1]) and M.in_weight_regularization(k)]
    assert 'attention/memory_layer/kernel' in wr and 'encoder/conv_0/batch_normalization/beta' in wr
    for k in wr:  # MSTTS_SV.py:145-159
        assert not any(x in k.lower() for x in ('bias', 'embedding', 'lstm', 'rnn', 'weight_w', 'projection'))
    assert not any('multi_rnn_cell' in k or 'linear_projection' in k for k in wr)
    assert s['decoder/decoder/attention_wrapper/multi_rnn_cell/cell_0/zoneout_lstm_cell/kernel'][0] == (2816, 4096)


def test_feeder_contracts():
    f = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(3, 20, 50))
    p = f.placeholder_Dict
    assert set(p) == {"Is_Training", "Token", "Token_Length", "Mel", "Mel_Length", "Speaker_Embedding_Mel"}
    d = f.Get_Train_Pattern()
    assert d[p['Token']].dtype == np.int32 and d[p['Token']].shape == (3, 20)
    assert (d[p['Token']][:, 0] == 0).all() and (d[p['Token']][:, -1] == 1).all()      # <S> ... <E>
    assert d[p['Mel']].shape == (3, 50, 80) and d[p['Mel']].dtype == np.float32
    assert d[p['Speaker_Embedding_Mel']].shape == (15, 64, 80)
    g = Feeder.Feeder(is_Training=False, synthetic=True)
    q = g.Get_Inference_Pattern(['nowhere_a.wav', 'nowhere_b.wav'], ['Hello, world!', 'Hi.'])
    tok = q[g.placeholder_Dict['Token']]
    assert tok.shape == (2, 15) and (tok[1, 5:] == 1).all()                           # padded with <E> (Feeder.py:155)
    assert q[g.placeholder_Dict['Mel']].shape == (2, 1, 80) and (q[g.placeholder_Dict['Mel_Length']] == 0).all()
    assert q[g.placeholder_Dict['Token_Length']].tolist() == [15, 5]


def test_speaker_embedding_windows():
    f = Feeder.Feeder(is_Training=False, synthetic=True)
    long = np.arange(400 * 80, dtype=np.float32).reshape(400, 80)
    short = np.ones((30, 80), np.float32)
    w = f.Speaker_Embedding_Mel([long, short]).reshape(2, 5, 64, 80)
    start = int((400 - 192) / 2)
    for s in range(5):
        assert np.array_equal(w[0, s], long[start + 32 * s:start + 32 * s + 64])
    assert (w[1, :, :30] == 1).all() and (w[1, :, 30:] == 0).all()


def test_trim_drops_quiet_edges():
    y = np.concatenate([np.zeros(320), np.sin(np.arange(1600) * 0.3), np.zeros(480)]).astype(np.float32)
    t = Feeder._trim(y)
    assert 1500 <= t.shape[0] <= 1700 and abs(t).max() > 0.9


@pytest.mark.parametrize("training", [True, False])
def test_encoder_and_postnet_match_oracle(training):
    from oracle import tacotron2_oracle as O
    from multi_speaker_tts_b200 import Modules
    torch.manual_seed(0)
    v = _variables()
    B, Te, T = 3, 11, 9
    tok = torch.randint(2, 42, (B, Te), dtype=torch.int32)
    tl = torch.tensor([11, 7, 4], dtype=torch.int32)
    masks_c = [(torch.rand(B, Te, 512) < 0.5).double() for _ in range(3)]
    mf = (torch.rand(Te, 2, B, 256) < 0.9).double()
    mb = (torch.rand(Te, 2, B, 256) < 0.9).double()
    # batched library-op path
    va = {k: t.clone() for k, t in v.items()}
    x = Modules.Encoder_Embedding(tok, va)
    x = Modules.Encoder_Conv(x, training, va, masks_c)
    x = Modules.Encoder_BiLSTM(x, tl, training, va, [(mf, mb)])
    # oracle
    y = v['encoder/embedding_variable'][tok.long()]
    y = O.conv_bn_stack(y, v, 'encoder', 3, torch.relu, training, 0.5, masks_c)
    p = 'encoder/bilstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn'
    fw = O.dynamic_rnn(y, tl, v[p + '/fw/zoneout_lstm_cell/kernel'], v[p + '/fw/zoneout_lstm_cell/bias'], training, mf)
    bw = O.dynamic_rnn(O.reverse_rows(y, tl), tl, v[p + '/bw/zoneout_lstm_cell/kernel'], v[p + '/bw/zoneout_lstm_cell/bias'],
                       training, mb)
    y = torch.cat([fw, O.reverse_rows(bw, tl)], dim=-1)
    assert x.shape == (B, Te, 512) and (x - y).abs().max() < 1e-10
    assert (x[1, 7:] == 0).all() and (x[2, 4:] == 0).all()  # zero beyond Token_Length
    if training:  # moving statistics moved by (1 - 0.99) towards the batch statistics
        assert not torch.equal(va['encoder/conv_0/batch_normalization/moving_mean'],
                               v['encoder/conv_0/batch_normalization/moving_mean'])
    lin = torch.randn(B, T, 80, dtype=torch.float64)
    masks_p = [(torch.rand(B, T, 512 if i < 4 else 80) < 0.5).double() for i in range(5)]
    a = Modules.Decoder_Conv(lin, training, va, masks_p)
    b = O.conv_bn_stack(lin, v, 'decoder', 5, torch.tanh, training, 0.5, masks_p)
    assert (a - b).abs().max() < 1e-10


def test_speaker_embedding_matches_oracle_and_is_normalised_over_the_batch():
    from oracle import tacotron2_oracle as O
    from multi_speaker_tts_b200.Speaker_Embedding import Modules as S
    v = _variables()
    se = torch.randn(2 * 5, 12, 80, dtype=torch.float64)
    x = S.Restructure(se, v)
    x = S.Stack_LSTM(x, torch.full((10,), 12), False, v)
    e = S.Inference(x)
    ref = O.speaker_embedding(v, se)
    assert e.shape == (2, 256) and (e - ref).abs().max() < 1e-10
    assert abs(float((e * e).sum()) - 1.0) < 1e-9  # whole-tensor l2 norm (quirk B-4), not per row


def test_zoneout_cell_class_matches_oracle_cell():
    from oracle import decoder_oracle as D
    from multi_speaker_tts_b200.ZoneoutLSTMCell import ZoneoutLSTMCell
    g = torch.Generator().manual_seed(1)
    cell = ZoneoutLSTMCell(16, is_training=True, cell_zoneout_rate=0.1, output_zoneout_rate=0.1, input_size=8, generator=g)
    x, c, h = torch.randn(4, 8), torch.randn(4, 16), torch.randn(4, 16)
    mc, mh = (torch.rand(4, 16) < 0.9).float(), (torch.rand(4, 16) < 0.9).float()
    m, (c2, h2) = cell(x, (c, h), masks=(mc, mh))
    rm, rc, rh = D.lstm_cell(x, c, h, cell.kernel, cell.bias, mc, mh)
    assert torch.allclose(m, rm, atol=1e-6) and torch.allclose(c2, rc, atol=1e-6) and torch.allclose(h2, rh, atol=1e-6)
    with pytest.raises(NotImplementedError):
        ZoneoutLSTMCell(16, num_proj=8)


def test_feeder_reads_the_reference_pickle_dataset(tmp_path):
    """the on-disk format of Pattern_Generate.py:66-76,245-272: one pickle per utterance ('Token', 'Mel', 'Text', 'Dataset') and
    METADATA.PICKLE with the length / dataset dictionaries; the Feeder thread batches, pads with <E> and adds <S>/<E>"""
    import pickle
    rng = np.random.default_rng(0)
    files, mel_len, tok_len, dataset = [], {}, {}, {}
    for i in range(5):
        name = 'VCTK.UTT%d.PICKLE' % i
        T = 45 + 7 * i  # frames; Use_Wav_Length_Range (500..9000 ms) / Frame_Shift 12.5 ms => 40..720 frames
        pat = {'Token': rng.integers(2, 42, size=6 + i).astype(np.int32), 'Mel': rng.standard_normal((T, 80)).astype(np.float32),
               'Text': 'X' * (6 + i), 'Dataset': 'VCTK'}
        with open(tmp_path / name, 'wb') as f:
            pickle.dump(pat, f, protocol=2)
        files.append(name)
        mel_len[name], tok_len[name], dataset[name] = T, 6 + i, 'VCTK'
    meta = {'Token_Index_Dict': dict(Feeder.TOKEN_INDEX_DICT), 'Spectrogram_Dim': hp.Sound.Spectrogram_Dim, 'Mel_Dim': hp.Sound.Mel_Dim,
            'Frame_Shift': hp.Sound.Frame_Shift, 'Frame_Length': hp.Sound.Frame_Length, 'Sample_Rate': hp.Sound.Sample_Rate,
            'File_List': files, 'Token_Length_Dict': tok_len, 'Mel_Length_Dict': mel_len, 'Dataset_Dict': dataset}
    with open(tmp_path / 'METADATA.PICKLE', 'wb') as f:
        pickle.dump(meta, f, protocol=2)
    old = (hp.Train.Pattern_Path, hp.Train.Batch_Size)
    hp.Train.Pattern_Path, hp.Train.Batch_Size = str(tmp_path), 3
    try:
        f = Feeder.Feeder(is_Training=True)
        assert not f.synthetic
        d = f.Get_Train_Pattern()
    finally:
        hp.Train.Pattern_Path, hp.Train.Batch_Size = old
    p = f.placeholder_Dict
    tok, tl, mel, ml = d[p['Token']], d[p['Token_Length']], d[p['Mel']], d[p['Mel_Length']]
    assert tok.shape[0] in (2, 3) and tok.dtype == np.int32 and mel.shape[2] == 80
    for b in range(tok.shape[0]):
        assert tok[b, 0] == 0 and tok[b, tl[b] - 1] == 1 and (tok[b, tl[b]:] == 1).all()   # <S> ... <E>, <E> padding
        assert (mel[b, ml[b]:] == 0).all()
    assert list(ml) == sorted(ml)                                                            # Pattern_Sorting_by_Mel_Length
    assert d[p['Speaker_Embedding_Mel']].shape == (tok.shape[0] * 5, 64, 80)


def test_stream_and_dense_helpers_are_transparent_on_the_cpu():
    """Modules.side_stream / Modules.dense / Feeder._pinned only change WHERE work runs on a CUDA device; on CPU tensors they must be
    the plain expressions (the CPU graph tests above go through them)"""
    from multi_speaker_tts_b200 import Modules
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 7, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(7, 4, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    y = Modules.dense(x, w, b)
    assert torch.equal(y, x @ w + b)
    assert torch.equal(Modules.dense(x, w), x @ w)
    y.sum().backward()
    assert x.grad is not None and w.grad is not None
    branch = Modules.side_stream(x.device, 'anything')
    with branch:
        z = x * 2
    assert branch.join(z) is z
    a, c = branch.join(z, y)
    assert a is z and c is y
    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    out = Feeder._pinned(arr)
    assert isinstance(out, np.ndarray) and out.dtype == arr.dtype and np.array_equal(out, arr)


def test_weight_regularisation_loss_is_one_reduction_over_the_flat_buffer():
    """MSTTS_SV.py:145-159: 0.5 * rate * sum over the regularised variables of sum(v ** 2).  The flat parameter buffer holds exactly
    those variables in its first n_l2 floats (padding zero), which the train step reduces in one call"""
    from multi_speaker_tts_b200 import MSTTS_SV as M
    shapes = M.variable_shapes()
    train = [k for k, (s, kind) in shapes.items() if M.is_trainable(k, kind)]
    reg = [k for k in train if M.in_weight_regularization(k)]
    gen = torch.Generator().manual_seed(1)
    vals = {k: torch.randn(shapes[k][0], generator=gen, dtype=torch.float64) for k in reg[:6]}
    flat = torch.zeros(sum((v.numel() + 3) // 4 * 4 for v in vals.values()), dtype=torch.float64)
    off = 0
    for v in vals.values():
        flat[off:off + v.numel()] = v.reshape(-1)
        off += (v.numel() + 3) // 4 * 4
    ref = sum(0.5 * (v ** 2).sum() for v in vals.values())
    got = 0.5 * torch.linalg.vector_norm(flat) ** 2
    assert abs(float(got) - float(ref)) <= 1e-12 * float(ref)
