"""CPU tests of the oracle: cross-checks against independent implementations, properties, golden fixtures.

The reference has no tests and TF1 cannot run here (parity unpinned); these checks pin the
restatement as far as this image allows."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import decoder_oracle as O
from multi_speaker_tts_b200 import synthetic as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lstm_cell_matches_torch_lstmcell():
    """zoneout 0 == plain LSTM; TF gate order i,j,f,o -> torch i,f,g,o; forget_bias 1.0 folded into torch bias."""
    torch.manual_seed(0)
    B, I, H = 3, 20, 16
    kernel = torch.randn(I + H, 4 * H, dtype=torch.float64) * 0.3
    bias = torch.randn(4 * H, dtype=torch.float64) * 0.1
    x, c, h = (torch.randn(B, n, dtype=torch.float64) for n in (I, H, H))
    m, zc, zh = O.lstm_cell(x, c, h, kernel, bias, None, None, zoneout=0.0)
    cell = torch.nn.LSTMCell(I, H).double()
    i_, j_, f_, o_ = torch.split(kernel, H, dim=1)
    bi, bj, bf, bo = torch.split(bias, H)
    Wt = torch.cat([i_, f_, j_, o_], dim=1)  # torch order: i, f, g, o
    with torch.no_grad():
        cell.weight_ih.copy_(Wt[:I].t())
        cell.weight_hh.copy_(Wt[I:].t())
        cell.bias_ih.copy_(torch.cat([bi, bf + 1.0, bj, bo]))
        cell.bias_hh.zero_()
        h2, c2 = cell(x, (h, c))
    assert torch.allclose(m, h2, atol=1e-12) and torch.allclose(zc, c2, atol=1e-12) and torch.allclose(zh, h2, atol=1e-12)


def test_zoneout_semantics():
    """state' = (1-r)*mask*(new-old)+old; output m is un-zoned; the (1-r) factor stays without a mask."""
    torch.manual_seed(1)
    B, I, H = 2, 8, 4
    kernel = torch.randn(I + H, 4 * H, dtype=torch.float64)
    bias = torch.zeros(4 * H, dtype=torch.float64)
    x, c, h = (torch.randn(B, n, dtype=torch.float64) for n in (I, H, H))
    m_plain, c_plain, _ = O.lstm_cell(x, c, h, kernel, bias, None, None, zoneout=0.0)
    mc = torch.tensor([[1, 0, 1, 0], [0, 0, 1, 1]], dtype=torch.float64)
    mh = 1 - mc
    m, zc, zh = O.lstm_cell(x, c, h, kernel, bias, mc, mh, zoneout=0.1)
    assert torch.allclose(m, m_plain)
    assert torch.allclose(zc, 0.9 * mc * (c_plain - c) + c)
    assert torch.allclose(zh, 0.9 * mh * (m_plain - h) + h)
    _, zc2, _ = O.lstm_cell(x, c, h, kernel, bias, None, None, zoneout=0.1)
    assert torch.allclose(zc2, 0.9 * (c_plain - c) + c)


def test_attention_step_against_loops():
    torch.manual_seed(2)
    B, Te, D = 2, 9, 12
    w = {k: torch.randn(s, dtype=torch.float64) * 0.3 for k, s in O.weight_shapes(D).items()}
    keys = torch.randn(B, Te, O.ATT, dtype=torch.float64)
    values = torch.randn(B, Te, D, dtype=torch.float64)
    text_len = torch.tensor([9, 5])
    query = torch.randn(B, O.CELL, dtype=torch.float64)
    cum = torch.rand(B, Te, dtype=torch.float64)
    a, new_cum, ctx = O.attention_step(w, keys, values, text_len, query, cum)
    Wc, bc, Wd = w['location/conv1d/kernel'], w['location/conv1d/bias'], w['location/dense/kernel']
    for b in range(B):
        q = query[b] @ w['query_layer/kernel']
        e = torch.full((Te,), -math.inf, dtype=torch.float64)
        for x in range(int(text_len[b])):
            f = bc.clone()
            for k in range(O.CONV_K):
                xx = x + k - O.CONV_K // 2
                if 0 <= xx < Te:
                    f = f + cum[b, xx] * Wc[k, 0]
            loc = f @ Wd
            e[x] = (w['score/weight_w'] * torch.tanh(keys[b, x] + q + loc + w['score/bias_b'])).sum()
        ref = torch.softmax(e, 0)
        assert torch.allclose(a[b], ref, atol=1e-12)
        assert (a[b, int(text_len[b]):] == 0).all()
        assert torch.allclose(ctx[b], ref @ values[b], atol=1e-12)
    assert torch.allclose(new_cum, cum + a)


def test_decoder_loop_properties():
    w = S.init_decoder_weights(0, bias_scale=0.05)
    b = S.synthetic_decoder_batch(3, 20, 9, seed=12, ragged=True)
    lin, stop, al, st = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'],
                                          b['zone_mask'], return_state=True)
    T = int(b['mel_len'].max()) + 1          # Modules.py:215,395 => max(L)+1 outputs
    assert lin.shape == (3, T, 80) and stop.shape == (3, T) and al.shape == (3, T, 20)
    assert torch.allclose(al.sum(-1), torch.ones(3, T), atol=1e-5)
    assert torch.allclose(al.sum(1), st['cum'], atol=1e-5)   # cumulative alignment == sum of the history
    for i in range(3):
        assert (al[i, :, int(b['text_len'][i]):] == 0).all()


def test_context_rows_fold():
    """Quirk B-1: the two 768-row context blocks of cell_0/kernel multiply the same vector -> summing them is
    exact in real arithmetic (this is what the CUDA path does)."""
    w = S.init_decoder_weights(3)
    D = 768
    K0 = w['cell_0/kernel'].double()
    ctx = torch.randn(2, D, dtype=torch.float64)
    pre = torch.randn(2, 256, dtype=torch.float64)
    h = torch.randn(2, 1024, dtype=torch.float64)
    full = torch.cat([pre, ctx, ctx, h], 1) @ K0
    folded = pre @ K0[:256] + ctx @ (K0[256:256 + D] + K0[256 + D:256 + 2 * D]) + h @ K0[256 + 2 * D:]
    assert torch.allclose(full, folded, atol=1e-10)


def test_tf_adam_and_lr():
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    m = torch.zeros(2, dtype=torch.float64)
    v = torch.zeros(2, dtype=torch.float64)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    O.tf_adam_step(p, m, v, g, step=1, lr=1e-3, eps=1e-6)
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = torch.tensor([1.0, -2.0], dtype=torch.float64) - lr_t * (0.1 * g) / ((0.001 * g * g).sqrt() + 1e-6)
    assert torch.allclose(p, exp, atol=1e-15)
    assert O.learning_rate(0) == 1e-3 and abs(O.learning_rate(10000) - 5e-4) < 1e-12
    assert O.learning_rate(10 ** 7) == 1e-5


def test_decoder_loss_matches_torch():
    torch.manual_seed(3)
    B, L = 2, 5
    lin = torch.randn(B, L + 1, 80)
    stop = torch.randn(B, L + 1)
    mel = torch.randn(B, L, 80)
    mel_len = torch.tensor([5, 3])
    ll, sl = O.decoder_loss(lin, stop, mel, mel_len)
    assert torch.allclose(ll, F.mse_loss(lin[:, :-1], mel) + F.l1_loss(lin[:, :-1], mel))
    tgt = torch.tensor([[0, 0, 0, 0, 0, 1], [0, 0, 0, 1, 1, 1]], dtype=torch.float32)
    x = stop
    manual = (torch.clamp(x, min=0) - x * tgt + torch.log1p(torch.exp(-x.abs()))).mean()
    assert torch.allclose(sl, manual, atol=1e-6)


@pytest.mark.parametrize("name", ["decoder_b2_te12_l6", "decoder_b3_te20_l9_ragged"])
def test_oracle_reproduces_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = mg.decoder_case(*mg.DECODER_CASES[name])
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    for k in ref.files:
        np.testing.assert_allclose(got[k], ref[k], rtol=2e-4, atol=2e-5, err_msg=k)
