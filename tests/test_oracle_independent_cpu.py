"""A second, independently structured reading of the reference's decoder graph, checked against oracle/decoder_oracle.py.

The oracle cannot be pinned against TensorFlow here (parity unpinned, see its header), so it is cross-examined instead: this
file re-derives the training-mode decoder loop from the reference sources with different building blocks --
``torch.nn.LSTMCell`` with converted weights instead of the concat-matmul-split cell, a batched (hoisted) prenet over all
teacher-forced frames instead of a per-step one, an explicit unfold-based location convolution instead of ``F.conv1d``,
``log_softmax`` over a length-sliced row instead of a -inf mask -- and must reproduce the oracle's outputs.  What both readings
share are the [TF-internal] facts of SURVEY Appendix A (AttentionWrapper wiring, BahdanauAttention masking); a slip in either
implementation's step order, state routing, gate order or masking shows up as a mismatch.

Reference lines re-read for this file: Modules.py:76-119 (wrapper arguments: output_attention=False, attention_layer_size=None,
alignment_history=True), :178-185 (first input), :212-237 (teacher forcing by ``time``), :239-255 (prenet, dropout always on),
:286-321 (projection of concat(cell output, context)), :387-443 (loop: finished OR, one extra step at time = max length);
ZoneoutLSTMCell.py:228-264 (gate order i,j,f,o; forget bias; output m un-zoned, state zoned);
Location_Sensitive_Attention.py:43-85 (query layer, conv over the CUMULATIVE alignment, energies, cumulate)."""
import pytest
import torch

from multi_speaker_tts_b200 import synthetic as S

H = 1024


def _torch_cell(kernel, bias, in_dim):
    """tf LSTM kernel [in + H, 4H] with gate columns i, j, f, o (forget bias 1.0 added at run time, ZoneoutLSTMCell.py:237)
    -> torch.nn.LSTMCell whose rows are ordered i, f, g, o."""
    cell = torch.nn.LSTMCell(in_dim, H).double()
    i, j, f, o = [kernel[:, k * H:(k + 1) * H] for k in range(4)]
    w = torch.cat([i, f, j, o], dim=1).t()              # [4H, in + H]
    bi, bj, bf, bo = [bias[k * H:(k + 1) * H] for k in range(4)]
    with torch.no_grad():
        cell.weight_ih.copy_(w[:, :in_dim])
        cell.weight_hh.copy_(w[:, in_dim:])
        cell.bias_ih.copy_(torch.cat([bi, bf + 1.0, bj, bo]))
        cell.bias_hh.zero_()
    return cell


def _second_reading(w, memory, text_len, mel, mel_len, prenet_mask, zone_mask):
    w = {k: v.double() for k, v in w.items()}
    memory, mel = memory.double(), mel.double()
    B, Te, D = memory.shape
    T = int(mel_len.max()) + 1                            # loop runs time = 0 .. max length (Modules.py:387-395, :215)
    # ---- everything that does not depend on the recurrence, batched over time ----
    frames = torch.zeros(T, B, 80, dtype=torch.float64)   # input of step t: zeros at t = 0, else mel frame t-1 (read(time) at t-1)
    for t in range(1, T):
        if t - 1 < int(mel_len.max()):                    # after the step at time = max length nothing is read any more
            frames[t] = mel[:, t - 1]
    pm = prenet_mask[:T].double()
    h1 = torch.relu(frames @ w['prenet_0/kernel'] + w['prenet_0/bias']) * 2.0 * pm[:, 0]
    pre = torch.relu(h1 @ w['prenet_1/kernel'] + w['prenet_1/bias']) * 2.0 * pm[:, 1]
    values = memory.clone()
    for b in range(B):
        values[b, int(text_len[b]):] = 0.0                # BahdanauAttention: memory masked beyond its length
    keys = values @ w['memory_layer/kernel']
    # location filter as one [31, 128] matrix: conv1d(1 -> 32, k = 31, 'same') followed by dense(32 -> 128, no bias)
    loc_filter = w['location/conv1d/kernel'][:, 0, :] @ w['location/dense/kernel']          # [31, 128]
    loc_bias = w['location/conv1d/bias'] @ w['location/dense/kernel']                        # [128]
    cell0 = _torch_cell(w['cell_0/kernel'], w['cell_0/bias'], 256 + 2 * D)
    cell1 = _torch_cell(w['cell_1/kernel'], w['cell_1/bias'], H)
    c0 = h0 = c1 = h1s = torch.zeros(B, H, dtype=torch.float64)
    attention = torch.zeros(B, D, dtype=torch.float64)    # AttentionWrapper.zero_state
    cum = torch.zeros(B, Te, dtype=torch.float64)
    lin, stop, aligns = [], [], []
    zm = zone_mask[:T].double()
    with torch.no_grad():
        for t in range(T):
            helper_input = torch.cat([pre[t], attention], -1)          # Modules.py:232-234
            cell_input = torch.cat([helper_input, attention], -1)      # default cell_input_fn: concat([inputs, attention])
            m0, cn0 = cell0(cell_input, (h0, c0))                       # torch returns (h', c')
            c0 = 0.9 * zm[t, 0, 0] * (cn0 - c0) + c0
            h0 = 0.9 * zm[t, 0, 1] * (m0 - h0) + h0
            m1, cn1 = cell1(m0, (h1s, c1))                              # MultiRNNCell: layer 1 reads layer 0's OUTPUT m (un-zoned)
            c1 = 0.9 * zm[t, 1, 0] * (cn1 - c1) + c1
            h1s = 0.9 * zm[t, 1, 1] * (m1 - h1s) + h1s
            q = m1 @ w['query_layer/kernel']
            padded = torch.nn.functional.pad(cum, (15, 15))
            windows = padded.unfold(1, 31, 1)                           # [B, Te, 31]: cross-correlation, 'same'
            loc = windows @ loc_filter + loc_bias
            e = (torch.tanh(keys + q[:, None, :] + loc + w['score/bias_b']) * w['score/weight_w']).sum(-1)
            a = torch.zeros(B, Te, dtype=torch.float64)
            for b in range(B):
                n = int(text_len[b])
                a[b, :n] = torch.log_softmax(e[b, :n], -1).exp()
            cum = cum + a
            attention = torch.einsum('bt,btd->bd', a, values)
            out = torch.cat([m1, attention], -1) @ w['projection/kernel'] + w['projection/bias']
            lin.append(out[:, :80]); stop.append(out[:, 80]); aligns.append(a)
    return torch.stack(lin, 1), torch.stack(stop, 1), torch.stack(aligns, 1), cum


@pytest.mark.parametrize("B,Te,L,ragged,seed", [(2, 12, 6, False, 1), (3, 20, 9, True, 2), (4, 33, 14, True, 3)])
def test_decoder_oracle_matches_second_reading(B, Te, L, ragged, seed):
    from oracle import decoder_oracle as O
    w = S.init_decoder_weights(seed, bias_scale=0.05)
    b = S.synthetic_decoder_batch(B, Te, L, seed=10 + seed, ragged=ragged)
    w64 = {k: v.double() for k, v in w.items()}
    with torch.no_grad():
        lin, stop, al, st = O.decoder_forward(w64, b['memory'].double(), b['text_len'], b['mel'].double(), b['mel_len'],
                                              b['prenet_mask'], b['zone_mask'], return_state=True)
    lin2, stop2, al2, cum2 = _second_reading(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], b['zone_mask'])
    assert lin.shape == lin2.shape and al.shape == al2.shape
    assert (lin - lin2).abs().max() < 1e-9
    assert (stop - stop2).abs().max() < 1e-9
    assert (al - al2).abs().max() < 1e-10
    assert (st['cum'] - cum2).abs().max() < 1e-9           # alignment history sums to the final attention state
    assert torch.equal(al.argmax(-1), al2.argmax(-1))
