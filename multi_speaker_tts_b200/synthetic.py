"""Seeded synthetic inputs and TF-default initialisers for the decoder hot path.

Nothing here touches the oracle; tests and bench feed the same tensors to both sides.
"""
import math

import numpy as np
import torch

MEL, PRENET, CELL, ATT, CONV_K, CONV_C = 80, 256, 1024, 128, 31, 32


def decoder_weight_shapes(mem_dim=768):
    """name -> shape in TF layouts; names are the suffixes of the TF variable names (SURVEY A-8)."""
    return {
        'prenet_0/kernel': (MEL, PRENET), 'prenet_0/bias': (PRENET,),
        'prenet_1/kernel': (PRENET, PRENET), 'prenet_1/bias': (PRENET,),
        'cell_0/kernel': (PRENET + 2 * mem_dim + CELL, 4 * CELL), 'cell_0/bias': (4 * CELL,),
        'cell_1/kernel': (2 * CELL, 4 * CELL), 'cell_1/bias': (4 * CELL,),
        'memory_layer/kernel': (mem_dim, ATT),
        'query_layer/kernel': (CELL, ATT),
        'location/conv1d/kernel': (CONV_K, 1, CONV_C), 'location/conv1d/bias': (CONV_C,),
        'location/dense/kernel': (CONV_C, ATT),
        'score/weight_w': (ATT,), 'score/bias_b': (ATT,),
        'projection/kernel': (CELL + mem_dim, MEL + 1), 'projection/bias': (MEL + 1,),
    }


# TF variable names the short keys stand for (checkpoint-key compatibility, SURVEY A-8)
TF_VARIABLE_NAMES = {
    'prenet_0/kernel': 'decoder/decoder/prenet_0/dense/kernel',
    'prenet_0/bias': 'decoder/decoder/prenet_0/dense/bias',
    'prenet_1/kernel': 'decoder/decoder/prenet_1/dense/kernel',
    'prenet_1/bias': 'decoder/decoder/prenet_1/dense/bias',
    'cell_0/kernel': 'decoder/decoder/attention_wrapper/multi_rnn_cell/cell_0/zoneout_lstm_cell/kernel',
    'cell_0/bias': 'decoder/decoder/attention_wrapper/multi_rnn_cell/cell_0/zoneout_lstm_cell/bias',
    'cell_1/kernel': 'decoder/decoder/attention_wrapper/multi_rnn_cell/cell_1/zoneout_lstm_cell/kernel',
    'cell_1/bias': 'decoder/decoder/attention_wrapper/multi_rnn_cell/cell_1/zoneout_lstm_cell/bias',
    'memory_layer/kernel': 'attention/memory_layer/kernel',
    'query_layer/kernel': 'decoder/decoder/attention_wrapper/location_sensitive_attention/query_layer/kernel',
    'location/conv1d/kernel': 'decoder/decoder/attention_wrapper/location_sensitive_attention/'
                              'attention_convolution_dense_layer/conv1d/kernel',
    'location/conv1d/bias': 'decoder/decoder/attention_wrapper/location_sensitive_attention/'
                            'attention_convolution_dense_layer/conv1d/bias',
    'location/dense/kernel': 'decoder/decoder/attention_wrapper/location_sensitive_attention/'
                             'attention_convolution_dense_layer/dense/kernel',
    'score/weight_w': 'decoder/decoder/attention_wrapper/location_sensitive_attention/score_layer/weight_w',
    'score/bias_b': 'decoder/decoder/attention_wrapper/location_sensitive_attention/score_layer/bias_b',
    'projection/kernel': 'decoder/decoder/linear_projection/dense/kernel',
    'projection/bias': 'decoder/decoder/linear_projection/dense/bias',
}


def _fans(shape):
    if len(shape) == 1:
        return shape[0], shape[0]
    rf = 1
    for d in shape[:-2]:
        rf *= d
    return shape[-2] * rf, shape[-1] * rf


def init_decoder_weights(seed=0, mem_dim=768, bias_scale=0.0, dtype=torch.float32):
    """glorot-uniform kernels, zero biases (TF get_variable / tf.layers defaults, SURVEY A-6).

    ``bias_scale`` > 0 draws biases ~ U(-s, s) instead, so parity tests exercise every bias add."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in decoder_weight_shapes(mem_dim).items():
        if name.endswith('bias') or name.endswith('bias_b'):
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bias_scale
        else:
            fi, fo = _fans(shape if name != 'score/weight_w' else (1, 1, ATT))
            lim = math.sqrt(6.0 / (fi + fo))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        out[name] = t.to(dtype)
    return out


def synthetic_decoder_batch(B, Te, L, seed=1234, rank=0, ragged=False, mem_dim=768, training=True):
    """memory [B,Te,mem_dim] (encoder-like activations ++ a unit-norm speaker vector tiled over time),
    text_len, mel ~ clip(N(0,1.5), +-4), mel_len, prenet_mask [T,2,B,256] u8, zone_mask [T,2,2,B,1024] u8."""
    rng = np.random.default_rng(seed + rank)
    enc = np.tanh(rng.standard_normal((B, Te, mem_dim - 256))).astype(np.float32) * 0.5
    spk = rng.standard_normal((B, 256)).astype(np.float32)
    spk /= np.linalg.norm(spk, axis=1, keepdims=True)
    memory = np.concatenate([enc, np.repeat(spk[:, None, :], Te, axis=1)], axis=2)
    mel = np.clip(rng.standard_normal((B, L, MEL)) * 1.5, -4, 4).astype(np.float32)
    if ragged:
        text_len = rng.integers(Te // 2, Te + 1, size=B).astype(np.int32)
        mel_len = rng.integers(L // 2, L + 1, size=B).astype(np.int32)
        text_len[0] = Te
        mel_len[0] = L
        for b in range(B):
            mel[b, mel_len[b]:] = 0.0
    else:
        text_len = np.full(B, Te, np.int32)
        mel_len = np.full(B, L, np.int32)
    T = L + 1
    prenet_mask = (rng.random((T, 2, B, PRENET)) < 0.5).astype(np.uint8)
    zone_mask = (rng.random((T, 2, 2, B, CELL)) < 0.9).astype(np.uint8)
    t = torch.from_numpy
    return {
        'memory': t(memory), 'text_len': t(text_len), 'mel': t(mel), 'mel_len': t(mel_len),
        'prenet_mask': t(prenet_mask), 'zone_mask': t(zone_mask),
    }
