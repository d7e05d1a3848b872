"""Location_Sensitive_Attention surface (Location_Sensitive_Attention.py:12-85).

In the reference this object owns the memory, its mask, the memory/query layers, the location conv + dense and the
score vectors, and is called once per decoder step by the AttentionWrapper.  Here it owns the same variables and the
same constructor arguments; the per-step ``__call__`` does not exist as a separate op -- ``Modules.Decoder_LSTM``
passes the whole mechanism to the fused persistent decoder kernel (csrc/decoder_fwd*.cu), which computes query
projection, location features, energies, masked softmax, cumulative alignment and context inside the loop.
"""
import math

import torch

VARIABLE_KEYS = ['memory_layer/kernel', 'query_layer/kernel', 'location/conv1d/kernel', 'location/conv1d/bias',
                 'location/dense/kernel', 'score/weight_w', 'score/bias_b']


def _glorot(shape, fan_in, fan_out, generator):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=generator, dtype=torch.float64) * 2 - 1) * lim).float()


class Location_Sensitive_Attention(object):
    def __init__(self, num_units, memory, memory_length, conv_kernel_size, conv_stride_size, conv_channel, dropout_rate,
                 is_training=False, normalize=False, probability_fn=None, score_mask_value=None, dtype=None,
                 name='location_sensitive_attention', variables=None, generator=None, query_depth=1024):
        if conv_stride_size != 1:
            # any other stride changes Te of the location features and breaks the broadcast at :82 (SURVEY B-14)
            raise ValueError("Location_Sensitive_Attention supports conv_stride_size == 1 only")
        if not memory.is_cuda:
            raise RuntimeError("Location_Sensitive_Attention: memory must be a CUDA tensor (no CPU fallback)")
        self.num_units = num_units
        self.memory = memory                    # [B, Te, D], un-masked; the kernel applies sequence_mask(memory_length)
        self.memory_length = memory_length      # [B] int32
        self.conv_kernel_size, self.conv_channel = conv_kernel_size, conv_channel
        self.dropout_rate = dropout_rate        # stored but never used by the reference (:31)
        self.is_training = is_training
        self.name = name
        D = memory.shape[2]
        if variables is None:
            g = generator
            variables = {
                'memory_layer/kernel': _glorot((D, num_units), D, num_units, g),
                'query_layer/kernel': _glorot((query_depth, num_units), query_depth, num_units, g),
                'location/conv1d/kernel': _glorot((conv_kernel_size, 1, conv_channel), conv_kernel_size,
                                                  conv_kernel_size * conv_channel, g),
                'location/conv1d/bias': torch.zeros(conv_channel),
                'location/dense/kernel': _glorot((conv_channel, num_units), conv_channel, num_units, g),
                'score/weight_w': _glorot((num_units,), num_units, num_units, g),   # xavier over [1,1,units] (:33)
                'score/bias_b': torch.zeros(num_units),
            }
            variables = {k: v.to(memory.device) for k, v in variables.items()}
        missing = [k for k in VARIABLE_KEYS if k not in variables]
        if missing:
            raise KeyError("Location_Sensitive_Attention: missing variables %s" % missing)
        self.variables = {k: variables[k] for k in VARIABLE_KEYS}

    @property
    def batch_size(self):
        return self.memory.shape[0]

    @property
    def alignments_size(self):
        return self.memory.shape[1]
