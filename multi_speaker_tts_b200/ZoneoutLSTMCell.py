"""ZoneoutLSTMCell surface of the reference (ZoneoutLSTMCell.py:48-56,188-271) on torch tensors.

The decoder's two cells never go through this class at run time: ``Modules.Decoder_LSTM`` hands their kernels to the
fused persistent CUDA loop.  The class is what the callers either side of the decoder use (encoder BiLSTM, speaker
embedding stack): a parameter holder with the reference's ``call(inputs, (c, h)) -> (m, (c', h'))``
contract, computed with torch ops on whatever device the tensors live on.
"""
import math

import torch


class LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return super(LSTMStateTuple, cls).__new__(cls, (c, h))

    c = property(lambda self: self[0])
    h = property(lambda self: self[1])


class ZoneoutLSTMCell(object):
    def __init__(self, num_units, is_training=False, cell_zoneout_rate=0.0, output_zoneout_rate=0.0, use_peepholes=False,
                 cell_clip=None, initializer=None, num_proj=None, proj_clip=None, num_unit_shards=None, num_proj_shards=None,
                 forget_bias=1.0, state_is_tuple=True, activation=None, reuse=None, name=None, input_size=None,
                 kernel=None, bias=None, device=None, generator=None):
        if use_peepholes or cell_clip is not None or num_proj is not None or not state_is_tuple:
            raise NotImplementedError("peepholes / clipping / projection / non-tuple state are unused by the reference "
                                      "configuration (Hyper_Parameters.py) and not implemented")
        self.num_units = num_units
        self.is_training = is_training
        self.cell_zoneout_rate = cell_zoneout_rate
        self.output_zoneout_rate = output_zoneout_rate
        self.forget_bias = forget_bias
        self.name = name
        self.kernel, self.bias = kernel, bias
        if kernel is None and input_size is not None:
            self.build(input_size, device, generator)

    @property
    def state_size(self):
        return LSTMStateTuple(self.num_units, self.num_units)

    @property
    def output_size(self):
        return self.num_units

    def build(self, input_size, device=None, generator=None):
        """kernel [input + units, 4 units] glorot-uniform, bias zeros (ZoneoutLSTMCell.py:160-166)"""
        rows, cols = input_size + self.num_units, 4 * self.num_units
        lim = math.sqrt(6.0 / (rows + cols))
        k = (torch.rand(rows, cols, generator=generator, dtype=torch.float64) * 2 - 1) * lim
        self.kernel = k.float().to(device)
        self.bias = torch.zeros(cols, device=device)

    def zero_state(self, batch_size, dtype=torch.float32):
        z = torch.zeros(batch_size, self.num_units, dtype=dtype, device=self.kernel.device)
        return LSTMStateTuple(z, z.clone())

    def call(self, inputs, state, masks=None):
        """One step.  ``masks`` = (mask_c, mask_h) 0/1 tensors (training only; drawn here when omitted).
        Gate order i, j, f, o; returns the UN-zoned m as output and the zoned (c, h) as state (:259-264)."""
        c_prev, h_prev = state
        lstm_matrix = torch.cat([inputs, h_prev], dim=1) @ self.kernel + self.bias
        i, j, f, o = torch.split(lstm_matrix, self.num_units, dim=1)
        c = torch.sigmoid(f + self.forget_bias) * c_prev + torch.sigmoid(i) * torch.tanh(j)
        m = torch.sigmoid(o) * torch.tanh(c)
        dc, dm = c - c_prev, m - h_prev
        if self.is_training:
            if masks is None:
                masks = (torch.floor(torch.rand_like(dc) + (1.0 - self.cell_zoneout_rate)),
                         torch.floor(torch.rand_like(dm) + (1.0 - self.output_zoneout_rate)))
            dc, dm = dc * masks[0], dm * masks[1]
        # dropout_no_scale at inference is the identity, the (1 - rate) factor stays (:266-271)
        c_z = (1.0 - self.cell_zoneout_rate) * dc + c_prev
        h_z = (1.0 - self.output_zoneout_rate) * dm + h_prev
        return m, LSTMStateTuple(c_z, h_z)

    __call__ = call
