"""Speaker-embedding forward (Speaker_Embedding/Modules.py:6-37,127-137): frozen producer of the 256-d conditioning
vector.  Reuses ``Modules.zoneout_lstm_sequence`` / ``Modules.dense`` (the library's own kernels)."""
import torch

from .. import Hyper_Parameters as hp
from ..Modules import dense, zoneout_lstm_sequence


def Restructure(inputs, variables):
    """dense 80 -> Embedding_Size so the first residual connection type-checks (:6-10)"""
    return dense(inputs, variables['speaker_embedding/dense/kernel'], variables['speaker_embedding/dense/bias'])


def Stack_LSTM(inputs, lengths, is_training=False, variables=None):
    """3 ZoneoutLSTMCells (256), ResidualWrapper on all but the last (:12-37)"""
    x = inputs
    n = hp.Speaker_Embedding.LSTM.Nums
    for i in range(n):
        p = 'speaker_embedding/lstm/rnn/multi_rnn_cell/cell_%d/lstmcell_%d' % (i, i)
        x, _ = zoneout_lstm_sequence(x, lengths, variables[p + '/kernel'], variables[p + '/bias'], is_training,
                                     hp.Speaker_Embedding.LSTM.Zoneout_Rate,
                                     residual=hp.Speaker_Embedding.LSTM.Use_Residual and i < n - 1)
    return x


def Inference(inputs):
    """last frame, mean over Sample_Nums windows, then l2-normalise over the WHOLE [B,256] tensor -- tf.nn.l2_normalize
    without an axis (:127-137, quirk B-4): the embedding scale is 1/sqrt(B)"""
    s = hp.Speaker_Embedding.Inference.Sample_Nums
    x = inputs[:, -1, :].reshape(inputs.shape[0] // s, s, inputs.shape[-1]).mean(dim=1)
    return x * torch.rsqrt(torch.clamp((x * x).sum(), min=1e-12))
