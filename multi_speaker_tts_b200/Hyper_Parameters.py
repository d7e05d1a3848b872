"""Configuration tree of the reference, without TensorFlow.

Mirrors the attribute tree of the reference's ``Hyper_Parameters.py:4-241`` (every
``hp.X.Y.Z`` path and value the hot path reads) using a plain namespace type instead of
``tf.contrib.training.HParams``.  Built from one nested literal so the whole tree is visible
in one place; ``HParams`` keeps the ``.values()`` / attribute access the callers use.
"""


class HParams(object):
    """Attribute bag standing in for ``tf.contrib.training.HParams``."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, HParams(**v) if isinstance(v, dict) else v)

    def values(self):
        return {k: (v.values() if isinstance(v, HParams) else v) for k, v in vars(self).items()}

    def __repr__(self):
        return "HParams(%r)" % (self.values(),)


def _lr(initial, minimum, step, rate, start=None):
    d = {'Initial': initial, 'Min': minimum, 'Decay_Step': step, 'Decay_Rate': rate}
    if start is not None:
        d['Decay_Start_Step'] = start
    return d


def _adam(eps):
    return {'Beta1': 0.9, 'Beta2': 0.999, 'Epsilon': eps}


_TREE = {
    'Sound': {  # Hyper_Parameters.py:4-11
        'Sample_Rate': 16000, 'Spectrogram_Dim': 1025, 'Mel_Dim': 80, 'Max_Abs_Mel': 4,
        'Frame_Shift': 12.5, 'Frame_Length': 50,
    },
    'Encoder': {  # :13-30
        'Embedding': {'Token_Size': 42, 'Embedding_Size': 512},
        'Conv': {'Nums': 3, 'Kernel_Size': 5, 'Stride': 1, 'Channel': 512, 'Dropout_Rate': 0.5},
        'BiLSTM': {'Nums': 1, 'Cell_Size': 256, 'Zoneout_Rate': 0.1},
    },
    'Attention': {  # :32-40
        'Memory_Size': 128,
        'Conv': {'Kernel_Size': 31, 'Stride': 1, 'Channel': 32, 'Dropout_Rate': 0.5},
    },
    'Decoder': {  # :42-62
        'PreNet': {'Nums': 2, 'Size': 256, 'Use_Dropout': True, 'Dropout_Rate': 0.5},
        'LSTM': {'Nums': 2, 'Cell_Size': 1024, 'Zoneout_Rate': 0.1, 'Max_Inference_Length': 1000},
        'Conv': {'Nums': 5, 'Kernel_Size': 5, 'Stride': 1, 'Channel': 512, 'Dropout_Rate': 0.5},
    },
    'Train': {  # :64-91
        'Pre_Step': 0, 'Use_Pre_in_Main_Train': False,
        'Pattern_Path': 'E:/MSTTS_SV.Data', 'Metadata_File': 'METADATA.PICKLE',
        'Batch_Size': 32, 'Pattern_Sorting_by_Mel_Length': True,
        'Use_Wav_Length_Range': (500, 9000),
        'Pre_Train_Dataset_List': ['LJ'], 'Main_Train_Dataset_List': ['VCTK', 'TIMIT'],
        'Max_Pattern_Queue': 20,
        'Learning_Rate': _lr(1e-3, 1e-5, 10000, 0.5, start=0),
        'Weight_Regularization_Rate': 1e-6,
        'ADAM': _adam(1e-6),
        'Use_L1_Loss': True, 'Inference_Timing': 1000, 'Checkpoint_Save_Timing': 1000,
    },
    'Speaker_Embedding': {  # :93-131
        'Embedding_Size': 256,
        'LSTM': {'Nums': 3, 'Cell_Size': 256, 'Zoneout_Rate': 0.1, 'Use_Residual': True},
        'Inference': {'Sample_Nums': 5, 'Mel_Frame': 64, 'Overlap_Frame': 32,
                      'Max_Embedding_per_Batch': 128},
        'Checkpoint_Path': 'E:/Speaker_Embedding/Checkpoint',
        'Train': {
            'Pattern_Path': 'E:/Speaker_Embedding.Data', 'Metadata_File': 'METADATA.PICKLE',
            'Batch_Speaker': 32, 'Batch_per_Speaker': 10, 'Max_Pattern_Queue': 20,
            'Frame_Range': (140, 180), 'Loss_Calc_Method': 'Softmax',
            'Learning_Rate': _lr(1e-3, 1e-5, 10000, 0.5),
            'ADAM': _adam(1e-8),
            'Inference_Path': 'E:/MSTTS_Checkpoints/Speaker_Embedding_Checkpoint',
            'Inference_Timing': 1000, 'Checkpoint_Save_Timing': 1000,
        },
    },
    'Taco1_Mel_to_Spect': {  # :133-193 (vocoder alternative; accepted as a flag only)
        'ConvBank': {
            'Nums': 1, 'Max_Kernel_Size': 8, 'Stride': 1, 'Channel': 128,
            'Pooling': {'Size': 2, 'Stride': 1},
            'Projection1': {'Kernel_Size': 3, 'Stride': 1, 'Channel': 256},
            'Projection2': {'Kernel_Size': 3, 'Stride': 1, 'Channel': 80},
            'Dropout_Rate': 0.5,
        },
        'Highway': {'Nums': 4},
        'BiRNN': {'Nums': 1, 'Cell_Size': 128, 'Zoneout_Rate': 0.1},
        'Griffin_Lim_Iteration': 100,
        'Checkpoint_Path': 'E:/MSTTS_Checkpoints/Mel_to_Spect_Checkpoint',
        'Train': {
            'Pattern_Path': 'E:/Taco1_Mel_to_Spect.Data/', 'Metadata_File': 'METADATA.PICKLE',
            'Batch_Size': 128, 'Pattern_Sorting_by_Length': True, 'Max_Mel_Length': 1000,
            'Max_Pattern_Queue': 20,
            'Learning_Rate': _lr(1e-3, 1e-5, 100, 0.5, start=50000),
            'Weight_Regularization_Rate': 1e-6,
            'ADAM': _adam(1e-6),
            'Inference_Timing': 1000, 'Checkpoint_Save_Timing': 1000,
            'Inference': {'Path': 'E:/MtS(20190817)', 'Batch_Size': 128},
        },
    },
    'WaveGlow': {  # :197-236
        'Flows': 12, 'Groups': 8, 'Early_Every': 4, 'Early_Size': 2,
        'Upsample': {'Kernel_Size': 1024, 'Strides': 256},
        'WaveNet': {'Layers': 8, 'Channels': 512, 'Kernel_Size': 3},
        'Export_Sample_Rate': 22050,
        'Checkpoint_Path': 'E:/MSTTS_SV_for_WaveGlow_Server/Checkpoint',
        'Train': {
            'Pattern_Path': 'E:/Multi_Speaker_TTS.Raw_Data/VCTK/wav48',
            'Max_Signal_Length': 16000 // 2, 'Batch_Size': 4, 'Max_Pattern_Queue': 20,
            'Learning_Rate': _lr(1e-3, 1e-5, 100000, 0.5),
            'ADAM': _adam(1e-8),
            'Inference_Timing': 1000, 'Checkpoint_Save_Timing': 1000,
        },
        'Inference': {'Path': 'E:/WaveGlow', 'Mel_Split_Length': 40, 'Batch_Size': 4},
    },
}

for _k, _v in _TREE.items():
    globals()[_k] = HParams(**_v)
del _k, _v

Use_Vocoder = 'Taco1_Mel_to_Spect'  # :238  ('WaveGlow' or 'Taco1_Mel_to_Spect')
Inference_Path = 'E:/MSTTS_Test(20190827)'  # :240
Checkpoint_Path = 'E:/MSTTS_Checkpoints/Multi_Speaker_TTS_Checkpoint'  # :241
