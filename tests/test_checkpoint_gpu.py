"""GPU: Save -> Restore -> the next train step is identical to the uninterrupted run, for both trainers
(MSTTS_SV.py:30-40,244-251,287-289; WaveGlow/WaveGlow.py Saver): variables AND the Adam slots travel through the checkpoint."""
import os

import numpy as np
import pytest
import torch

from multi_speaker_tts_b200 import Feeder, Hyper_Parameters as hp

pytestmark = pytest.mark.gpu


def test_tacotron2_save_restore_next_step_identical(cuda_dev, tmp_path):
    from multi_speaker_tts_b200 import MSTTS_SV as M
    old = hp.Checkpoint_Path
    hp.Checkpoint_Path = str(tmp_path / "taco")
    try:
        shape = (2, 12, 20)
        a = M.Tacotron2(is_Training=True, device=cuda_dev, seed=3, feeder=Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=shape))
        feeds = [a.feeder.Get_Train_Pattern() for _ in range(3)]
        a.Run_Train_Step(feeds[0])
        a.Run_Train_Step(feeds[1])
        path = a.Save()
        assert os.path.basename(path) == 'CHECKPOINT-2.pt'          # Saver.save(..., global_step = Global_Step + 1)
        torch.manual_seed(123)                                       # encoder / postnet dropout bits come from torch's generator
        ra = a.Run_Train_Step(feeds[2])
        # A second model picks the checkpoint up and must reproduce step 3 bit for bit.  Same seed: the frozen speaker-embedding
        # net is not part of this checkpoint (the reference's Saver excludes it, MSTTS_SV.py:30-38) and the decoder's dropout /
        # zoneout stream is keyed by (seed, step).  Every trainable variable and Adam slot is scrambled first, so only Restore
        # can make the step agree.
        b = M.Tacotron2(is_Training=True, device=cuda_dev, seed=3, feeder=a.feeder)   # feed dicts are keyed by the feeder's placeholders
        b.flat_p.add_(0.05)
        b.flat_m.fill_(1.0)
        b.flat_v.fill_(1.0)
        b.Restore()
        assert b.global_Step == 2
        torch.manual_seed(123)
        rb = b.Run_Train_Step(feeds[2])
        for k in ('Global_Step', 'Learning_Rate', 'Linear_Loss', 'Postnet_Loss', 'Stop_Loss', 'Weight_Regularization_Loss'):
            assert ra[k] == rb[k], (k, ra[k], rb[k])
        for k in a.trainable:      # (the flat buffers also hold alignment gaps, which the scramble above touched)
            assert torch.equal(a.variables[k], b.variables[k]), k
        assert torch.equal(a.flat_m, b.flat_m) and torch.equal(a.flat_v, b.flat_v)
        # rotation: five newest files stay
        for _ in range(6):
            a.Run_Train_Step(feeds[0])
            a.Save()
        files = sorted(f for f in os.listdir(hp.Checkpoint_Path) if f.endswith('.pt'))
        assert len(files) == 5 and 'CHECKPOINT-9.pt' in files and 'CHECKPOINT-2.pt' not in files
    finally:
        hp.Checkpoint_Path = old


def test_waveglow_save_restore_keeps_adam_slots(cuda_dev, tmp_path):
    from multi_speaker_tts_b200.WaveGlow import WaveGlow as WG
    old = hp.WaveGlow.Checkpoint_Path
    hp.WaveGlow.Checkpoint_Path = str(tmp_path / "wg")
    try:
        a = WG.WaveGlow(device=cuda_dev, seed=1, feeder=WG.Feeder(seed=5, batch_size=1, signal_length=2048))
        feeds = [a.feeder.Get_Train_Pattern() for _ in range(3)]
        a.Run_Train_Step(feeds[0])
        a.Run_Train_Step(feeds[1])
        path = a.Save()
        assert os.path.basename(path) == 'CHECKPOINT-2.pt'
        ra = a.Run_Train_Step(feeds[2])
        b = WG.WaveGlow(device=cuda_dev, seed=77, feeder=a.feeder)
        b.Restore()
        assert b.global_Step == 2 and float(b.flat_v.abs().sum()) > 0.0   # the moments came back, not zeros
        rb = b.Run_Train_Step(feeds[2])
        for k in ('Log_S_Loss', 'Log_Det_W_Loss', 'Audio_Loss', 'Global_Norm'):
            assert abs(ra[k] - rb[k]) <= 1e-6 * max(1.0, abs(ra[k])), (k, ra[k], rb[k])
        assert (a.flat_p - b.flat_p).abs().max().item() < 1e-7
    finally:
        hp.WaveGlow.Checkpoint_Path = old
