"""cProfile of the host side of one full MSTTS train step (config 5): where the Python time between kernel launches goes"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_speaker_tts_b200 import MSTTS_SV, Feeder  # noqa: E402

dev = torch.device("cuda:0")
feeder = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(16, 128, 800), rank=0)
model = MSTTS_SV.Tacotron2(is_Training=True, device=dev, feeder=feeder, process_group=None)
pat = feeder.Get_Train_Pattern()
for _ in range(3):
    model.Run_Train_Step(pat)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    model.Run_Train_Step(pat)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(40)
