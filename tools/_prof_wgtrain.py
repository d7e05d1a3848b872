import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_speaker_tts_b200.WaveGlow import WaveGlow as WG
dev = torch.device("cuda:0")
feeder = WG.Feeder(seed=1, batch_size=8, signal_length=16000)
model = WG.WaveGlow(device=dev, feeder=feeder, seed=0)
pat = feeder.Get_Train_Pattern()
model.Run_Train_Step(pat)
torch.cuda.synchronize()
