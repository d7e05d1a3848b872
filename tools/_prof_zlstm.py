"""driver for ncu: encoder-sized zoneout-LSTM sequence forward + reverse (B=16, T=128, 512 -> 256 units)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_speaker_tts_b200 import Modules  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, T, In, H = 16, 128, 512, 256
x = torch.randn(B, T, In, generator=g).to(dev).requires_grad_(True)
k = ((torch.rand(In + H, 4 * H, generator=g) * 2 - 1) * 0.08).to(dev).requires_grad_(True)
b = (torch.randn(4 * H, generator=g) * 0.1).to(dev).requires_grad_(True)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    y, _ = Modules.zoneout_lstm_sequence(x, ln, k, b, True, 0.1)
    y.sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y, _ = Modules.zoneout_lstm_sequence(x, ln, k, b, True, 0.1)
    y.sum().backward()
e1.record()
torch.cuda.synchronize()
print("fwd + bwd sequence: %.3f ms" % (e0.elapsed_time(e1) / 10))
