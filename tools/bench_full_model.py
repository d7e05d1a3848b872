#!/usr/bin/env python
"""Full Tacotron2 train step through the reference surface (MSTTS_SV.Tacotron2.Run_Train_Step) at BASELINE config 2 shapes:
speaker-embedding net + encoder + fused decoder + postnet + losses + backward + TF Adam, host feed dict in, losses out.
python tools/bench_full_model.py [--steps K] [--B 32 --Te 128 --L 800]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--Te", type=int, default=128)
ap.add_argument("--L", type=int, default=800)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
from multi_speaker_tts_b200 import MSTTS_SV, Feeder
from multi_speaker_tts_b200.decoder import set_profiling, kernel_ms

dev = torch.device("cuda:0")
feeder = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(args.B, args.Te, args.L))
model = MSTTS_SV.Tacotron2(is_Training=True, device=dev, feeder=feeder)
pat = feeder.Get_Train_Pattern()
for _ in range(3):
    model.Run_Train_Step(pat)
torch.cuda.synchronize()
set_profiling(True)
t0 = time.perf_counter()
for _ in range(args.steps):
    r = model.Run_Train_Step(pat)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / args.steps * 1e3
f_ms, f_n = kernel_ms(0)
b_ms, b_n = kernel_ms(1)
out = {"metric": "Tacotron2 full-model train step (reference surface)", "ms_per_step": ms, "frames_per_s": args.B * args.L / (ms * 1e-3),
       "decoder_fwd_loop_ms": f_ms / max(f_n, 1), "decoder_bwd_loop_ms": b_ms / max(b_n, 1),
       "config": {"B": args.B, "Te": args.Te, "L": args.L}, "losses": {k: r[k] for k in ("Linear_Loss", "Postnet_Loss", "Stop_Loss")}}
print(json.dumps(out))
if args.profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model.Run_Train_Step(pat)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
