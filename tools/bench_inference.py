#!/usr/bin/env python
"""Serving-side numbers: free-running Tacotron2 decode (decoder only, random weights: the stop token never fires reliably, so the
step cap sets the length) and WaveGlow vocoding of the resulting frames.  python tools/bench_inference.py [--steps N]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--cap", type=int, default=400)
args = ap.parse_args()
from multi_speaker_tts_b200 import synthetic as S
from multi_speaker_tts_b200.decoder import decoder_forward

dev = torch.device("cuda:0")
w = S.init_decoder_weights(0)
w['projection/bias'][80] = -50.0   # keep every row running to the cap
wd = {k: v.to(dev) for k, v in w.items()}
for mode, B in [(m, B) for m in ("bf16x3", "fp32") for B in (1, 4, 16, 32)]:
    b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(B, 64, args.cap, seed=3).items()}
    for _ in range(2):
        decoder_forward(wd, b['memory'], b['text_len'], None, None, b['prenet_mask'], None, is_training=False, n_steps=args.cap + 1, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lin, _, _, _ = decoder_forward(wd, b['memory'], b['text_len'], None, None, b['prenet_mask'], None, is_training=False,
                                   n_steps=args.cap + 1, mode=mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    T = lin.shape[1]
    print(json.dumps({"metric": "free-running decode (%s persistent kernel)" % ("tcgen05 bf16x3" if mode == "bf16x3" else "fp32 SIMT"), "B": B, "Te": 64, "steps": T, "ms": ms,
                      "us_per_step": ms * 1e3 / T, "frames_per_s": B * T / (ms * 1e-3),
                      "x_realtime_12.5ms_frames": B * T * 0.0125 / (ms * 1e-3)}))
