#!/usr/bin/env python
"""Decoder train step beyond the headline shape: longer texts (129 .. 256 positions: two clusters per row, 16 rows per launch) and
larger batches (row chunks), bf16x3 tcgen05 loops vs the fp32 SIMT loops.  python tools/bench_envelope.py [--steps 5]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_speaker_tts_b200 import synthetic as S
from multi_speaker_tts_b200.trainer import DecoderTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--fp32", action="store_true", help="also time the fp32 SIMT loops (slow)")
args = ap.parse_args()
dev = torch.device("cuda:0")
L = 800
for B, Te in ((32, 128), (16, 256), (32, 160), (32, 256), (64, 128), (64, 256)):
    for mode in (("bf16x3", "fp32") if args.fp32 and (B, Te) in ((32, 256), (32, 128)) else ("bf16x3",)):
        tr = DecoderTrainer(dev, mem_dim=768, mode=mode, seed=0)
        b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(B, Te, L, seed=1).items()}
        for _ in range(2):
            tr.train_step(b['memory'], b['text_len'], b['mel'], b['mel_len'], L + 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = tr.train_step(b['memory'], b['text_len'], b['mel'], b['mel_len'], L + 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"B": B, "text_len": Te, "mel_len": L, "mode": mode, "ms_per_step": ms, "frames_per_s": B * L / (ms * 1e-3),
                          "loss": [float(x) for x in loss.tolist()]}))
        del tr, b
        torch.cuda.empty_cache()
