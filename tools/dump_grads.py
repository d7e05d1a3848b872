#!/usr/bin/env python
"""Dump the config-2 decoder gradients + a few workspace regions of whichever build this file sits in (used for A/B runs of two
worktrees): python tools/dump_grads.py out.pt [--mode bf16x3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from multi_speaker_tts_b200 import synthetic as S, _lib
from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss

out_path = sys.argv[1]
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
B, TE, L = 32, 128, 800
w, b = S.init_decoder_weights(0, bias_scale=0.05), S.synthetic_decoder_batch(B, TE, L, seed=1234, ragged=True)
dev = torch.device("cuda:0")
T = int(b['mel_len'].max()) + 1
wd = {k: v.to(dev) for k, v in w.items()}
bd = {k: v.to(dev) for k, v in b.items()}
lin, stop, align, st = decoder_forward(wd, bd['memory'], bd['text_len'], bd['mel'], bd['mel_len'], bd['prenet_mask'][:T].contiguous(),
                                       bd['zone_mask'][:T].contiguous(), is_training=True, n_steps=T, mode=mode)
loss2, dlin, dstop = decoder_loss(lin, stop, bd['mel'], bd['mel_len'])
grads, dmem = decoder_backward(st, wd, dlin, dstop)
torch.cuda.synchronize()
lib = _lib.lib()
m = _lib.MODES[mode]


def reg(name, cols, rows=T * B):
    off = lib.mstts_decoder_ws_offset(name.encode(), B, TE, L, 768, T, m)
    if off == 2 ** 64 - 1:
        return None
    return st.ws[off:off + rows * cols * 4].view(torch.float32).view(rows, cols).cpu().clone()


blob = {"grads": {k: v.cpu() for k, v in grads.items()}, "dmem": dmem.cpu(), "lin": lin.cpu(), "stop": stop.cpu(), "dlin": dlin.cpu(),
        "dstop": dstop.cpu(), "loss": loss2.cpu()}
for name, cols in (("dG0", 4096), ("dG1", 4096), ("g0pre", 4096), ("m1", 1024), ("act0", 4096)):
    r = reg(name, cols)
    if r is not None:
        blob[name] = r[:, :256].clone() if name == "act0" else r
torch.save(blob, out_path)
print("saved", out_path)
