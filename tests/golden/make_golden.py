"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here: TF1 absent).

Run from the repo root:  python tests/golden/make_golden.py
Inputs are regenerated from seeds by multi_speaker_tts_b200.synthetic; the fixture stores the oracle's
outputs (and loss / gradient checks) so that (a) oracle drift is caught on CPU and (b) the GPU path is
compared with committed numbers, not only with a live oracle run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import decoder_oracle as O  # noqa: E402
from multi_speaker_tts_b200 import synthetic as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

DECODER_CASES = {
    # name: (B, Te, L, ragged, seed)
    "decoder_b2_te12_l6": (2, 12, 6, False, 11),
    "decoder_b3_te20_l9_ragged": (3, 20, 9, True, 12),
}


def decoder_case(B, Te, L, ragged, seed):
    w = S.init_decoder_weights(0, bias_scale=0.05)
    b = S.synthetic_decoder_batch(B, Te, L, seed=seed, ragged=ragged)
    wd = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    mem = b['memory'].clone().requires_grad_(True)
    lin, stop, al = O.decoder_forward(wd, mem, b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], b['zone_mask'])
    ll, sl = O.decoder_loss(lin, stop, b['mel'], b['mel_len'])
    (ll + sl).backward()
    out = {"linear": lin.detach().numpy(), "stop": stop.detach().numpy(), "align": al.detach().numpy(),
           "linear_loss": np.float32(ll.item()), "stop_loss": np.float32(sl.item()),
           "d_memory": mem.grad.numpy()}
    for k, v in wd.items():
        out["grad/" + k.replace('/', '.')] = v.grad.numpy() if v.numel() <= 4096 else \
            np.array([v.grad.double().norm().item(), v.grad.double().sum().item()])
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in DECODER_CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **decoder_case(*cfg))
        print("wrote", name)
