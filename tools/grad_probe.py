#!/usr/bin/env python
"""Which dense product limits the full-size decoder gradients?  Computes the fp32-oracle gradients of BASELINE config 2 once, then
re-runs the CUDA train path in sub-processes with the precision level of individual front-end calls overridden
(MSTTS_GEMM_FORCE / MSTTS_GEMM_FORCE_SITES, csrc/tc_gemm.cu) and prints the error of every gradient tensor.
python tools/grad_probe.py [--mode bf16x3] ; child: python tools/grad_probe.py --child ref.pt"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from multi_speaker_tts_b200 import synthetic as S

B, TE, L = 32, 128, 800


def inputs():
    return S.init_decoder_weights(0, bias_scale=0.05), S.synthetic_decoder_batch(B, TE, L, seed=1234, ragged=True)


def child(path, mode):
    from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss
    ref = torch.load(path)
    w, b = inputs()
    dev = torch.device("cuda:0")
    T = int(b['mel_len'].max()) + 1
    wd = {k: v.to(dev) for k, v in w.items()}
    bd = {k: v.to(dev) for k, v in b.items()}
    lin, stop, align, st = decoder_forward(wd, bd['memory'], bd['text_len'], bd['mel'], bd['mel_len'], bd['prenet_mask'][:T].contiguous(),
                                           bd['zone_mask'][:T].contiguous(), is_training=True, n_steps=T, mode=mode)
    loss2, dlin, dstop = decoder_loss(lin, stop, bd['mel'], bd['mel_len'])
    grads, dmem = decoder_backward(st, wd, dlin, dstop)
    torch.cuda.synchronize()
    out = []
    for k, r in ref.items():
        x = dmem.cpu() if k == 'd_memory' else grads[k].cpu()
        out.append("%s=%.1e" % (k.replace('/kernel', '/k').replace('/bias', '/b'), (x - r).abs().max().item() / r.abs().max().item()))
    print(" ".join(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", default=None)
    ap.add_argument("--mode", default="bf16x3")
    ap.add_argument("--sites", default="")
    args = ap.parse_args()
    if args.child:
        return child(args.child, args.mode)
    from oracle import decoder_oracle as O
    w, b = inputs()
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    mem = b['memory'].clone().requires_grad_(True)
    lin, stop, al = O.decoder_forward(wr, mem, b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], b['zone_mask'])
    ll, sl = O.decoder_loss(lin, stop, b['mel'], b['mel_len'])
    (ll + sl).backward()
    ref = {k: v.grad for k, v in wr.items()}
    ref['d_memory'] = mem.grad
    path = "/tmp/grad_probe_ref.pt"
    torch.save(ref, path)

    def run(tag, env):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, "--mode", args.mode], env=e, capture_output=True, text=True)
        print("%-28s %s" % (tag, r.stdout.strip() or r.stderr[-400:]))
        return r

    r = run("as built", {"MSTTS_GEMM_TRACE": "1"})
    calls = [l for l in r.stderr.splitlines() if l.startswith("[mstts gemm")]
    print("\n".join(calls))
    run("all precise", {"MSTTS_GEMM_FORCE": "2"})
    run("all chained", {"MSTTS_GEMM_FORCE": "1"})
    sites = [int(x) for x in args.sites.split(",") if x] or list(range(len(calls)))
    for i in sites:
        run("site %d precise" % i, {"MSTTS_GEMM_FORCE": "2", "MSTTS_GEMM_FORCE_SITES": str(i)})


if __name__ == "__main__":
    main()
