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
    if os.environ.get("PROBE_CHAIN"):
        # re-derive the prenet gradient chain from the workspace's own dG0 in fp64 (torch on the GPU as the checker)
        from multi_speaker_tts_b200 import _lib
        lib = _lib.lib()
        Bn, Te, D = bd['memory'].shape
        m = _lib.MODES[mode]

        def reg(name, cols):
            off = lib.mstts_decoder_ws_offset(name.encode(), Bn, Te, bd['mel'].shape[1], D, T, m)
            return st.ws[off:off + T * Bn * cols * 4].view(torch.float32).view(T * Bn, cols)
        dG0, pre, pre_h, frames, dpre, dpre_h = (reg(n, c) for n, c in (("dG0", 4096), ("pre", 256), ("pre_h", 256), ("frames", 80),
                                                                         ("dpre", 256), ("dpre_h", 256)))
        K0, W1 = wd['cell_0/kernel'].double(), wd['prenet_1/kernel'].double()
        r_dpre = (dG0.double() @ K0[:256].t()) * 2 * (pre > 0)
        r_dpre_h = (r_dpre @ W1.t()) * 2 * (pre_h > 0)

        def cmp(tag, x, r):
            e = (x.double() - r).abs()
            rows = e.max(dim=1).values if e.dim() == 2 and e.shape[0] == T * Bn else None
            msg = "%-10s max|ref| %.3e err %.3e" % (tag, r.abs().max().item(), e.max().item())
            if rows is not None:
                bad = (rows > 1e-3 * r.abs().max()).nonzero().flatten()
                msg += " bad rows %d %s" % (bad.numel(), bad[:8].tolist())
            print(msg)
        dproj, dm1p, proj_tm, m1, ctx = (reg(n, c) for n, c in (("dproj_tm", 81), ("dm1_proj", 1024), ("proj_tm", 81), ("m1", 1024), ("ctx", D)))
        Wp = wd['projection/kernel'].double()
        cmp("dm1_proj", dm1p, dproj.double() @ Wp[:1024].t())
        off = lib.mstts_decoder_ws_offset(b"ctx", Bn, Te, bd['mel'].shape[1], D, T, m)
        ctx1 = st.ws[off + Bn * D * 4: off + (T + 1) * Bn * D * 4].view(torch.float32).view(T * Bn, D)
        cmp("proj_tm", proj_tm, m1.double() @ Wp[:1024] + ctx1.double() @ Wp[1024:])
        # upstream gradient as the loss kernel produced it vs torch
        lin64, stop64 = lin.double(), stop.double()
        cmp("dpre", dpre, r_dpre)
        cmp("dpre_h", dpre_h, r_dpre_h)
        cmp("dW1", grads['prenet_1/kernel'], pre_h.double().t() @ r_dpre)
        cmp("db1", grads['prenet_1/bias'], r_dpre.sum(0))
        cmp("dW0", grads['prenet_0/kernel'], frames.double().t() @ r_dpre_h)
        cmp("db0", grads['prenet_0/bias'], r_dpre_h.sum(0))
        cmp("dK0pre", grads['cell_0/kernel'][:256], pre.double().t() @ dG0.double())
        cmp("db_c0", grads['cell_0/bias'], dG0.double().sum(0))
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
    ap.add_argument("--quick", action="store_true")
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

    r = run("as built", {"MSTTS_GEMM_TRACE": "1", "PROBE_CHAIN": "1"})
    calls = [l for l in r.stderr.splitlines() if l.startswith("[mstts gemm")]
    print("\n".join(calls))
    if not args.quick:
        run("all precise", {"MSTTS_GEMM_FORCE": "2"})
        run("all chained", {"MSTTS_GEMM_FORCE": "1"})
    sites = [int(x) for x in args.sites.split(",") if x]
    for i in sites:
        run("site %d precise" % i, {"MSTTS_GEMM_FORCE": "2", "MSTTS_GEMM_FORCE_SITES": str(i)})


if __name__ == "__main__":
    main()
