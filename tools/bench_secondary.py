#!/usr/bin/env python
"""Secondary benches (BASELINE configs 3 and 4): WaveGlow forward + NLL and STFT + mel extraction on one B200.
Prints one JSON line per workload with the roofline that bounds it and the oracle timed on the host CPU
(bounded sample).  python tools/bench_secondary.py [--steps K] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def time_gpu(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_waveglow(args, pk, src, emit=True):
    from oracle import waveglow_oracle as W  # parameters + cpu baseline only
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    dev = torch.device("cuda:0")
    N, S, Tm = 8, 16000, 64
    raws, upk, upb = W.init_waveglow(0, end_scale=0.01, g_mode="unit", inv_mode="orthogonal")
    params = M.WaveGlowParams(raws, upk, upb, dev)
    audio, mel = W.synthetic_batch(N, S, Tm)
    ad, md = audio.to(dev), mel.to(dev)

    def step():
        a, m = M.Restructure_Train_Data(ad, md, params)
        z, ls, ld, ss = M.Glow_Train(a, m, params)
        return M.Glow_Loss(z, ls, ld, ss)

    ms = time_gpu(step, min(args.steps, 10))
    flops = 8.357e12  # 21.76 M MAC / position / flow, 16 000 positions, 12 flows
    ach = flops / (ms * 1e-3) / 1e12
    out = {"metric": "WaveGlow forward+NLL samples/s", "value": N * S / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "dtype": "bf16x3 (3 bf16 tensor-core GEMMs per product, fp32 accumulate)", "data": "synthetic",
           "config": {"workload": "WaveGlow forward+NLL, 12 flows, 8 WN layers x 512 ch, N=8 x 16000 samples (BASELINE config 3)"},
           "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": src,
                        "note": "algorithmic FLOPs (1x); the bf16x3 split executes 3x of them on the tensor pipe"}}
    # ---- synthesis direction (Glow_Inference): the same 12 flows inverted, 4 + 2 + 2 noise channels in, audio out ----
    try:
        zin = torch.randn(N, S // 8, 4, device=dev)
        noise = {f: torch.randn(N, S // 8, 2, device=dev) for f in (8, 4)}
        a0, m0 = M.Restructure_Train_Data(ad, md, params)
        ims = time_gpu(lambda: M.Glow_Inference(zin, m0, params, sigma=0.6, early_noise=noise), min(args.steps, 5))
        out["inference"] = {"ms_per_step": ims, "samples_per_s": N * S / (ims * 1e-3),
                            "note": "Glow_Inference on the up-sampled conditioning of the same batch (tcgen05 path)"}
    except Exception as e:
        out["inference"] = {"error": repr(e)}
    # ---- the same workload as a full training step (forward with saved activations, reverse pass, clip, TF Adam) ----
    try:
        from multi_speaker_tts_b200.WaveGlow import WaveGlow as WG
        del params
        torch.cuda.empty_cache()
        feeder = WG.Feeder(seed=1, batch_size=N, signal_length=S)
        model = WG.WaveGlow(device=dev, feeder=feeder, seed=0)
        pat = feeder.Get_Train_Pattern()
        for _ in range(2):
            model.Run_Train_Step(pat)
        torch.cuda.synchronize()
        nst = min(args.steps, 5)
        t0 = time.perf_counter()
        for _ in range(nst):
            model.Run_Train_Step(pat)
        torch.cuda.synchronize()
        tms = (time.perf_counter() - t0) / nst * 1e3
        tach = 3 * flops / (tms * 1e-3) / 1e12
        out["train_step"] = {"metric": "WaveGlow train step samples/s", "ms_per_step": tms, "value": N * S / (tms * 1e-3), "unit": "samples/s",
                             "roofline": {"bound": "tensor", "achieved": tach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                          "frac": tach / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": src,
                                          "note": "algorithmic FLOPs: fwd + bwd = 3x the forward (25.07 TFLOP); bf16x3 executes 3x of them"},
                             "note": "Restructure_Train_Data + flows with saved activations + reverse pass + clip + TF Adam through "
                                     "WaveGlow.Run_Train_Step; host pinned->device copies and the loss read inside the timed region"}
        del model
        torch.cuda.empty_cache()
    except Exception as e:  # keep the forward line even if the trainer cannot allocate
        out["train_step"] = {"error": repr(e)}
    if not args.no_cpu:
        flows = [W.effective_params(r) for r in raws]
        torch.set_num_threads(os.cpu_count() or 1)
        a, m = W.restructure_train_data(audio[:1, :8000], mel[:1, :32], upk, upb)
        t0 = time.perf_counter()
        z, ls, ld = W.glow_train(a, m, flows)
        W.glow_loss(z, ls, ld)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 8000 / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "N=1 x 8000 samples (1/16 of the workload), forward + NLL, %.1f s" % dt}
    if emit:
        print(json.dumps(out))
    return out


def bench_stft(args, pk, src, emit=True):
    from multi_speaker_tts_b200 import Audio as G
    dev = torch.device("cuda:0")
    B, S, n_fft, hop = 64, 220500, 1024, 256
    wav = (torch.rand(B, S, device=dev) * 1.98 - 0.99)
    frames = 1 + S // hop

    def step():
        return G.stft_features(wav, n_fft, hop, n_fft, 22050, 80, 4.0)

    ms = time_gpu(step, max(args.steps, 20))
    alg = 4 * B * S + 4 * B * frames * 80
    ach = alg / (ms * 1e-3) / 1e9
    out = {"metric": "STFT+mel HBM GB/s", "value": ach, "unit": "GB/s", "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "Audio.melspectrogram 64 x 10 s @ 22050 Hz, n_fft 1024 hop 256, 80 mel (BASELINE config 4)"},
           "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "traffic": None, "peak_source": src, "algorithmic_bytes": alg}}
    if not args.no_cpu:
        from oracle import audio_oracle as A
        x = wav[0].cpu().numpy()
        t0 = time.perf_counter()
        A.melspectrogram(x, 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (4 * S + 4 * frames * 80) / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                               "sample": "1 of the 64 waveforms, %.2f s" % dt}
    if emit:
        print(json.dumps(out))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    pk, src = peaks()
    bench_waveglow(args, pk, src)
    bench_stft(args, pk, src)
