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


def waveglow_case():
    """tiny WaveGlow forward: N=1, 192 samples, orthogonal 1x1 kernels, small random end conv (seed 3)"""
    from oracle import waveglow_oracle as W
    raws, upk, upb = W.init_waveglow(3, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    flows = [W.effective_params(r) for r in raws]
    audio, mel = W.synthetic_batch(1, 8 * 24, 2, seed=77)
    a, m = W.restructure_train_data(audio, mel, upk, upb)
    z, ls, ld = W.glow_train(a, m, flows)
    l = W.glow_loss(z, ls, ld)
    return {"z": z.numpy(), "mel640_checksum": np.float64(m.double().sum().item()), "log_s_sum": np.float64(torch.stack(ls).sum().item()),
            "losses": np.array([float(x) for x in l])}


def audio_case():
    """Audio.melspectrogram of 0.4 s of seeded noise at the reference defaults and at the config-4 parameters"""
    from oracle import audio_oracle as A
    x = np.random.default_rng(5).uniform(-0.9, 0.9, 6400).astype(np.float32)
    return {"mel_ref_defaults": A.melspectrogram(x, 1025, 12.5, 50, 80, 16000, max_abs_value=4).astype(np.float32),
            "mel_config4": A.melspectrogram(x, 513, 256 / 22050 * 1000, 1024 / 22050 * 1000, 80, 22050, max_abs_value=4).astype(np.float32),
            "spec_ref_defaults_checksum": np.float64(A.spectrogram(x, 1025, 12.5, 50, 16000).sum())}


def inference_case():
    """free-running decode (Modules.py:212-237): B=2, Te=40, cap 60, stop bias -0.05 -> rows stop at steps 0 and 5"""
    w = S.init_decoder_weights(0, bias_scale=0.05)
    w['projection/bias'][80] = -0.05
    b = S.synthetic_decoder_batch(2, 40, 60, seed=3, ragged=True)
    lin, stop, al = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], None,
                                      is_training=False, max_steps=60)
    return {"linear": lin.numpy(), "stop": stop.numpy(), "align": al.numpy(), "steps": np.int32(lin.shape[1])}


def waveglow_grad_case():
    """gradients of the three WaveGlow loss terms (N=1, 320 samples, seed 0) w.r.t. a few raw variables, fp64 autograd"""
    from oracle import waveglow_oracle as W
    raws, upk, upb = W.init_waveglow(0, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    audio, mel = W.synthetic_batch(1, 8 * 40, 2)
    watch = {"f0/inv_w": raws[0]['inv_w'], "f5/in_3/v": raws[5]['in'][3]['v'], "f5/in_3/g": raws[5]['in'][3]['g'],
             "f11/cond_7/b": raws[11]['cond'][7]['b'], "f7/res_0/v": raws[7]['res'][0]['v'], "f2/end_w": raws[2]['end_w'],
             "f9/start/v": raws[9]['start']['v'], "up_kernel": upk}
    leaves = {}
    for k, t in watch.items():
        leaves[k] = t.double().clone().requires_grad_(True)
    raws[0]['inv_w'] = leaves["f0/inv_w"]; raws[5]['in'][3]['v'] = leaves["f5/in_3/v"]; raws[5]['in'][3]['g'] = leaves["f5/in_3/g"]
    raws[11]['cond'][7]['b'] = leaves["f11/cond_7/b"]; raws[7]['res'][0]['v'] = leaves["f7/res_0/v"]; raws[2]['end_w'] = leaves["f2/end_w"]
    raws[9]['start']['v'] = leaves["f9/start/v"]

    def dbl(d):
        return {k: (dbl(v) if isinstance(v, dict) else [dbl(x) for x in v] if isinstance(v, list) else
                    (v if v.dtype == torch.float64 else v.double())) for k, v in d.items()}
    flows = [W.effective_params(dbl(r)) for r in raws]
    a, m = W.restructure_train_data(audio.double(), mel.double(), leaves["up_kernel"], upb.double())
    z, ls, ld = W.glow_train(a, m, flows)
    losses = W.glow_loss(z, ls, [x.double() for x in ld])
    sum(losses).backward()
    out = {"losses": np.array([float(x) for x in losses])}
    for k, t in leaves.items():
        g = t.grad
        out["grad/" + k.replace('/', '.')] = g.numpy().astype(np.float32) if g.numel() <= 8192 else \
            np.array([g.norm().item(), g.sum().item(), g.flatten()[::997].abs().sum().item()])
    return out


def zlstm_case():
    """zoneout-LSTM sequence (dynamic_rnn semantics): B=3, T=7, In=512, H=256, ragged, training masks, both directions"""
    from oracle import tacotron2_oracle as T2
    g = torch.Generator().manual_seed(21)
    x = torch.randn(3, 7, 512, generator=g)
    lengths = torch.tensor([7, 4, 2], dtype=torch.int32)
    kernel = (torch.rand(768, 1024, generator=g) * 2 - 1) * 0.08
    bias = torch.randn(1024, generator=g) * 0.1
    masks = (torch.rand(7, 2, 3, 256, generator=g) < 0.9).float()
    fw = T2.dynamic_rnn(x, lengths, kernel, bias, True, masks)
    bw = T2.reverse_rows(T2.dynamic_rnn(T2.reverse_rows(x, lengths), lengths, kernel, bias, True, masks), lengths)
    return {"fw": fw.numpy(), "bw": bw.numpy()}


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in DECODER_CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **decoder_case(*cfg))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "waveglow_n1_t24.npz"), **waveglow_case())
    np.savez_compressed(os.path.join(HERE, "audio_mel.npz"), **audio_case())
    print("wrote waveglow / audio fixtures")
    np.savez_compressed(os.path.join(HERE, "decoder_inference_b2_te40.npz"), **inference_case())
    np.savez_compressed(os.path.join(HERE, "waveglow_grads_n1_t40.npz"), **waveglow_grad_case())
    np.savez_compressed(os.path.join(HERE, "zlstm_b3_t7.npz"), **zlstm_case())
    print("wrote inference / waveglow-gradient / zlstm fixtures")
