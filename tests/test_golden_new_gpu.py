"""GPU path against the committed fixtures (no live oracle involved): free-running decode, WaveGlow gradients, zoneout-LSTM."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_free_running_decode_golden(cuda_dev):
    from multi_speaker_tts_b200 import synthetic as S
    from multi_speaker_tts_b200.decoder import decoder_forward
    g = np.load(os.path.join(GOLD, "decoder_inference_b2_te40.npz"))
    w = S.init_decoder_weights(0, bias_scale=0.05)
    w['projection/bias'][80] = -0.05
    b = S.synthetic_decoder_batch(2, 40, 60, seed=3, ragged=True)
    wd = {k: v.to(cuda_dev) for k, v in w.items()}
    lin, stop, align, _ = decoder_forward(wd, b['memory'].to(cuda_dev), b['text_len'].to(cuda_dev), None, None,
                                          b['prenet_mask'].to(cuda_dev), None, is_training=False, n_steps=61)
    assert lin.shape[1] == int(g["steps"])
    assert np.abs(lin.cpu().numpy() - g["linear"]).max() < 1e-3
    assert np.array_equal(stop.cpu().numpy() >= 0, g["stop"] >= 0)
    assert np.array_equal(align.cpu().numpy().argmax(-1), g["align"].argmax(-1))


def test_waveglow_gradients_golden(cuda_dev):
    from oracle import waveglow_oracle as W  # parameters only (seeded initialiser)
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    g = np.load(os.path.join(GOLD, "waveglow_grads_n1_t40.npz"))
    raws, upk, upb = W.init_waveglow(0, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    audio, mel = W.synthetic_batch(1, 8 * 40, 2)
    params = M.WaveGlowParams(raws, upk, upb, cuda_dev)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z, losses, grads, d_mel = M.Glow_Train_Backward(a, m, params)
    dk, db = M.Upsample_Mel_Backward(mel.to(cuda_dev), d_mel.reshape(1, 320, 80), params)
    assert np.allclose([float(x) for x in losses], g["losses"], rtol=1e-5, atol=1e-6)
    mine = {"grad/f0.inv_w": grads[0]['inv_w'], "grad/f5.in_3.g": grads[5]['in'][3]['g'], "grad/f11.cond_7.b": grads[11]['cond'][7]['b'],
            "grad/f2.end_w": grads[2]['end_w'], "grad/f9.start.v": grads[9]['start']['v']}
    for k, t in mine.items():
        ref = g[k]
        assert np.abs(t.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max(), k
    for k, t in (("grad/f5.in_3.v", grads[5]['in'][3]['v']), ("grad/f7.res_0.v", grads[7]['res'][0]['v']), ("grad/up_kernel", dk)):
        t = t.double().cpu()
        got = np.array([t.norm().item(), t.sum().item(), t.flatten()[::997].abs().sum().item()])
        assert np.allclose(got, g[k], rtol=2e-3, atol=1e-7), (k, got, g[k])


def test_zlstm_golden(cuda_dev):
    from multi_speaker_tts_b200 import Modules
    g = np.load(os.path.join(GOLD, "zlstm_b3_t7.npz"))
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(3, 7, 512, generator=gen)
    lengths = torch.tensor([7, 4, 2], dtype=torch.int32)
    kernel = (torch.rand(768, 1024, generator=gen) * 2 - 1) * 0.08
    bias = torch.randn(1024, generator=gen) * 0.1
    masks = (torch.rand(7, 2, 3, 256, generator=gen) < 0.9).float()
    d = cuda_dev
    for key, rev in (("fw", False), ("bw", True)):
        out, _ = Modules.zoneout_lstm_sequence(x.to(d), lengths.to(d), kernel.to(d), bias.to(d), True, 0.1, masks.to(d), reverse=rev)
        assert np.abs(out.cpu().numpy() - g[key]).max() < 1e-4, key
