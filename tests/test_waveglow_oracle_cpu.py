"""CPU tests of the WaveGlow oracle: cross-checks against explicit loops / independent implementations and the
flow-invertibility property (the reference has no tests; parity unpinned)."""
import math

import numpy as np
import torch

from oracle import waveglow_oracle as W


def test_weight_norm_matches_definition():
    torch.manual_seed(0)
    v = torch.randn(3, 5, 7)
    g = torch.randn(7)
    w = W.weight_norm(v, g)
    for o in range(7):
        ss = float((v[:, :, o] ** 2).sum())
        assert torch.allclose(w[:, :, o], g[o] * v[:, :, o] / math.sqrt(max(ss, 1e-5)), atol=1e-6)
    tiny = torch.full((1, 1, 2), 1e-4)
    w2 = W.weight_norm(tiny, torch.ones(2))  # epsilon clamps the SUM OF SQUARES (tf.nn.l2_normalize)
    assert torch.allclose(w2, tiny / math.sqrt(1e-5))


def test_dilated_conv_against_loops():
    torch.manual_seed(1)
    N, T, Ci, Co, d = 2, 11, 3, 4, 2
    x = torch.randn(N, T, Ci, dtype=torch.float64)
    w = torch.randn(3, Ci, Co, dtype=torch.float64)
    b = torch.randn(Co, dtype=torch.float64)
    y = W.conv1d_same(x, w, b, dilation=d)
    ref = torch.zeros(N, T, Co, dtype=torch.float64)
    for t in range(T):
        for k in range(3):
            s = t + (k - 1) * d
            if 0 <= s < T:
                ref[:, t] += x[:, s] @ w[k]
    assert torch.allclose(y, ref + b, atol=1e-12)


def test_upsample_against_loops():
    torch.manual_seed(2)
    N, Tm = 1, 3
    mel = torch.randn(N, Tm, 80, dtype=torch.float64)
    K = torch.randn(1024, 80, 80, dtype=torch.float64) * 0.01
    b = torch.randn(80, dtype=torch.float64)
    y = W.upsample_mel(mel, K, b)
    assert y.shape == (N, (Tm - 1) * 256 + 1024, 80)
    ref = torch.zeros_like(y)
    for t in range(Tm):
        ref[:, t * 256:t * 256 + 1024] += torch.einsum('ni,koi->nko', mel[:, t], K)
    assert torch.allclose(y, ref + b, atol=1e-10)


def test_inv1x1_logdet_and_inverse():
    torch.manual_seed(3)
    Wm = torch.randn(6, 6)
    if torch.linalg.det(Wm) < 0:
        Wm[:, 0] *= -1
    x = torch.randn(2, 5, 6)
    y, ld = W.inv1x1(x, Wm)
    ref = 2 * 5 * (math.log(float(torch.linalg.det(Wm.double() * 1e3)) + 1e-6) - 6 * math.log(1e3))
    assert abs(float(ld) - ref) < 1e-3 * max(1, abs(ref))
    assert torch.allclose(W.inv1x1(y, Wm, reverse=True), x, atol=1e-4)


def _small_model(seed=0, end_scale=0.05):
    raws, upk, upb = W.init_waveglow(seed, end_scale=end_scale, g_mode="unit", inv_mode="orthogonal")
    return [W.effective_params(r) for r in raws], raws, upk, upb


def test_flow_invertibility_and_loss():
    """Glow_Inference(Glow_Train(x)) == x when the early outputs are fed back as the 'noise' (log_s < 8)."""
    flows, raws, upk, upb = _small_model()
    audio, mel = W.synthetic_batch(1, 8 * 24, 2)
    a, m = W.restructure_train_data(audio, mel, upk, upb)
    assert a.shape == (1, 24, 8) and m.shape == (1, 24, 640)
    z, ls, ld = W.glow_train(a, m, flows)
    assert z.shape == a.shape and len(ls) == 12 and len(ld) == 12
    # channel bookkeeping: early outputs of flows 4 and 8 come first
    x = W.glow_inference(z[..., 4:], m, flows, {4: z[..., 0:2], 8: z[..., 2:4]})
    assert torch.allclose(x.reshape(a.shape), a, atol=2e-3), (x.reshape(a.shape) - a).abs().max()
    l1, l2, l3 = W.glow_loss(z, ls, ld)
    assert all(torch.isfinite(v) for v in (l1, l2, l3))
    assert abs(float(l3) - float((z ** 2).sum() / 2 / z.numel())) < 1e-6


def test_zero_end_conv_gives_identity_coupling():
    """reference initialisation: end conv zero => log_s = 0, b = 0, so a flow is just the 1x1 mix"""
    flows, raws, upk, upb = _small_model(end_scale=0.0)
    x = torch.randn(1, 9, 8)
    mel = torch.randn(1, 9, 640)
    y, ls, ld = W.affine_coupling(x, mel, flows[0])
    assert float(ls) == 0.0
    assert torch.allclose(y, x @ flows[0]['inv_w'], atol=1e-6)


def test_golden_fixture_reproduces():
    """oracle drift guard: tests/golden/waveglow_n1_t24.npz (made by tests/golden/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waveglow_n1_t24.npz"))
    raws, upk, upb = W.init_waveglow(3, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    flows = [W.effective_params(r) for r in raws]
    audio, mel = W.synthetic_batch(1, 8 * 24, 2, seed=77)
    a, m = W.restructure_train_data(audio, mel, upk, upb)
    z, ls, ld = W.glow_train(a, m, flows)
    assert np.abs(z.numpy() - g["z"]).max() < 1e-5
    assert abs(float(torch.stack(ls).sum()) - float(g["log_s_sum"])) < 1e-4
    assert np.allclose([float(x) for x in W.glow_loss(z, ls, ld)], g["losses"], rtol=1e-5, atol=1e-6)
