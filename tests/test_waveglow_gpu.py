"""GPU parity of the WaveGlow flows (through the C ABI) vs the CPU oracle on identical inputs and parameters.

Gates: z L_inf < 1e-3 (orthogonal 1x1 kernels keep |z| = O(1), so the absolute gate is meaningful), loss terms
rel. 1e-5; reverse direction reproduces the input (invertibility through the CUDA path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dev, N, S, Tm, end_scale=0.02, inv_mode="orthogonal", g_mode="unit", seed=0):
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    raws, upk, upb = W.init_waveglow(seed, end_scale=end_scale, g_mode=g_mode, inv_mode=inv_mode)
    flows = [W.effective_params(r) for r in raws]
    audio, mel = W.synthetic_batch(N, S, Tm)
    params = M.WaveGlowParams(raws, upk, upb, dev)
    return W, M, flows, params, (audio, mel), (upk, upb)


@pytest.mark.parametrize("N,S,Tm", [(1, 8 * 40, 2), (2, 8 * 333 + 5, 8)])
def test_glow_train_parity(cuda_dev, N, S, Tm):
    W, M, flows, params, (audio, mel), (upk, upb) = _setup(cuda_dev, N, S, Tm)
    a_ref, m_ref = W.restructure_train_data(audio, mel, upk, upb)
    z_ref, ls_ref, ld_ref = W.glow_train(a_ref, m_ref, flows)
    l_ref = W.glow_loss(z_ref, ls_ref, ld_ref)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    assert (m.cpu() - m_ref).abs().max() < 1e-4  # Upsample_Mel parity
    z, ls_sum, ld_list, ss = M.Glow_Train(a, m, params)
    losses = M.Glow_Loss(z, ls_sum, ld_list, ss)
    torch.cuda.synchronize()
    err = (z.cpu() - z_ref).abs().max().item()
    print("z Linf %.3e (|z|max %.2f)" % (err, z_ref.abs().max().item()))
    assert err < 1e-3
    for got, ref in zip(losses, l_ref):
        assert abs(float(got) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref))), (float(got), float(ref))


def test_reference_init_relative_parity(cuda_dev):
    """reference initialisation (N(0,1) 1x1 kernels, glorot g, zero end conv): |z| grows ~1e4, so compare relatively"""
    W, M, flows, params, (audio, mel), (upk, upb) = _setup(cuda_dev, 1, 8 * 64, 3, end_scale=0.0, inv_mode="reference", g_mode="glorot")
    a_ref, m_ref = W.restructure_train_data(audio, mel, upk, upb)
    z_ref, ls_ref, ld_ref = W.glow_train(a_ref, m_ref, flows)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z, ls_sum, ld_list, ss = M.Glow_Train(a, m, params)
    assert float(ls_sum) == 0.0
    assert (z.cpu() - z_ref).abs().max() <= 1e-4 * z_ref.abs().max()


def test_inference_inverts_training_direction(cuda_dev):
    W, M, flows, params, (audio, mel), (upk, upb) = _setup(cuda_dev, 2, 8 * 96, 3)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z, _, _, _ = M.Glow_Train(a, m, params)
    x = M.Glow_Inference(z[..., 4:].contiguous(), m, params, sigma=1.0,
                         early_noise={4: z[..., 0:2].contiguous(), 8: z[..., 2:4].contiguous()})
    assert (x.reshape(a.shape) - a).abs().max() < 2e-3
    # and against the oracle's reverse pass on the same z
    x_ref = W.glow_inference(z[..., 4:].cpu(), m.cpu(), flows, {4: z[..., 0:2].cpu(), 8: z[..., 2:4].cpu()})
    assert (x.cpu() - x_ref).abs().max() < 2e-3


def test_upsample_keep_out_of_range_is_refused(cuda_dev):
    W, M, flows, params, (audio, mel), _ = _setup(cuda_dev, 1, 8 * 40, 2)
    from multi_speaker_tts_b200._lib import MsttsError
    with pytest.raises(MsttsError):
        M.Upsample_Mel(mel.to(cuda_dev), params, keep=10 ** 6)


def test_against_committed_golden(cuda_dev):
    import os
    import numpy as np
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waveglow_n1_t24.npz"))
    raws, upk, upb = W.init_waveglow(3, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    params = M.WaveGlowParams(raws, upk, upb, cuda_dev)
    audio, mel = W.synthetic_batch(1, 8 * 24, 2, seed=77)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z, ls_sum, ld_list, ss = M.Glow_Train(a, m, params)
    assert np.abs(z.cpu().numpy() - g["z"]).max() < 1e-3
    got = [float(x) for x in M.Glow_Loss(z, ls_sum, ld_list, ss)]
    assert np.allclose(got, g["losses"], rtol=1e-5, atol=1e-6)
