"""GPU parity of the WaveGlow reverse pass (C ABI) against torch.autograd through the fp64 CPU oracle: gradients of
L = -sum(log_s)/n - sum(logdet W)/n + sum(z^2)/(2 sigma^2 n) w.r.t. every raw variable (weight-norm g / v / bias, end
conv, invertible 1x1 kernels) and, through the conditioning, the up-sampling kernel.  Gate: 2e-3 of max|ref| per tensor."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _rel(a, b):
    return (a.double().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("N,S,Tm,seed", [(1, 8 * 40, 2, 0), (2, 8 * 301, 8, 1)])
def test_gradients_match_autograd(cuda_dev, N, S, Tm, seed):
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    raws, upk, upb = W.init_waveglow(seed, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    audio, mel = W.synthetic_batch(N, S, Tm)
    params = M.WaveGlowParams(raws, upk, upb, cuda_dev)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z, losses, grads, d_mel = M.Glow_Train_Backward(a, m, params)
    S8 = (S // 8) * 8
    dk, db = M.Upsample_Mel_Backward(mel.to(cuda_dev), d_mel.reshape(N, S8, 80), params)
    torch.cuda.synchronize()
    # ---- fp64 oracle with autograd over the raw variables ----
    leaves = []

    def leaf(t):
        t = t.double().clone().requires_grad_(True)
        leaves.append(t)
        return t

    raws64 = [{'start': {k: leaf(v) for k, v in r['start'].items()},
               'in': [{k: leaf(v) for k, v in x.items()} for x in r['in']],
               'cond': [{k: leaf(v) for k, v in x.items()} for x in r['cond']],
               'res': [{k: leaf(v) for k, v in x.items()} for x in r['res']],
               'end_w': leaf(r['end_w']), 'end_b': leaf(r['end_b']), 'inv_w': leaf(r['inv_w'])} for r in raws]
    upk64, upb64 = leaf(upk), leaf(upb)
    flows = [W.effective_params(r) for r in raws64]
    a_ref, m_ref = W.restructure_train_data(audio.double(), mel.double(), upk64, upb64)
    z_ref, ls, ld = W.glow_train(a_ref, m_ref, flows)
    l_ref = W.glow_loss(z_ref, ls, [x.double() for x in ld])
    sum(l_ref).backward()
    for got, ref in zip(losses, l_ref):
        assert abs(float(got) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    worst = {}
    for f in range(12):
        r64, g = raws64[f], grads[f]
        pairs = [('inv_w', g['inv_w'], r64['inv_w']), ('end_w', g['end_w'], r64['end_w']), ('end_b', g['end_b'], r64['end_b'])]
        for k in ('g', 'v', 'b'):
            pairs.append(('start/' + k, g['start'][k], r64['start'][k]))
            for i in range(8):
                for grp in ('in', 'cond', 'res'):
                    pairs.append(('%s_%d/%s' % (grp, i, k), g[grp][i][k], r64[grp][i][k]))
        for name, mine, ref in pairs:
            e = _rel(mine, ref.grad)
            key = name.split('_')[0] + '/' + name.split('/')[-1] if '/' in name else name
            worst[key] = max(worst.get(key, 0.0), e)
            assert e < TOL, "flow %d %s: rel err %.3e" % (f, name, e)
    print({k: "%.1e" % v for k, v in worst.items()})
    assert _rel(dk, upk64.grad) < TOL and _rel(db, upb64.grad) < TOL


def test_forward_values_unchanged_by_saving(cuda_dev):
    """the training forward (activations kept for the reverse pass; row-major stacked operands packed per product) and Glow_Train
    (fused-epilogue tcgen05 GEMMs over tiled operands) compute the same bf16x3 products in a different summation order"""
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    raws, upk, upb = W.init_waveglow(2, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    audio, mel = W.synthetic_batch(2, 8 * 120, 4)
    params = M.WaveGlowParams(raws, upk, upb, cuda_dev)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    z0, ls0, ld0, ss0 = M.Glow_Train(a, m, params)
    z1, losses, _, _ = M.Glow_Train_Backward(a, m, params, want_d_mel=False)
    assert (z0 - z1).abs().max().item() < 1e-4
    l0 = M.Glow_Loss(z0, ls0, ld0, ss0)
    for x, y in zip(l0, losses):
        assert abs(float(x) - float(y)) <= 5e-6 * max(1.0, abs(float(y)))   # fp32 summation-order noise of a 0.13-sized mean


def test_trainer_step_clip_and_adam(cuda_dev, capsys):
    """WaveGlow.WaveGlow.Run_Train_Step: flat gradient buffer == Glow_Train_Backward's gradients, global-norm clip 0.1 and the
    TF Adam update (eps 1e-8) reproduced on the host; Train() prints the reference's line."""
    from oracle import waveglow_oracle as W, decoder_oracle as D
    from multi_speaker_tts_b200.WaveGlow import WaveGlow as WG
    raws, upk, upb = W.init_waveglow(5, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    feeder = WG.Feeder(seed=3, batch_size=2, signal_length=8 * 96)
    model = WG.WaveGlow(device=cuda_dev, feeder=feeder, raws=raws, up_kernel=upk, up_bias=upb)
    assert model.n_params == 268294760  # 261.7 M (WN + 1x1) + 6.55 M (upsampling)
    p0 = model.flat_p.clone()
    r = model.Run_Train_Step(feeder.Get_Train_Pattern())
    g = model.flat_g.clone()
    gnorm = float(g.double().square().sum().sqrt())
    assert abs(gnorm - r['Global_Norm']) < 1e-6 * max(1.0, gnorm)
    scale = 0.1 / max(gnorm, 0.1)
    p_ref, m, v = p0.cpu().clone(), torch.zeros_like(p0).cpu(), torch.zeros_like(p0).cpu()
    D.tf_adam_step(p_ref, m, v, g.cpu() * scale, 1, WG.learning_rate(0), eps=1e-8)
    assert (model.flat_p.cpu() - p_ref).abs().max().item() < 2e-6
    assert r['Global_Step'] == 0 and abs(r['Learning_Rate'] - 1e-3) < 1e-12
    # the loop itself on the reference's own initialisation (glorot g keeps the first sign-like Adam steps harmless; the
    # g = 1 test initialisation above makes one lr-sized step on every weight move log_s by O(1) per flow)
    del model
    torch.cuda.empty_cache()
    model = WG.WaveGlow(device=cuda_dev, feeder=feeder, seed=1)
    model.Train(max_Steps=2)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('Time:')]
    assert len(lines) == 2
    for key in ('Global step: 1', 'Learning rate:', 'Log S Loss:', 'Log Det W Loss:', 'Audio Loss:'):
        assert key in lines[1]
    assert torch.isfinite(model.flat_p).all()
    out = model.Run_Inference(torch.randn(1, 3, 80))
    assert out['Audio'].shape == (1, (2 * 256 + 1024) // 8 * 8) and torch.isfinite(out['Audio']).all()
