"""GPU parity of the fused activation -> batch norm -> dropout kernels (csrc/act_bn_dropout.cu) against the library-op
composition of the same tf.layers semantics in fp64: outputs, moving statistics and gradients w.r.t. x / gamma / beta."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, gamma, beta, mm, mv, mask, act, training, rate):
    a = torch.relu(x) if act == 0 else torch.tanh(x)
    if training:
        flat = a.reshape(-1, a.shape[-1])
        mean = flat.mean(0)
        var = ((flat - mean) ** 2).mean(0)
        mm_new = mm * 0.99 + mean.detach() * 0.01
        mv_new = mv * 0.99 + var.detach() * 0.01
    else:
        mean, var, mm_new, mv_new = mm, mv, mm, mv
    y = (a - mean) / torch.sqrt(var + 1e-3) * gamma + beta
    if training:
        y = y / (1.0 - rate) * mask
    return y, mm_new, mv_new


@pytest.mark.parametrize("B,T,C,act,training", [(3, 37, 512, 0, True), (2, 50, 80, 1, True), (4, 21, 512, 1, False), (1, 9, 8, 0, True)])
def test_fused_matches_composition(cuda_dev, B, T, C, act, training):
    from multi_speaker_tts_b200 import Modules
    g = torch.Generator().manual_seed(C + T + act)
    x = torch.randn(B, T, C, generator=g)
    gamma = 1 + 0.2 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    mm = 0.1 * torch.randn(C, generator=g)
    mv = 0.5 + torch.rand(C, generator=g)
    mask = (torch.rand(B, T, C, generator=g) < 0.5).float()
    R = torch.randn(B, T, C, generator=g)
    d = cuda_dev
    v = {'p/gamma': gamma.to(d).requires_grad_(True), 'p/beta': beta.to(d).requires_grad_(True), 'p/moving_mean': mm.to(d).clone(),
         'p/moving_variance': mv.to(d).clone()}
    xd = x.to(d).requires_grad_(True)
    y = Modules._act_bn_dropout(xd, v, 'p', act, training, 0.5, mask.to(d))
    (y * R.to(d)).sum().backward()
    x64 = x.double().requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    b64 = beta.double().requires_grad_(True)
    yr, mmr, mvr = _ref(x64, g64, b64, mm.double(), mv.double(), mask.double(), act, training, 0.5)
    (yr * R.double()).sum().backward()
    assert (y.detach().cpu().double() - yr.detach()).abs().max().item() < 1e-4
    assert (v['p/moving_mean'].cpu().double() - mmr).abs().max().item() < 1e-5
    assert (v['p/moving_variance'].cpu().double() - mvr).abs().max().item() < 1e-5
    if training:
        for name, mine, r in (("dx", xd.grad, x64.grad), ("dgamma", v['p/gamma'].grad, g64.grad), ("dbeta", v['p/beta'].grad, b64.grad)):
            err = (mine.cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
            assert err < 1e-4, (name, err)
