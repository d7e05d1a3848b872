"""GPU parity of the tensor-core conv1d (csrc/conv1d.cu, tf.layers.conv1d 'same' semantics) against F.conv1d in fp64:
forward L_inf < 1e-4 relative, gradients w.r.t. input / kernel / bias within 1e-4 of max|ref|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,Cin,Cout,k", [(3, 37, 512, 512, 5), (2, 50, 80, 512, 5), (4, 21, 512, 80, 5), (1, 9, 16, 24, 3)])
def test_conv1d_matches_library_fp64(cuda_dev, B, T, Cin, Cout, k):
    from multi_speaker_tts_b200 import Modules
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, T, Cin, generator=g)
    w = torch.randn(k, Cin, Cout, generator=g) * (2.0 / (k * Cin)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    R = torch.randn(B, T, Cout, generator=g)
    xd, wd, bd = (t.to(cuda_dev).requires_grad_(True) for t in (x, w, b))
    y = Modules._conv1d_same(xd, wd, bd)
    (y * R.to(cuda_dev)).sum().backward()
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    ref = F.conv1d(x64.transpose(1, 2), w64.permute(2, 1, 0), b64, padding=k // 2).transpose(1, 2)
    (ref * R.double()).sum().backward()
    assert (y.detach().cpu().double() - ref.detach()).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    for name, mine, r in (("dx", xd.grad, x64.grad), ("dkernel", wd.grad, w64.grad), ("dbias", bd.grad, b64.grad)):
        err = (mine.cpu().double() - r).abs().max().item() / r.abs().max().item()
        assert err < 1e-4, (name, err)
