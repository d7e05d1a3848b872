"""GPU parity of the zoneout-LSTM sequence kernels (csrc/zlstm.cu, through the C ABI and Modules.zoneout_lstm_sequence)
against the row-by-row CPU oracle (tf.nn.dynamic_rnn semantics over ZoneoutLSTMCell.call): outputs L_inf < 1e-4 and, through
torch.autograd over the fp64 oracle, gradients w.r.t. inputs / kernel / bias within 1e-3 of max|ref|."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle(x, lengths, kernel, bias, training, masks, residual, reverse):
    from oracle import tacotron2_oracle as O
    xin = O.reverse_rows(x, lengths) if reverse else x
    y = O.dynamic_rnn(xin, lengths, kernel, bias, training, masks, residual=residual)
    return O.reverse_rows(y, lengths) if reverse else y


@pytest.mark.parametrize("B,T,In,training,residual,reverse", [
    (3, 11, 512, True, False, False),    # encoder forward direction, ragged
    (3, 11, 512, True, False, True),     # encoder backward direction
    (10, 7, 256, False, True, False),    # speaker-embedding cell: inference, residual wrapper, 2 clusters
    (9, 5, 256, True, True, True),
])
def test_sequence_matches_oracle(cuda_dev, B, T, In, training, residual, reverse):
    from multi_speaker_tts_b200 import Modules
    g = torch.Generator().manual_seed(B * 100 + T)
    H = 256
    x = torch.randn(B, T, In, generator=g)
    lengths = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
    lengths[0] = T
    kernel = (torch.rand(In + H, 4 * H, generator=g) * 2 - 1) * 0.08
    bias = torch.randn(4 * H, generator=g) * 0.1
    masks = (torch.rand(T, 2, B, H, generator=g) < 0.9).float() if training else None
    R = torch.randn(B, T, H, generator=g)
    # GPU
    xd = x.to(cuda_dev).requires_grad_(True)
    kd = kernel.to(cuda_dev).requires_grad_(True)
    bd = bias.to(cuda_dev).requires_grad_(True)
    out, _ = Modules.zoneout_lstm_sequence(xd, lengths.to(cuda_dev), kd, bd, training, 0.1, None if masks is None else masks.to(cuda_dev),
                                           residual=residual, reverse=reverse)
    (out * R.to(cuda_dev)).sum().backward()
    # fp64 oracle
    x64 = x.double().requires_grad_(True)
    k64 = kernel.double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    ref = _oracle(x64, lengths, k64, b64, training, None if masks is None else masks.double(), residual, reverse)
    (ref * R.double()).sum().backward()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() < 1e-4
    for b in range(B):
        assert (out[b, int(lengths[b]):] == 0).all()  # zero beyond the sequence length
    for name, mine, r in (("d inputs", xd.grad, x64.grad), ("d kernel", kd.grad, k64.grad), ("d bias", bd.grad, b64.grad)):
        err = (mine.cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
        print(name, "%.2e" % err)
        assert err < 1e-3, name


def test_unsupported_width_uses_library_path_and_abi_refuses(cuda_dev):
    """H != 256 is not built as a kernel: the module runs the step-by-step library path; the C ABI itself refuses loudly"""
    import ctypes as C
    from multi_speaker_tts_b200 import Modules, _lib
    x = torch.randn(2, 4, 16, device=cuda_dev)
    k = torch.randn(16 + 128, 512, device=cuda_dev) * 0.05
    out, _ = Modules.zoneout_lstm_sequence(x, torch.tensor([4, 2], device=cuda_dev), k, torch.zeros(512, device=cuda_dev), False, 0.1)
    assert out.shape == (2, 4, 128) and (out[1, 2:] == 0).all()
    rc = _lib.lib().mstts_zlstm_fwd(_lib.ptr(x), _lib.ptr(k), _lib.ptr(x), None, None, 2, 4, 128, 0, C.c_float(0.9), _lib.ptr(out), None,
                                    None, None, None)
    assert rc == -4
