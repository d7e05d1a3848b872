"""GPU parity of the reverse pass: CUDA gradients (C ABI) vs torch.autograd through the fp64 CPU oracle.

Tolerance: every gradient tensor within 2e-4 * max|ref| + 1e-7 (fp32 kernels vs an fp64 reference; the
forward gate is the north_star's 1e-3)."""
import pytest
import torch

from multi_speaker_tts_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _oracle_grads(w, b, T, dtype=torch.float64):
    from oracle import decoder_oracle as O
    wd = {k: v.to(dtype).requires_grad_(True) for k, v in w.items()}
    mem = b['memory'].to(dtype).requires_grad_(True)
    lin, stop, al = O.decoder_forward(wd, mem, b['text_len'], b['mel'].to(dtype), b['mel_len'], b['prenet_mask'],
                                      b['zone_mask'])
    ll, sl = O.decoder_loss(lin, stop, b['mel'].to(dtype), b['mel_len'])
    (ll + sl).backward()
    return {k: v.grad for k, v in wd.items()}, mem.grad, (ll.item(), sl.item())


@pytest.mark.parametrize("B,Te,L,ragged", [(2, 32, 24, False), (3, 40, 17, True), (8, 50, 9, True), (32, 128, 5, True),
                                           (36, 24, 4, True), (40, 128, 5, True), (64, 60, 3, True), (65, 16, 2, True),
                                           (7, 160, 5, True), (16, 256, 6, True), (5, 224, 9, True), (3, 129, 4, True), (9, 140, 4, False),
                                           (40, 160, 3, True), (20, 256, 4, True),
                                           (3, 20, 1, False), (2, 9, 2, False)])   # shortest loops: T = 2 and T = 3
@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_decoder_gradients(cuda_dev, B, Te, L, ragged, mode):
    # B > 32 in bf16x3 mode: two / three row chunks through the one-tile tcgen05 loops, weight gradients summed
    from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss
    w = S.init_decoder_weights(0, bias_scale=0.05)
    b = S.synthetic_decoder_batch(B, Te, L, seed=B + Te, ragged=ragged)
    T = int(b['mel_len'].max()) + 1
    assert T == L + 1
    ref_g, ref_dmem, (rll, rsl) = _oracle_grads(w, b, T)
    dev = cuda_dev
    wd = {k: v.to(dev) for k, v in w.items()}
    bd = {k: v.to(dev) for k, v in b.items()}
    lin, stop, align, st = decoder_forward(wd, bd['memory'], bd['text_len'], bd['mel'], bd['mel_len'], bd['prenet_mask'],
                                           bd['zone_mask'], True, T, mode)
    loss2, dlin, dstop = decoder_loss(lin, stop, bd['mel'], bd['mel_len'])
    grads, dmem = decoder_backward(st, wd, dlin, dstop)
    torch.cuda.synchronize()
    l2 = loss2.cpu()
    assert abs(l2[0].item() - rll) < 1e-5 * max(1, abs(rll)) and abs(l2[1].item() - rsl) < 1e-5
    worst = 0.0
    for k, rg in list(ref_g.items()) + [('d_memory', ref_dmem)]:
        gg = (dmem if k == 'd_memory' else grads[k]).cpu().double()
        assert gg.shape == rg.shape, k
        assert torch.isfinite(gg).all(), k
        scale = rg.abs().max().item()
        err = (gg - rg).abs().max().item()
        rel = err / (scale + 1e-30)
        worst = max(worst, rel)
        print("%-24s max|ref| %.3e  err %.3e  rel %.2e" % (k, scale, err, rel))
        assert err <= 2e-4 * scale + 1e-7, (k, err, scale)
    print("worst rel err %.2e" % worst)


@pytest.mark.parametrize("D", [256, 512])
@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_other_memory_widths(cuda_dev, D, mode):
    """The reference's memory is 768 wide (encoder 512 + speaker embedding 256, MSTTS_SV.py:70-71); the kernels also take
    256 and 512 (a decoder without the speaker concat): forward outputs and every gradient against the oracle."""
    from oracle import decoder_oracle as O
    from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss
    B, Te, L = 5, 40, 7
    w = S.init_decoder_weights(0, mem_dim=D, bias_scale=0.05)
    b = S.synthetic_decoder_batch(B, Te, L, seed=D, ragged=True, mem_dim=D)
    T = L + 1
    ref_g, ref_dmem, (rll, rsl) = _oracle_grads(w, b, T)
    with torch.no_grad():
        rl, rs, ra = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], b['zone_mask'])
    wd = {k: v.to(cuda_dev) for k, v in w.items()}
    bd = {k: v.to(cuda_dev) for k, v in b.items()}
    lin, stop, align, st = decoder_forward(wd, bd['memory'], bd['text_len'], bd['mel'], bd['mel_len'], bd['prenet_mask'],
                                           bd['zone_mask'], True, T, mode)
    assert (lin.cpu() - rl).abs().max() < 1e-3 and (stop.cpu() - rs).abs().max() < 1e-3 and (align.cpu() - ra).abs().max() < 1e-3
    assert torch.equal(align.cpu().argmax(-1), ra.argmax(-1))
    loss2, dlin, dstop = decoder_loss(lin, stop, bd['mel'], bd['mel_len'])
    grads, dmem = decoder_backward(st, wd, dlin, dstop)
    torch.cuda.synchronize()
    for k, rg in list(ref_g.items()) + [('d_memory', ref_dmem)]:
        gg = (dmem if k == 'd_memory' else grads[k]).cpu().double()
        scale = rg.abs().max().item()
        assert (gg - rg).abs().max().item() <= 2e-4 * scale + 1e-7, (k, D, mode)
