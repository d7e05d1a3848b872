"""GPU parity at BASELINE.json's FULL sizes (configs 2, 3 and 4), through the C ABI, against the CPU oracle and
against the domain's size-independent properties.

The oracle finishes these sizes in seconds (decoder forward 801 steps x 32 rows: ~5 s on the GPU box's host cores;
WaveGlow N=8 x 16000: ~10 s; STFT 64 x 10 s: ~1 s), so the comparison is direct, not sampled.  Gates as in the
small-shape tests (north_star): outputs L_inf < 1e-3, stop decision bit-exact, alignment argmax bit-exact.

The bit-exact reference is the oracle run in fp64 (its own rounding cannot decide a comparison).  At 25 632 (row, step)
pairs a handful of attention rows hold two largest entries closer than fp32 rounding; the fp32 run of the SAME oracle code
shows which ones (where fp32 and fp64 oracles disagree, "the reference in fp32" is itself undecided).  Gate: the CUDA argmax /
stop decision must equal the fp64 oracle at EVERY position where the two oracle precisions agree, and the total number of
positions where CUDA differs from the fp64 oracle is printed and hard-bounded by a constant (expected 0)."""
import numpy as np
import pytest
import torch

from multi_speaker_tts_b200 import synthetic as S

pytestmark = pytest.mark.gpu
TOL = 1e-3
MAX_FLIPS = 3     # hard bound on decisions that differ from the fp64 oracle over all 25 632 positions (expected: 0)
B, TE, L = 32, 128, 800


def _decoder_inputs(ragged):
    w = S.init_decoder_weights(0, bias_scale=0.05)
    b = S.synthetic_decoder_batch(B, TE, L, seed=1234, ragged=ragged)
    return w, b


def _gpu_forward(w, b, dev, mode):
    from multi_speaker_tts_b200.decoder import decoder_forward
    T = int(b['mel_len'].max()) + 1
    wd = {k: v.to(dev) for k, v in w.items()}
    bd = {k: v.to(dev) for k, v in b.items()}
    lin, stop, align, st = decoder_forward(wd, bd['memory'], bd['text_len'], bd['mel'], bd['mel_len'],
                                           bd['prenet_mask'][:T].contiguous(), bd['zone_mask'][:T].contiguous(),
                                           is_training=True, n_steps=T, mode=mode)
    torch.cuda.synchronize()
    return wd, bd, lin, stop, align, st


@pytest.mark.parametrize("ragged", [False, True])
def test_decoder_config2_forward_vs_oracle(cuda_dev, ragged):
    """BASELINE config 2 (B=32, text_len=128, mel_len=800): both CUDA implementations against the oracle."""
    from oracle import decoder_oracle as O
    w, b = _decoder_inputs(ragged)
    with torch.no_grad():
        rl, rs, ra = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'],
                                       b['zone_mask'])
        w64 = {k: v.double() for k, v in w.items()}
        rl64, rs64, ra64 = O.decoder_forward(w64, b['memory'].double(), b['text_len'], b['mel'].double(), b['mel_len'],
                                             b['prenet_mask'], b['zone_mask'])
    arg64, stop64 = ra64.argmax(-1), rs64 >= 0
    agree = ra.argmax(-1) == arg64                        # positions the oracle decides the same way in fp32 and fp64
    stop_agree = (rs >= 0) == stop64
    print("oracle fp32 vs fp64: argmax undecided at %d of %d positions, stop at %d" %
          (int((~agree).sum()), agree.numel(), int((~stop_agree).sum())))
    assert int((~agree).sum()) <= MAX_FLIPS and int((~stop_agree).sum()) <= MAX_FLIPS, "the oracle itself is unstable"
    for mode in ("bf16x3", "fp32"):
        _, _, lin, stop, align, _ = _gpu_forward(w, b, cuda_dev, mode)
        gl, gs, ga = lin.cpu(), stop.cpu(), align.cpu()
        assert gl.shape == rl.shape == (B, L + 1, 80) and ga.shape == ra.shape == (B, L + 1, TE)
        assert torch.isfinite(gl).all() and torch.isfinite(gs).all() and torch.isfinite(ga).all()
        e = [(gl - rl64).abs().max().item(), (gs - rs64).abs().max().item(), (ga - ra64).abs().max().item()]
        same = ga.argmax(-1) == arg64
        stop_same = (gs >= 0) == stop64
        flips, stop_flips = int((~same).sum()), int((~stop_same).sum())
        print("%s ragged=%s: Linf vs fp64 oracle linear %.3e stop %.3e align %.3e; argmax flips %d, stop flips %d (of %d)"
              % (mode, ragged, e[0], e[1], e[2], flips, stop_flips, same.numel()))
        assert max(e) < TOL
        assert bool(same[agree].all()), "alignment argmax differs from the fp64 oracle where the fp32 oracle agrees with it"
        assert bool(stop_same[stop_agree].all()), "stop decision differs from the fp64 oracle where the fp32 oracle agrees with it"
        assert flips <= MAX_FLIPS and stop_flips <= MAX_FLIPS, (flips, stop_flips)
        # properties: rows sum to 1, exactly 0 beyond text_len
        assert (ga.sum(-1) - 1).abs().max() < 1e-5
        beyond = torch.arange(TE)[None, None, :] >= b['text_len'][:, None, None]
        assert (ga[beyond.expand_as(ga)] == 0).all()


def _ws_region(st, name, cols, mode):
    """[T*B, cols] view of a named region of the decoder workspace (mstts_decoder_ws_offset)"""
    from multi_speaker_tts_b200 import _lib
    Bn, Te, Ln, Dn, T = st.shape
    off = _lib.lib().mstts_decoder_ws_offset(name.encode(), Bn, Te, Ln, Dn, T, _lib.MODES[mode])
    assert off != 2 ** 64 - 1, name
    return st.ws[off:off + T * Bn * cols * 4].view(torch.float32).view(T, Bn, cols)


class _PinnedPrenet(object):
    """The oracle's prenet (Modules.py:239-255) with the ReLU decisions taken from the CUDA forward.  A ReLU's derivative is
    discontinuous at 0: any two correct fp32 implementations disagree on the sign of a few of the 13 M pre-activations
    (|z| below their rounding noise), and ONE such element moves the cancellation-heavy prenet weight gradients by 1e-3 of
    their max (measured on the oracle alone: 5e-7 relative noise on z -> 3.6e-3 on d prenet_0/kernel).  So the decisions
    are pinned, and the disagreements are counted and must all be near-ties."""

    def __init__(self, pos_h, pos_p):
        self.pos = (pos_h, pos_p)   # [T, B, 256] bool: CUDA forward's (output > 0) of the two layers
        self.t = 0
        self.flips = 0
        self.worst = 0.0

    def __call__(self, x, w, m0, m1):
        h = x
        for layer, m in enumerate((m0, m1)):
            z = h @ w['prenet_%d/kernel' % layer] + w['prenet_%d/bias' % layer]
            kept = m > 0
            cuda_pos = self.pos[layer][self.t].to(z.device)
            dis = kept & ((z > 0) != cuda_pos)
            if bool(dis.any()):
                self.flips += int(dis.sum())
                self.worst = max(self.worst, float(z.detach()[dis].abs().max()))
            gate = torch.where(kept, cuda_pos, z > 0).to(z.dtype)
            h = (z * gate / 0.5) * m
        self.t += 1
        return h


def test_decoder_config2_gradients(cuda_dev):
    """Full-size reverse pass: the tcgen05 (bf16x3) and the fp32 SIMT kernels are independent implementations and must
    agree on every gradient tensor within 2e-4 of its max; the fp32 oracle (torch.autograd over 801 steps, itself
    carrying fp32 rounding) within 1e-3 of max.  The two non-smooth points of the graph (L1 sign, prenet ReLU) are pinned
    to the CUDA forward's decisions and their disagreements counted (see _PinnedPrenet and the L1 note below)."""
    from oracle import decoder_oracle as O
    from multi_speaker_tts_b200.decoder import decoder_backward, decoder_loss
    w, b = _decoder_inputs(True)
    res, upstream = {}, None
    for mode in ("bf16x3", "fp32"):
        wd, bd, lin, stop, align, st = _gpu_forward(w, b, cuda_dev, mode)
        loss2, dlin, dstop = decoder_loss(lin, stop, bd['mel'], bd['mel_len'])
        if upstream is None:
            upstream = (dlin, dstop)     # both reverse passes differentiate the same upstream gradient (see the L1 note below)
        grads, dmem = decoder_backward(st, wd, upstream[0].to(cuda_dev), upstream[1].to(cuda_dev))
        torch.cuda.synchronize()
        res[mode] = ({k: v.cpu() for k, v in grads.items()}, dmem.cpu(), loss2.cpu(), upstream[0].cpu(), upstream[1].cpu())
        if mode == "bf16x3":
            relu_pos = ((_ws_region(st, "pre_h", 256, mode) > 0).cpu(), (_ws_region(st, "pre", 256, mode) > 0).cpu())
        del st, grads, dmem
        torch.cuda.empty_cache()
    (ga, ma, la, dlin_a, dstop_a), (gb, mb, lb, _, _) = res["bf16x3"], res["fp32"]
    assert (la - lb).abs().max() < 1e-5 * max(1.0, lb.abs().max().item())
    worst = 0.0
    for k in list(gb) + ['d_memory']:
        x, y = (ma, mb) if k == 'd_memory' else (ga[k], gb[k])
        assert torch.isfinite(x).all() and torch.isfinite(y).all(), k
        scale = y.abs().max().item()
        err = (x - y).abs().max().item()
        worst = max(worst, err / (scale + 1e-30))
        assert err <= 2e-4 * scale + 1e-7, (k, err, scale)
    print("bf16x3 vs fp32 kernels, worst rel err %.2e" % worst)
    # the oracle's own gradients (fp32 autograd)
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    mem = b['memory'].clone().requires_grad_(True)
    pinned, plain_prenet = _PinnedPrenet(*relu_pos), O.prenet
    O.prenet = pinned
    try:
        lin, stop, al = O.decoder_forward(wr, mem, b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], b['zone_mask'])
    finally:
        O.prenet = plain_prenet
    print("prenet ReLU decisions: %d of %d differ between the CUDA forward and the fp32 oracle, largest |z| among them %.2e"
          % (pinned.flips, 2 * relu_pos[0].numel(), pinned.worst))
    assert pinned.flips <= 2000 and pinned.worst < 1e-4
    ll, sl = O.decoder_loss(lin, stop, b['mel'], b['mel_len'])
    assert abs(ll.item() - la[0].item()) < 1e-5 * max(1, abs(ll.item())) and abs(sl.item() - la[1].item()) < 1e-5
    # The L1 term makes the loss gradient discontinuous: d|x|/dx = sign(lin - mel) flips wherever the CUDA forward (L_inf 5e-6
    # from the oracle) lands on the other side of a mel value.  Over 2 M elements that is a handful of positions, each moving
    # d_linear by 2/n -- enough to move the cancellation-heavy prenet gradients by 3e-3 (measured), so the two stages are
    # gated separately: (1) the loss kernel's d_linear / d_stop equal the oracle's autograd except at near-ties, which are
    # counted and bounded; (2) the reverse pass is compared with the oracle's backward of THE SAME upstream gradient.
    dlin_ref, dstop_ref = torch.autograd.grad(ll + sl, [lin, stop], retain_graph=True)
    tie = (lin.detach()[:, :-1] - b['mel']).abs() < 2e-5
    differs = (dlin_a - dlin_ref).abs() > 1e-9
    n_flip = int(differs.sum())
    print("d_linear: %d of %d elements differ from the oracle's autograd (L1 sign at near-ties)" % (n_flip, differs.numel()))
    assert n_flip <= 32 and bool(tie[differs[:, :-1]].all()) and not bool(differs[:, -1].any())
    assert (dstop_a - dstop_ref).abs().max() <= 1e-4 * dstop_ref.abs().max()
    torch.autograd.backward([lin, stop], [dlin_a, dstop_a])
    worst, bad = 0.0, []
    for k in list(gb) + ['d_memory']:
        ref = mem.grad if k == 'd_memory' else wr[k].grad
        x = ma if k == 'd_memory' else ga[k]
        scale = ref.abs().max().item()
        err = (x - ref).abs().max().item()
        worst = max(worst, err / (scale + 1e-30))
        print("%-24s max|ref| %.3e err %.3e rel %.2e" % (k, scale, err, err / (scale + 1e-30)))
        if err > 1e-3 * scale + 1e-7:
            bad.append((k, err, scale))
    print("bf16x3 kernel vs fp32 oracle autograd, worst rel err %.2e" % worst)
    assert not bad, bad


def test_waveglow_config3_vs_oracle_and_round_trip(cuda_dev):
    """BASELINE config 3 (N=8 x 16000 samples, 12 flows, 8 x 512-ch WN): z / loss terms against the oracle, then the
    encode -> decode round trip through Glow_Inference reproduces the audio."""
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    N, S_, Tm = 8, 16000, 64
    raws, upk, upb = W.init_waveglow(0, end_scale=0.02, g_mode="unit", inv_mode="orthogonal")
    flows = [W.effective_params(r) for r in raws]
    audio, mel = W.synthetic_batch(N, S_, Tm)
    params = M.WaveGlowParams(raws, upk, upb, cuda_dev)
    a, m = M.Restructure_Train_Data(audio.to(cuda_dev), mel.to(cuda_dev), params)
    assert a.shape == (N, S_ // 8, 8) and m.shape == (N, S_ // 8, 640)
    z, ls_sum, ld_list, ss = M.Glow_Train(a, m, params)
    losses = M.Glow_Loss(z, ls_sum, ld_list, ss)
    x = M.Glow_Inference(z[..., 4:].contiguous(), m, params, sigma=1.0,
                         early_noise={4: z[..., 0:2].contiguous(), 8: z[..., 2:4].contiguous()})
    torch.cuda.synchronize()
    rt = (x.reshape(a.shape) - a).abs().max().item()
    print("round trip |x - audio| max %.3e" % rt)
    assert rt < 2e-3
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        a_ref, m_ref = W.restructure_train_data(audio, mel, upk, upb)
        z_ref, ls_ref, ld_ref = W.glow_train(a_ref, m_ref, flows)
        l_ref = W.glow_loss(z_ref, ls_ref, ld_ref)
    assert (m.cpu() - m_ref).abs().max() < 1e-4
    err = (z.cpu() - z_ref).abs().max().item()
    print("z Linf %.3e (|z|max %.2f)" % (err, z_ref.abs().max().item()))
    assert err < TOL
    for got, ref in zip(losses, l_ref):
        assert abs(float(got) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref))), (float(got), float(ref))


def test_stft_config4_vs_oracle(cuda_dev):
    """BASELINE config 4 (64 x 10 s @ 22 050 Hz, n_fft 1024, hop 256, 80 mels): every waveform against the numpy
    oracle, plus two properties of the transform: batch rows are independent of their neighbours (row b of the batched
    launch equals the single-waveform launch bit for bit) and the result is deterministic."""
    from oracle import audio_oracle as A
    from multi_speaker_tts_b200 import Audio as G
    Bw, S_ = 64, 220500
    wav = np.random.default_rng(11).uniform(-0.99, 0.99, (Bw, S_)).astype(np.float32)
    wd = torch.from_numpy(wav).to(cuda_dev)
    shift, length = 256 / 22050 * 1000, 1024 / 22050 * 1000
    got = G.melspectrogram(wd, 513, shift, length, 80, 22050, max_abs_value=4)
    again = G.melspectrogram(wd, 513, shift, length, 80, 22050, max_abs_value=4)
    assert got.shape == (Bw, 80, 1 + S_ // 256)
    assert torch.equal(got, again)
    one = G.melspectrogram(wd[17:18].contiguous(), 513, shift, length, 80, 22050, max_abs_value=4)
    assert torch.equal(one[0], got[17])
    g = got.cpu().numpy()
    worst = 0.0
    for i in range(Bw):
        ref = A.melspectrogram(wav[i], 513, shift, length, 80, 22050, max_abs_value=4)
        worst = max(worst, float(np.abs(g[i] - ref).max()))
    print("mel Linf over 64 waveforms %.3e" % worst)
    assert worst < TOL
