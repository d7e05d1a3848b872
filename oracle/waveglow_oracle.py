"""CPU ORACLE (test infrastructure, not product code) -- WaveGlow affine-coupling / WN stack.

Straight-line PyTorch-CPU restatement of WaveGlow/Modules.py and WaveGlow/Inv1x1.py of the reference (TF1).  Only
``tests/``, ``__graft_entry__.smoke()`` and the cpu_baseline legs of the bench scripts may import it.

PARITY UNPINNED: the reference ships no tests or golden vectors and TensorFlow 1.x cannot run in this image.  The
restatement is pinned by cross-checks against independent implementations (explicit loops, torch.linalg) and by
invertibility / property tests in ``tests/test_waveglow_oracle_cpu.py`` plus fixtures under ``tests/golden``.

Reference lines followed (WaveGlow/Modules.py unless noted):
  * weight norm          :9-33   g * v / sqrt(max(sum_{k,in} v^2, 1e-5)) per output channel (tf.nn.l2_normalize)
  * upsample             :198-208 conv2d_transpose k=1024 stride=256 VALID, kernel [1,k,out,in], + bias
  * data restructuring   :135-195
  * affine coupling      :210-250 (forward clamps log_s at 8, reverse does not)
  * WaveNet              :252-327 (residual added to the GATED activation, quirk B-6; plain end conv)
  * flows / early output :329-371
  * loss                 :373-384
  * invertible 1x1       Inv1x1.py:9-32 (logdet = n*(log(det64(1e3 W) + 1e-6) - c*log 1e3), no abs)
Layouts: activations [N, T, C] (TF NWC); conv kernels [k, in, out]; dense/1x1 kernels [in, out].
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

FLOWS, GROUPS, EARLY_EVERY, EARLY_SIZE = 12, 8, 4, 2
UP_K, UP_S = 1024, 256
WN_LAYERS, WN_CH, WN_K = 8, 512, 3
MEL = 80


def flow_channels(f):
    """number of audio channels entering flow f (2 are split off before flows 4 and 8)"""
    return GROUPS - EARLY_SIZE * (f // EARLY_EVERY)


def weight_norm(v, g):
    """v [k, in, out], g [out] -> g * v / sqrt(max(sum over (k, in) of v^2, 1e-5))"""
    ss = (v * v).sum(dim=(0, 1), keepdim=True)
    return g * v * torch.rsqrt(torch.clamp(ss, min=1e-5))


def conv1d_same(x, w, bias, dilation=1):
    """x [N,T,Cin], w [k,Cin,Cout] (TF layout, cross-correlation), SAME padding"""
    k = w.shape[0]
    pad = dilation * (k - 1) // 2
    y = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), bias, padding=pad, dilation=dilation)
    return y.transpose(1, 2)


def upsample_mel(mel, kernel, bias):
    """mel [N,Tm,80], kernel [1024, out 80, in 80] (tf conv2d_transpose kernel [1,k,out,in] squeezed), bias [80]
    -> [N,(Tm-1)*256+1024,80]:  out[n, t*256+k, co] += mel[n,t,ci] * kernel[k,co,ci]"""
    w = kernel.permute(2, 1, 0)  # torch conv_transpose1d weight [in, out, k]
    y = F.conv_transpose1d(mel.transpose(1, 2), w, bias, stride=UP_S)
    return y.transpose(1, 2)


def restructure_train_data(audio, mel, up_kernel, up_bias):
    """audio [N,S], mel [N,Tm,80] -> audio [N,S/8,8], mel [N,S/8,640]"""
    N, S = audio.shape
    S8 = (S // GROUPS) * GROUPS
    a = audio[:, :S8]
    m = upsample_mel(mel, up_kernel, up_bias)
    assert m.shape[1] >= S8, "upsampled mel shorter than the audio (the reference's tf.slice would fail)"
    m = m[:, :S8]
    return a.reshape(N, S8 // GROUPS, GROUPS), m.reshape(N, S8 // GROUPS, GROUPS * MEL)


def inv1x1(x, W, reverse=False):
    """x [N,T,c], W [c,c]: y = x @ W (conv2d with a [1,1,c,c] kernel); logdet per Inv1x1.py:25-27"""
    if reverse:
        return x @ torch.linalg.inv(W)
    c = W.shape[0]
    det = torch.linalg.det((W * 1e3).double())   # Inv1x1.py:25: the scaling happens in the kernel's dtype, then the cast
    logdet = (torch.log(det + 1e-6)).float() - math.log(1e3) * c
    logdet = logdet * float(x.shape[0] * x.shape[1])  # tf.shape(inputs)[0:3] of the [N,1,T,c] tensor = N*1*T
    return x @ W, logdet


def wavenet(x0, mel, p):
    """p: dict with effective (weight-normalised) kernels: start_w [1,c/2,512], start_b, in_w[i] [3,512,1024], in_b[i],
    cond_w[i] [1,640,1024], cond_b[i], res_w[i] [1,512,1024|512], res_b[i], end_w [1,512,c], end_b"""
    h = conv1d_same(x0, p['start_w'], p['start_b'])
    out = 0
    for i in range(WN_LAYERS):
        a = conv1d_same(h, p['in_w'][i], p['in_b'][i], dilation=2 ** i) + conv1d_same(mel, p['cond_w'][i], p['cond_b'][i])
        g = torch.tanh(a[..., :WN_CH]) * torch.sigmoid(a[..., WN_CH:])
        rs = conv1d_same(g, p['res_w'][i], p['res_b'][i])
        if i < WN_LAYERS - 1:
            h = g + rs[..., :WN_CH]
            out = out + rs[..., WN_CH:]
        else:
            out = out + rs
    o = conv1d_same(out, p['end_w'], p['end_b'])
    half = o.shape[-1] // 2
    return o[..., :half], o[..., half:]


def affine_coupling(x, mel, p, reverse=False):
    if not reverse:
        x, logdet = inv1x1(x, p['inv_w'])
    half = x.shape[-1] // 2
    x0, x1 = x[..., :half], x[..., half:]
    log_s, b = wavenet(x0, mel, p)
    if not reverse:
        log_s = torch.clamp(log_s, max=8.0)
        x1 = torch.exp(log_s) * x1 + b
        return torch.cat([x0, x1], dim=-1), log_s.sum(), logdet
    x1 = (x1 - b) / torch.exp(log_s)
    return inv1x1(torch.cat([x0, x1], dim=-1), p['inv_w'], reverse=True)


def glow_train(audio, mel, flows):
    outs, log_s_list, logdet_list = [], [], []
    x = audio
    for f in range(FLOWS):
        if f % EARLY_EVERY == 0 and f > 0:
            outs.append(x[..., :EARLY_SIZE])
            x = x[..., EARLY_SIZE:]
        x, ls, ld = affine_coupling(x, mel, flows[f])
        log_s_list.append(ls)
        logdet_list.append(ld)
    outs.append(x)
    return torch.cat(outs, dim=-1), log_s_list, logdet_list


def glow_inference(z_last, mel, flows, early_noise, sigma=1.0):
    """z_last [N,T,4]; early_noise: dict {8: [N,T,2], 4: [N,T,2]} standing in for tf.random.normal (explicit so the CUDA
    path and the oracle consume identical bits)"""
    x = z_last
    for f in reversed(range(FLOWS)):
        x = affine_coupling(x, mel, flows[f], reverse=True)
        if f % EARLY_EVERY == 0 and f > 0:
            x = torch.cat([early_noise[f] * sigma, x], dim=-1)
    return x.reshape(x.shape[0], -1)


def glow_loss(z, log_s_list, logdet_list, sigma=1.0):
    n = float(z.numel())
    log_s_loss = -torch.stack(log_s_list).sum() / n
    logdet_loss = -torch.stack(logdet_list).sum() / n
    audio_loss = (z * z).sum() / (2 * sigma ** 2) / n
    return log_s_loss, logdet_loss, audio_loss


# ---- seeded parameters in the reference's variable layout (g, v) ------------------------------------------------
def _glorot(shape, gen):
    rf = 1
    for d in shape[:-2]:
        rf *= d
    lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).float() * lim


def init_flow_raw(f, gen, end_scale=0.0, g_mode="glorot", inv_mode="reference"):
    """Raw variables of flow f as the reference creates them: weight-normed convs have (g, v, bias), g and v from the same
    (glorot) initialiser (WaveGlow/Modules.py:17-30); end conv zero-init (:319-320) unless end_scale > 0; invertible 1x1
    kernel N(0,1) with the first column flipped when det < 0 (Inv1x1.py:13-15)."""
    c = flow_channels(f)
    half = c // 2

    def wn(shape):
        v = _glorot(shape, gen)
        if g_mode == "glorot":
            lim = math.sqrt(6.0 / (shape[-1] + shape[-1]))
            g = (torch.rand(shape[-1], generator=gen, dtype=torch.float64) * 2 - 1).float() * lim
        else:  # "unit": g = 1 gives O(1) activations, a better conditioned numerics test
            g = torch.ones(shape[-1])
        return {'g': g, 'v': v, 'b': (torch.rand(shape[-1], generator=gen, dtype=torch.float64) * 0.02 - 0.01).float()}

    raw = {'start': wn((1, half, WN_CH)), 'in': [], 'cond': [], 'res': []}
    for i in range(WN_LAYERS):
        raw['in'].append(wn((WN_K, WN_CH, 2 * WN_CH)))
        raw['cond'].append(wn((1, GROUPS * MEL, 2 * WN_CH)))
        raw['res'].append(wn((1, WN_CH, 2 * WN_CH if i < WN_LAYERS - 1 else WN_CH)))
    raw['end_w'] = (torch.randn((1, WN_CH, c), generator=gen, dtype=torch.float64) * end_scale).float()
    raw['end_b'] = (torch.randn((c,), generator=gen, dtype=torch.float64) * end_scale).float()
    W = torch.randn((c, c), generator=gen, dtype=torch.float64)
    if inv_mode == "orthogonal":  # well-conditioned variant for numerics tests: 12 chained N(0,1) kernels grow z ~1e4x
        W = torch.linalg.qr(W)[0]
    if torch.linalg.det(W) < 0:
        W[:, 0] *= -1
    raw['inv_w'] = W.float()
    return raw


def effective_params(raw):
    """weight-normalised kernels of one flow (what the per-step graph recomputes, :31-33)"""
    def eff(d):
        return weight_norm(d['v'], d['g'])
    return {
        'start_w': eff(raw['start']), 'start_b': raw['start']['b'],
        'in_w': [eff(d) for d in raw['in']], 'in_b': [d['b'] for d in raw['in']],
        'cond_w': [eff(d) for d in raw['cond']], 'cond_b': [d['b'] for d in raw['cond']],
        'res_w': [eff(d) for d in raw['res']], 'res_b': [d['b'] for d in raw['res']],
        'end_w': raw['end_w'], 'end_b': raw['end_b'], 'inv_w': raw['inv_w'],
    }


def init_waveglow(seed=0, end_scale=0.0, g_mode="glorot", inv_mode="reference"):
    gen = torch.Generator().manual_seed(seed)
    raws = [init_flow_raw(f, gen, end_scale, g_mode, inv_mode) for f in range(FLOWS)]
    up_kernel = torch.rand((UP_K, MEL, MEL), generator=gen, dtype=torch.float64).float() * 0.02  # U(0, 0.02), :205
    up_bias = torch.zeros(MEL)
    return raws, up_kernel, up_bias


def synthetic_batch(N, S, Tm, seed=1234):
    rng = np.random.default_rng(seed)
    audio = np.clip(rng.standard_normal((N, S)) * 0.3, -0.99, 0.99).astype(np.float32)
    mel = np.clip(rng.standard_normal((N, Tm, MEL)) * 1.5, -4, 4).astype(np.float32)
    return torch.from_numpy(audio), torch.from_numpy(mel)
