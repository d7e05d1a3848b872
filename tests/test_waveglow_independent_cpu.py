"""A second, independently structured reading of the reference's WaveGlow graph, checked against oracle/waveglow_oracle.py.

Same purpose as tests/test_oracle_independent_cpu.py (the oracle cannot be pinned against TensorFlow in this image): the flow
stack is re-derived from WaveGlow/Modules.py and WaveGlow/Inv1x1.py with different building blocks -- channel-last matmuls over
explicitly shifted copies instead of ``F.conv1d`` (dilated k=3 conv, SAME), per-output-channel weight norm from column norms of
the flattened kernel, the transposed-conv up-sampling as an overlap-add of per-frame outer products, numpy's determinant /
inverse for the invertible 1x1 -- and must reproduce the oracle's z, log_s sums, log-determinants, loss terms and the reverse direction.

Reference lines re-read: Modules.py:9-33 (g * l2_normalize(v, axes 0,1,2, eps 1e-5)), :135-175 (crop, fold 8 samples into
channels), :198-208 (conv2d_transpose k 1024 stride 256, VALID), :210-250 (1x1 first, split halves, x1' = exp(min(log_s, 8)) x1 + b,
reverse without the min), :252-327 (start conv, 8 layers dilation 2^i, tanh * sigmoid, residual onto the GATED activation, skip sum,
last layer all skip, plain end conv split into log_s | b), :329-352 (2 channels leave before flows 4 and 8, concatenated in
front-to-back order), :354-371 (reverse order, noise concatenated IN FRONT), :373-384 (loss terms over tf.size(output));
Inv1x1.py:13-31 (y = x W, logdet = N T (log(det64(1e3 W) + 1e-6) - c log 1e3), reverse with inv(W))."""
import math

import numpy as np
import torch

CH = 512


def _wn(d):
    v, g = d['v'].double(), d['g'].double()
    cols = v.reshape(-1, v.shape[-1])                   # one column per output channel: the norm runs over (k, in)
    norm = torch.linalg.vector_norm(cols, dim=0)
    scale = g / torch.sqrt(torch.clamp(norm * norm, min=1e-5))
    return (cols * scale).reshape(v.shape), d['b'].double()


def _conv_k3(x, w, b, d):
    """x [N,T,C], w [3,C,Co]: y[t] = x[t-d] w0 + x[t] w1 + x[t+d] w2 (zeros outside), TF SAME with dilation d"""
    N, T, C = x.shape
    z = torch.zeros(N, d, C, dtype=x.dtype)
    xp = torch.cat([z, x, z], 1)
    return xp[:, 0:T] @ w[0] + xp[:, d:d + T] @ w[1] + xp[:, 2 * d:2 * d + T] @ w[2] + b


def _wavenet(x0, mel, raw):
    w, b = _wn(raw['start'])
    h = x0 @ w[0] + b
    skip = torch.zeros(x0.shape[0], x0.shape[1], CH, dtype=torch.float64)
    for i in range(8):
        wi, bi = _wn(raw['in'][i])
        wc, bc = _wn(raw['cond'][i])
        a = _conv_k3(h, wi, bi, 2 ** i) + (mel @ wc[0] + bc)
        g = torch.tanh(a[..., :CH]) * torch.sigmoid(a[..., CH:])
        wr, br = _wn(raw['res'][i])
        rs = g @ wr[0] + br
        if i < 7:
            h = g + rs[..., :CH]                      # the reference adds the residual to the gated activation
            skip = skip + rs[..., CH:]
        else:
            skip = skip + rs
    o = skip @ raw['end_w'].double()[0] + raw['end_b'].double()
    c2 = o.shape[-1] // 2
    return o[..., :c2], o[..., c2:]


def _flow_forward(x, mel, raw):
    W = raw['inv_w'].double().numpy()
    c = W.shape[0]
    logdet = (math.log(np.linalg.det(W * 1e3) + 1e-6) - c * math.log(1e3)) * x.shape[0] * x.shape[1]
    x = x @ torch.from_numpy(W)
    h = c // 2
    log_s, b = _wavenet(x[..., :h], mel, raw)
    log_s = torch.minimum(log_s, torch.tensor(8.0, dtype=torch.float64))
    return torch.cat([x[..., :h], torch.exp(log_s) * x[..., h:] + b], -1), float(log_s.sum()), logdet


def _flow_reverse(y, mel, raw):
    h = y.shape[-1] // 2
    log_s, b = _wavenet(y[..., :h], mel, raw)
    x = torch.cat([y[..., :h], (y[..., h:] - b) / torch.exp(log_s)], -1)
    return x @ torch.from_numpy(np.linalg.inv(raw['inv_w'].double().numpy()))


def _upsample(mel, kernel, bias):
    """conv2d_transpose, VALID: out[n, 256 t + k, co] += sum_ci mel[n, t, ci] kernel[k, co, ci]"""
    N, Tm, _ = mel.shape
    out = torch.zeros(N, (Tm - 1) * 256 + 1024, 80, dtype=torch.float64)
    for t in range(Tm):
        out[:, 256 * t:256 * t + 1024] += torch.einsum('ni,koi->nko', mel[:, t].double(), kernel.double())
    return out + bias.double()


def test_waveglow_oracle_matches_second_reading():
    from oracle import waveglow_oracle as W
    torch.manual_seed(0)
    N, S, Tm = 2, 8 * 37 + 3, 2                                  # 3 trailing samples are cropped (:135-142)
    raws, upk, upb = W.init_waveglow(5, end_scale=0.05, g_mode="glorot", inv_mode="orthogonal")
    upb = torch.randn(80) * 0.01                                  # a non-zero up-sampling bias
    audio, mel = W.synthetic_batch(N, S, Tm, seed=9)
    # ---- second reading ----
    S8 = S // 8 * 8
    up = _upsample(mel, upk, upb)[:, :S8]
    a2 = audio[:, :S8].double().reshape(N, S8 // 8, 8)
    m2 = up.reshape(N, S8 // 8, 640)
    x, outs, ls2, ld2 = a2, [], [], []
    for f in range(12):
        if f in (4, 8):
            outs.append(x[..., :2])
            x = x[..., 2:]
        x, ls, ld = _flow_forward(x, m2, raws[f])
        ls2.append(ls)
        ld2.append(ld)
    outs.append(x)
    z2 = torch.cat(outs, -1)
    n = z2.numel()
    loss2 = (-sum(ls2) / n, -sum(ld2) / n, float((z2 ** 2).sum()) / 2 / n)
    # reverse direction from the same z (the early outputs stand in for the noise)
    y = z2[..., 4:]
    for f in reversed(range(12)):
        y = _flow_reverse(y, m2, raws[f])
        if f == 8:
            y = torch.cat([z2[..., 2:4], y], -1)
        if f == 4:
            y = torch.cat([z2[..., 0:2], y], -1)
    # ---- oracle (fp64 parameters) ----
    raws64 = [{k: ([{kk: vv.double() for kk, vv in d.items()} for d in v] if isinstance(v, list) else
                   ({kk: vv.double() for kk, vv in v.items()} if isinstance(v, dict) else v.double())) for k, v in r.items()}
              for r in raws]
    flows = [W.effective_params(r) for r in raws64]
    a_ref, m_ref = W.restructure_train_data(audio.double(), mel.double(), upk.double(), upb.double())
    z_ref, ls_ref, ld_ref = W.glow_train(a_ref, m_ref, flows)
    l_ref = W.glow_loss(z_ref, ls_ref, ld_ref)
    x_ref = W.glow_inference(z_ref[..., 4:], m_ref, flows, {4: z_ref[..., 0:2], 8: z_ref[..., 2:4]})
    assert (m_ref - m2).abs().max() < 1e-12 and torch.equal(a_ref, a2)
    assert (z_ref - z2).abs().max() < 1e-9 * max(1.0, float(z_ref.abs().max()))
    for f in range(12):
        assert abs(float(ls_ref[f]) - ls2[f]) < 1e-8 * max(1.0, abs(ls2[f]))
        assert abs(float(ld_ref[f]) - ld2[f]) < 1e-3 * max(1.0, abs(ld2[f]))   # the oracle rounds log(det) to fp32 as the reference does
    for got, ref in zip(loss2, l_ref):
        assert abs(got - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert (x_ref - y.reshape(N, -1)).abs().max() < 1e-8
    assert (y.reshape(N, -1) - audio[:, :S8].double()).abs().max() < 1e-7      # and the flow inverts
