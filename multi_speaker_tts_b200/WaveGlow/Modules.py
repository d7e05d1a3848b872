"""Host side of the WaveGlow hot path, keeping the reference's function names (WaveGlow/Modules.py):
``Restructure_Train_Data``, ``Restructure_Inference_Data``, ``Upsample_Mel``, ``Glow_Train``, ``Glow_Inference``,
``Glow_Loss``, ``Reshaped_Mel``.  Tensors are torch CUDA tensors, parameters a ``WaveGlowParams`` object holding the
reference's raw variables (g, v, bias per weight-normalised conv).  There is no CPU fallback.
"""
import ctypes as C
import math

import torch

from .. import _lib
from .. import Hyper_Parameters as hp

FLOWS, LAYERS = 12, 8


def flow_channels(f):
    return hp.WaveGlow.Groups - hp.WaveGlow.Early_Size * (f // hp.WaveGlow.Early_Every)


class WaveGlowParams(object):
    """raws: list (12) of dicts {'start','in'[8],'cond'[8],'res'[8]: {'g','v','b'}, 'end_w','end_b','inv_w'};
    up_kernel [1024,80,80] (k, out, in), up_bias [80]  -- fp32 CUDA tensors in the reference layouts."""

    def __init__(self, raws, up_kernel, up_bias, device):
        def dev(t):
            return t.to(device=device, dtype=torch.float32).contiguous()
        self.raws = []
        for r in raws:
            d = {'start': {k: dev(v) for k, v in r['start'].items()},
                 'in': [{k: dev(v) for k, v in x.items()} for x in r['in']],
                 'cond': [{k: dev(v) for k, v in x.items()} for x in r['cond']],
                 'res': [{k: dev(v) for k, v in x.items()} for x in r['res']],
                 'end_w': dev(r['end_w']), 'end_b': dev(r['end_b']), 'inv_w': dev(r['inv_w'])}
            self.raws.append(d)
        self.up_kernel, self.up_bias = dev(up_kernel), dev(up_bias)
        self.device = device

    def struct(self, inverse, raws=None, cls=None):
        """ctypes view of the variables (or of a gradient tree `raws`).  The forward-direction struct only holds pointers, which
        are stable for tensors that are updated in place (the trainer's flat-buffer views): it is built once per tree."""
        cache = getattr(self, '_struct_cache', None)
        if cache is None:
            cache = self._struct_cache = {}
        key = cls
        cacheable = not inverse and raws is None  # gradient trees are caller-owned and may be fresh per call: never cached
        if cacheable and key in cache:
            return cache[key]
        s = (cls or _lib.MsttsWaveGlowWeights)()
        keep = []
        for f, r in enumerate(raws if raws is not None else self.raws):
            W = r['inv_w']
            if inverse:
                W = torch.linalg.inv(W.double()).float().contiguous()  # tf.linalg.inv(kernel), Inv1x1.py:31
                keep.append(W)
            s.inv_w[f] = W.data_ptr()
            s.start_g[f], s.start_v[f], s.start_b[f] = (r['start'][k].data_ptr() for k in ('g', 'v', 'b'))
            for i in range(LAYERS):
                s.in_g[f][i], s.in_v[f][i], s.in_b[f][i] = (r['in'][i][k].data_ptr() for k in ('g', 'v', 'b'))
                s.cond_g[f][i], s.cond_v[f][i], s.cond_b[f][i] = (r['cond'][i][k].data_ptr() for k in ('g', 'v', 'b'))
                s.res_g[f][i], s.res_v[f][i], s.res_b[f][i] = (r['res'][i][k].data_ptr() for k in ('g', 'v', 'b'))
            s.end_w[f], s.end_b[f] = r['end_w'].data_ptr(), r['end_b'].data_ptr()
        if cacheable:
            cache[key] = (s, keep)
        return s, keep

    def log_det_w(self, n_positions):
        """per flow: N*T*(log(det64(1e3 W) + 1e-6) - c log 1e3)  (Inv1x1.py:25-27; float64 det, no abs)"""
        out = []
        for r in self.raws:
            W = r['inv_w']
            c = W.shape[0]
            det = torch.linalg.det((W * 1e3).double())   # scaled in fp32, then cast (Inv1x1.py:25)
            out.append(((torch.log(det + 1e-6)).float() - math.log(1e3) * c) * float(n_positions))
        return out


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def Upsample_Mel(inputs, params, keep=None):
    """[N,Tm,80] -> [N,(Tm-1)*256+1024,80] (or only the first `keep` frames)"""
    assert inputs.is_cuda, "no CPU fallback"
    N, Tm, _ = inputs.shape
    L = (Tm - 1) * hp.WaveGlow.Upsample.Strides + hp.WaveGlow.Upsample.Kernel_Size
    keep = L if keep is None else keep
    x = inputs.contiguous().float()
    out = torch.empty(N, keep, hp.Sound.Mel_Dim, device=x.device)
    ws = torch.empty(_lib.lib().mstts_upsample_mel_workspace_bytes(N, Tm), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        rc = _lib.lib().mstts_upsample_mel(_lib.ptr(x), _lib.ptr(params.up_kernel), _lib.ptr(params.up_bias), N, Tm, keep,
                                           _lib.ptr(out), C.c_void_p(ws.data_ptr()), ws.numel(), _stream(x))
    _lib.check(rc, "mstts_upsample_mel")
    return out


def Restructure_Train_Data(audios, mels, params):
    G = hp.WaveGlow.Groups
    N, S = audios.shape
    S8 = (S // G) * G
    a = audios[:, :S8].contiguous().float()
    m = Upsample_Mel(mels, params, keep=S8)
    return a.reshape(N, S8 // G, G), m.reshape(N, S8 // G, G * hp.Sound.Mel_Dim)


def Restructure_Inference_Data(mels, params, generator=None):
    G = hp.WaveGlow.Groups
    m = Upsample_Mel(mels, params)
    N, L, _ = m.shape
    m = m[:, :(L // G) * G].contiguous().reshape(N, L // G, G * hp.Sound.Mel_Dim)
    nch = G - (math.ceil(hp.WaveGlow.Flows / hp.WaveGlow.Early_Every) - 1) * hp.WaveGlow.Early_Size
    z = torch.randn(N, L // G, nch, device=m.device, generator=generator)
    return z, m


def _run_flows(params, x, mel, direction, early_noise=None):
    lib = _lib.lib()
    N, T, _ = x.shape
    dev = x.device
    x = x.contiguous().float()
    mel = mel.contiguous().float()
    assert mel.shape == (N, T, hp.WaveGlow.Groups * hp.Sound.Mel_Dim)
    nbytes = lib.mstts_waveglow_workspace_bytes(N, T)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    out = torch.empty(N, T, hp.WaveGlow.Groups, device=dev)
    sums = torch.zeros(2, device=dev, dtype=torch.float64)
    wstruct, keep = params.struct(inverse=(direction == 1))
    noise_arr = None
    if direction == 1:
        noise = [early_noise[8].contiguous().float(), early_noise[4].contiguous().float()]
        noise_arr = (C.c_void_p * 2)(noise[0].data_ptr(), noise[1].data_ptr())
    with torch.cuda.device(dev):
        rc = lib.mstts_waveglow_flows(C.byref(wstruct), _lib.ptr(x), _lib.ptr(mel), N, T, direction, noise_arr, _lib.ptr(out),
                                      _lib.ptr(sums), C.c_void_p(ws.data_ptr()), ws.numel(), _stream(x))
    _lib.check(rc, "mstts_waveglow_flows")
    return out, sums


def Glow_Train(audio_Tensor, mel_Tensor, params):
    """-> (z [N,T,8], sum of log_s over all flows (device scalar), list of per-flow log|det W| terms)"""
    z, sums = _run_flows(params, audio_Tensor, mel_Tensor, 0)
    return z, sums[0], params.log_det_w(audio_Tensor.shape[0] * audio_Tensor.shape[1]), sums[1]


def Glow_Inference(audio_Tensor, mel_Tensor, params, sigma=1.0, early_noise=None, generator=None):
    N, T, _ = audio_Tensor.shape
    if early_noise is None:
        early_noise = {f: torch.randn(N, T, hp.WaveGlow.Early_Size, device=audio_Tensor.device, generator=generator) for f in (8, 4)}
    noise = {f: v * sigma for f, v in early_noise.items()}
    x, _ = _run_flows(params, audio_Tensor, mel_Tensor, 1, noise)
    return x.reshape(N, -1)


def Glow_Loss(output_Audio_Tensor, log_S_sum, log_Det_W_List, sum_sq=None, sigma=1.0):
    n = float(output_Audio_Tensor.numel())
    log_S_Loss = -log_S_sum / n
    log_Det_W_Loss = -torch.stack([x.double() for x in log_Det_W_List]).sum() / n
    ss = sum_sq if sum_sq is not None else (output_Audio_Tensor.double() ** 2).sum()
    audio_Loss = ss / (2 * sigma ** 2) / n
    return log_S_Loss, log_Det_W_Loss, audio_Loss


def Reshaped_Mel(mel_Tensor):
    """pad the time axis to a multiple of Mel_Split_Length and fold the chunks into the batch (Modules.py:416-438)"""
    L = hp.WaveGlow.Inference.Mel_Split_Length
    B, T, D = mel_Tensor.shape
    pad = (L - T % L) % L
    if pad:
        mel_Tensor = torch.cat([mel_Tensor, mel_Tensor.new_zeros(B, pad, D)], dim=1)
    return mel_Tensor.reshape(B * (mel_Tensor.shape[1] // L), L, D)


# ---- training: forward with saved activations + reverse pass (WaveGlow/WaveGlow.py:48-70) ---------------------------------
_TRAIN_WS = {}


def zeros_like_raws(raws):
    """gradient holders with the structure of ``WaveGlowParams.raws``"""
    def z(t):
        return torch.zeros_like(t)
    return [{'start': {k: z(v) for k, v in r['start'].items()},
             'in': [{k: z(v) for k, v in x.items()} for x in r['in']],
             'cond': [{k: z(v) for k, v in x.items()} for x in r['cond']],
             'res': [{k: z(v) for k, v in x.items()} for x in r['res']],
             'end_w': z(r['end_w']), 'end_b': z(r['end_b']), 'inv_w': z(r['inv_w'])} for r in raws]


def Glow_Train_Backward(audio_Tensor, mel_Tensor, params, sigma=1.0, grads=None, want_d_mel=True):
    """Glow_Train + Glow_Loss + their gradients.  Returns (z, (log_S_Loss, log_Det_W_Loss, audio_Loss), grads, d_mel) where
    ``grads`` mirrors ``params.raws`` (gradient of the sum of the three loss terms w.r.t. every raw variable g / v / bias /
    end conv / invertible 1x1 kernel) and d_mel [N,T,640] is the gradient w.r.t. the folded up-sampled conditioning."""
    lib = _lib.lib()
    x = audio_Tensor.contiguous().float()
    mel = mel_Tensor.contiguous().float()
    assert x.is_cuda and mel.is_cuda, "no CPU fallback"
    N, T, _ = x.shape
    dev = x.device
    nbytes = lib.mstts_waveglow_train_workspace_bytes(N, T)
    ws = _TRAIN_WS.get(dev)
    if ws is None or ws.numel() < nbytes:
        _TRAIN_WS[dev] = ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    z = torch.empty(N, T, hp.WaveGlow.Groups, device=dev)
    sums = torch.zeros(2, device=dev, dtype=torch.float64)
    wstruct, keep = params.struct(inverse=False)
    if grads is None:
        grads = zeros_like_raws(params.raws)
    gstruct, _ = params.struct(inverse=False, raws=grads, cls=_lib.MsttsWaveGlowGrads)
    d_mel = torch.empty(N, T, hp.WaveGlow.Groups * hp.Sound.Mel_Dim, device=dev) if want_d_mel else None
    with torch.cuda.device(dev):
        rc = lib.mstts_waveglow_train_fwd(C.byref(wstruct), _lib.ptr(x), _lib.ptr(mel), N, T, _lib.ptr(z), _lib.ptr(sums),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), _stream(x))
        _lib.check(rc, "mstts_waveglow_train_fwd")
        rc = lib.mstts_waveglow_train_bwd(C.byref(wstruct), C.byref(gstruct), _lib.ptr(z), N, T, float(sigma), _lib.ptr(d_mel),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), _stream(x))
        _lib.check(rc, "mstts_waveglow_train_bwd")
    n = float(z.numel())
    # log-det term of the invertible 1x1 kernels: host-side c x c math in float64 (Inv1x1.py:25-27), forward and gradient
    log_dets = []
    for r, gr in zip(params.raws, grads):
        W = r['inv_w'].double()
        c = W.shape[0]
        det3 = torch.linalg.det((r['inv_w'] * 1e3).double())   # scaled in fp32, then cast (Inv1x1.py:25)
        log_dets.append(((torch.log(det3 + 1e-6)).float() - math.log(1e3) * c) * float(N * T))
        gr['inv_w'].add_((-(float(N * T) / n) * (det3 / (det3 + 1e-6)) * torch.linalg.inv(W).t()).float())
    losses = Glow_Loss(z, sums[0], log_dets, sums[1], sigma)
    return z, losses, grads, d_mel


def Upsample_Mel_Backward(mels, d_up, params):
    """gradients of Upsample_Mel's kernel [1024,80,80] and bias [80] given d_up [N, keep, 80]"""
    lib = _lib.lib()
    N, Tm, _ = mels.shape
    keep = d_up.shape[1]
    x = mels.contiguous().float()
    d = d_up.contiguous().float()
    dk = torch.empty_like(params.up_kernel)
    db = torch.empty_like(params.up_bias)
    ws = torch.empty(lib.mstts_upsample_mel_bwd_workspace_bytes(N, Tm), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        rc = lib.mstts_upsample_mel_bwd(_lib.ptr(x), _lib.ptr(d), N, Tm, keep, _lib.ptr(dk), _lib.ptr(db), C.c_void_p(ws.data_ptr()),
                                        ws.numel(), _stream(x))
    _lib.check(rc, "mstts_upsample_mel_bwd")
    return dk, db
