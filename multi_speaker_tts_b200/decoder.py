"""Host side of the decoder op: torch tensors in, C ABI call, torch tensors out.

``decoder_forward`` / ``DecoderFunction`` are what ``Modules.Decoder_LSTM`` dispatches to.
"""
import ctypes as C

import torch

from . import _lib

MEL = 80


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("multi_speaker_tts_b200 decoder runs on CUDA tensors only (no CPU fallback)")


def pack_weights(weights, struct_cls=_lib.MsttsDecoderWeights):
    s = struct_cls()
    keep = []
    for field, key in _lib.DECODER_WEIGHT_FIELDS:
        t = weights[key]
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.contiguous().float()
        keep.append(t)
        setattr(s, field, t.data_ptr())
    return s, keep


class DecoderState(object):
    """Everything one fwd call produced that bwd needs (workspace holds the saved activations)."""
    __slots__ = ("io", "ws", "wstruct", "keep", "shape")


def decoder_forward(weights, memory, text_len, mel, mel_len, prenet_mask, zone_mask, is_training=True,
                    n_steps=None, mode="fp32", stream=None, workspace=None):
    """Run the decoder loop on the GPU.

    weights: dict (keys as synthetic.decoder_weight_shapes) of fp32 CUDA tensors in TF layouts.
    Returns (linear [B,T,80], stop [B,T], align [B,T,Te], state).

    ``is_training=False`` is the free-running decode of Modules.py:212-237 (``mel`` / ``mel_len`` / ``zone_mask`` are
    ignored and may be None): ``n_steps`` is the step cap + 1 (default Max_Inference_Length + 1 = 1001), the loop
    stops once every row has emitted ``stop >= 0`` and the outputs are cut to the executed steps (one host sync).
    ``mode="bf16x3"`` runs it on the tcgen05 loop (projection + prenet inside the kernel; B <= 32, Te <= 128), ``"fp32"`` on
    the SIMT kernel.
    """
    _require_cuda(memory, text_len, prenet_mask)
    lib = _lib.lib()
    B, Te, D = memory.shape
    dev = memory.device
    if not is_training:
        mel, mel_len, zone_mask = None, None, None
        if n_steps is None:
            n_steps = 1001
    else:
        _require_cuda(mel, mel_len, zone_mask)
    L = mel.shape[1] if mel is not None else 0
    if n_steps is None:
        n_steps = int(mel_len.max().item()) + 1
    T = n_steps
    memory = memory.contiguous().float()
    text_len = text_len.contiguous().to(torch.int32)
    if mel is not None:
        mel = mel.contiguous().float()
        mel_len = mel_len.contiguous().to(torch.int32)
    prenet_mask = prenet_mask.contiguous()
    assert prenet_mask.dtype == torch.uint8 and tuple(prenet_mask.shape) == (T, 2, B, 256), prenet_mask.shape
    if zone_mask is not None:
        zone_mask = zone_mask.contiguous()
        assert zone_mask.dtype == torch.uint8 and tuple(zone_mask.shape) == (T, 2, 2, B, 1024), zone_mask.shape
    linear = torch.empty(B, T, MEL, device=dev, dtype=torch.float32)
    stop = torch.empty(B, T, device=dev, dtype=torch.float32)
    align = torch.empty(B, T, Te, device=dev, dtype=torch.float32)
    steps_done = torch.zeros(1, device=dev, dtype=torch.int32)
    m = _lib.MODES[mode] if isinstance(mode, str) else int(mode)
    nbytes = lib.mstts_decoder_workspace_bytes(B, Te, L, D, T, m)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    wstruct, keep = pack_weights(weights)
    io = _lib.MsttsDecoderIO(
        B=B, Te=Te, L=L, D=D, n_steps=T, is_training=int(bool(is_training)), mode=m,
        memory=memory.data_ptr(), text_len=text_len.data_ptr(), mel=mel.data_ptr() if mel is not None else None,
        mel_len=mel_len.data_ptr() if mel_len is not None else None,
        prenet_mask=prenet_mask.data_ptr(), zone_mask=zone_mask.data_ptr() if zone_mask is not None else None,
        linear=linear.data_ptr(), stop=stop.data_ptr(), align=align.data_ptr(), steps_done=steps_done.data_ptr())
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(dev)
    with _lib.on_device(dev):
        rc = lib.mstts_decoder_fwd(C.byref(wstruct), C.byref(io), C.c_void_p(workspace.data_ptr()), workspace.numel(),
                                   s)
    _lib.check(rc, "mstts_decoder_fwd")
    st = DecoderState()
    st.io, st.ws, st.wstruct = io, workspace, wstruct
    st.keep = keep + [memory, text_len, mel, mel_len, prenet_mask, zone_mask, steps_done]
    st.shape = (B, Te, L, D, T)
    if not is_training:
        n = int(steps_done.item())
        linear, stop, align = linear[:, :n].contiguous(), stop[:, :n].contiguous(), align[:, :n].contiguous()
    return linear, stop, align, st


def decoder_backward(state, weights, d_linear, d_stop, stream=None, want_d_memory=True, grad_out=None):
    """Reverse pass for a ``decoder_forward`` call.  Returns (grads dict keyed like ``weights``, d_memory).
    ``grad_out``: optional dict of preallocated fp32 tensors (e.g. views of one flat buffer) to write into."""
    lib = _lib.lib()
    B, Te, L, D, T = state.shape
    dev = d_linear.device
    d_linear = d_linear.contiguous().float()
    d_stop = d_stop.contiguous().float()
    assert tuple(d_linear.shape) == (B, T, MEL) and tuple(d_stop.shape) == (B, T)
    grads = grad_out if grad_out is not None else {
        key: torch.empty_like(weights[key], dtype=torch.float32).contiguous() for _, key in _lib.DECODER_WEIGHT_FIELDS}
    gstruct = _lib.MsttsDecoderWeightGrads()
    for field, key in _lib.DECODER_WEIGHT_FIELDS:
        setattr(gstruct, field, grads[key].data_ptr())
    d_memory = torch.empty(B, Te, D, device=dev, dtype=torch.float32) if want_d_memory else None
    g = _lib.MsttsDecoderGrads(d_linear=d_linear.data_ptr(), d_stop=d_stop.data_ptr(),
                               d_memory=d_memory.data_ptr() if d_memory is not None else None)
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(dev)
    with _lib.on_device(dev):
        rc = lib.mstts_decoder_bwd(C.byref(state.wstruct), C.byref(state.io), C.byref(g), C.byref(gstruct),
                                   C.c_void_p(state.ws.data_ptr()), state.ws.numel(), s)
    _lib.check(rc, "mstts_decoder_bwd")
    return grads, d_memory


def decoder_loss(linear, stop, mel, mel_len, use_l1=True, stream=None):
    """Decoder terms of MSTTS_SV.py:127-144 and their gradients, on device.
    Returns (loss2 [linear_loss, stop_loss], d_linear, d_stop)."""
    lib = _lib.lib()
    B, T, _ = linear.shape
    L = mel.shape[1]
    dev = linear.device
    loss2 = torch.empty(2, device=dev, dtype=torch.float32)
    d_linear = torch.empty_like(linear)
    d_stop = torch.empty_like(stop)
    mel = mel.contiguous().float()
    mel_len = mel_len.contiguous().to(torch.int32)
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(dev)
    with _lib.on_device(dev):
        rc = lib.mstts_decoder_loss(_lib.ptr(linear), _lib.ptr(stop), _lib.ptr(mel), _lib.ptr(mel_len), B, L, T,
                                    int(bool(use_l1)), _lib.ptr(loss2), _lib.ptr(d_linear), _lib.ptr(d_stop),
                                    s)
    _lib.check(rc, "mstts_decoder_loss")
    return loss2, d_linear, d_stop


def fill_mask(out, keep_prob, seed, stream=None):
    """out (uint8 CUDA tensor) <- Bernoulli(keep_prob) bits from the counter-based generator."""
    lib = _lib.lib()
    assert out.dtype == torch.uint8 and out.is_cuda and out.is_contiguous()
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(out.device)
    with _lib.on_device(out.device):
        rc = lib.mstts_fill_mask(_lib.ptr(out), out.numel(), float(keep_prob), int(seed) & (2 ** 64 - 1),
                                 s)
    _lib.check(rc, "mstts_fill_mask")
    return out


def adam_tf(p, m, v, g, lr_t, b1=0.9, b2=0.999, eps=1e-6, grad_scale=1.0, l2=0.0, stream=None):
    """In-place tf.train.AdamOptimizer update on flat fp32 CUDA buffers."""
    lib = _lib.lib()
    for t in (p, m, v, g):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    s = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(p.device)
    with _lib.on_device(p.device):
        rc = lib.mstts_adam_tf(_lib.ptr(p), _lib.ptr(m), _lib.ptr(v), _lib.ptr(g), p.numel(), float(lr_t), float(b1),
                               float(b2), float(eps), float(grad_scale), float(l2), s)
    _lib.check(rc, "mstts_adam_tf")


def set_profiling(on):
    _lib.check(_lib.lib().mstts_set_profiling(int(bool(on))), "mstts_set_profiling")


def kernel_ms(which):
    """(sum of device ms, launch count) of the persistent kernel `which` (0 fwd loop, 1 reverse loop)."""
    ms, n = C.c_float(0), C.c_int(0)
    _lib.check(_lib.lib().mstts_kernel_ms(int(which), C.byref(ms), C.byref(n)), "mstts_kernel_ms")
    return ms.value, n.value
