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
    """
    _require_cuda(memory, text_len, mel, mel_len, prenet_mask, zone_mask)
    lib = _lib.lib()
    B, Te, D = memory.shape
    L = mel.shape[1]
    if n_steps is None:
        n_steps = int(mel_len.max().item()) + 1 if is_training else 1001
    T = n_steps
    dev = memory.device
    memory = memory.contiguous().float()
    mel = mel.contiguous().float()
    text_len = text_len.contiguous().to(torch.int32)
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
        memory=memory.data_ptr(), text_len=text_len.data_ptr(), mel=mel.data_ptr(), mel_len=mel_len.data_ptr(),
        prenet_mask=prenet_mask.data_ptr(), zone_mask=zone_mask.data_ptr() if zone_mask is not None else None,
        linear=linear.data_ptr(), stop=stop.data_ptr(), align=align.data_ptr(), steps_done=steps_done.data_ptr())
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = lib.mstts_decoder_fwd(C.byref(wstruct), C.byref(io), C.c_void_p(workspace.data_ptr()), workspace.numel(),
                                   C.c_void_p(s.cuda_stream))
    _lib.check(rc, "mstts_decoder_fwd")
    st = DecoderState()
    st.io, st.ws, st.wstruct = io, workspace, wstruct
    st.keep = keep + [memory, text_len, mel, mel_len, prenet_mask, zone_mask, steps_done]
    st.shape = (B, Te, L, D, T)
    return linear, stop, align, st
