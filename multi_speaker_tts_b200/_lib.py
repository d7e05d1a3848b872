"""ctypes binding of libmstts_b200.so (the C ABI in include/mstts_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmstts_b200.so")

MODE_FP32, MODE_BF16X3, MODE_BF16 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "bf16x3": MODE_BF16X3, "bf16": MODE_BF16}

_fp = C.c_void_p

DECODER_WEIGHT_FIELDS = [
    # (struct field, key in the python weight dict)
    ("prenet0_kernel", "prenet_0/kernel"), ("prenet0_bias", "prenet_0/bias"),
    ("prenet1_kernel", "prenet_1/kernel"), ("prenet1_bias", "prenet_1/bias"),
    ("cell0_kernel", "cell_0/kernel"), ("cell0_bias", "cell_0/bias"),
    ("cell1_kernel", "cell_1/kernel"), ("cell1_bias", "cell_1/bias"),
    ("memory_kernel", "memory_layer/kernel"), ("query_kernel", "query_layer/kernel"),
    ("loc_conv_kernel", "location/conv1d/kernel"), ("loc_conv_bias", "location/conv1d/bias"),
    ("loc_dense_kernel", "location/dense/kernel"),
    ("score_w", "score/weight_w"), ("score_b", "score/bias_b"),
    ("proj_kernel", "projection/kernel"), ("proj_bias", "projection/bias"),
]


class MsttsDecoderWeights(C.Structure):
    _fields_ = [(f, _fp) for f, _ in DECODER_WEIGHT_FIELDS]


class MsttsDecoderWeightGrads(C.Structure):
    _fields_ = [(f, _fp) for f, _ in DECODER_WEIGHT_FIELDS]


class MsttsDecoderIO(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("Te", C.c_int), ("L", C.c_int), ("D", C.c_int), ("n_steps", C.c_int),
        ("is_training", C.c_int), ("mode", C.c_int),
        ("memory", _fp), ("text_len", _fp), ("mel", _fp), ("mel_len", _fp),
        ("prenet_mask", _fp), ("zone_mask", _fp),
        ("linear", _fp), ("stop", _fp), ("align", _fp), ("steps_done", _fp),
    ]


class MsttsDecoderGrads(C.Structure):
    _fields_ = [("d_linear", _fp), ("d_stop", _fp), ("d_memory", _fp)]


class MsttsWaveGlowWeights(C.Structure):
    _fields_ = ([("inv_w", _fp * 12), ("start_g", _fp * 12), ("start_v", _fp * 12), ("start_b", _fp * 12)] +
                [(n + "_" + k, (_fp * 8) * 12) for n in ("in", "cond", "res") for k in ("g", "v", "b")] +
                [("end_w", _fp * 12), ("end_b", _fp * 12)])


class MsttsWaveGlowGrads(C.Structure):
    _fields_ = MsttsWaveGlowWeights._fields_


EXPORTS = {
    # name: (restype, argtypes)
    "mstts_version": (C.c_int, []),
    "mstts_last_error": (C.c_char_p, []),
    "mstts_device_check": (C.c_int, [C.c_int]),
    "mstts_set_profiling": (C.c_int, [C.c_int]),
    "mstts_kernel_ms": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "mstts_decoder_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "mstts_decoder_ws_offset": (C.c_size_t, [C.c_char_p] + [C.c_int] * 6),
    "mstts_decoder_fwd": (C.c_int, [C.POINTER(MsttsDecoderWeights), C.POINTER(MsttsDecoderIO), _fp, C.c_size_t, _fp]),
    "mstts_decoder_bwd": (C.c_int, [C.POINTER(MsttsDecoderWeights), C.POINTER(MsttsDecoderIO),
                                    C.POINTER(MsttsDecoderGrads), C.POINTER(MsttsDecoderWeightGrads),
                                    _fp, C.c_size_t, _fp]),
    "mstts_decoder_loss": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "mstts_waveglow_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mstts_waveglow_flows": (C.c_int, [C.POINTER(MsttsWaveGlowWeights), _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp,
                                       C.c_size_t, _fp]),
    "mstts_waveglow_set_path": (C.c_int, [C.c_int]),
    "mstts_waveglow_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mstts_waveglow_train_fwd": (C.c_int, [C.POINTER(MsttsWaveGlowWeights), _fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "mstts_waveglow_train_bwd": (C.c_int, [C.POINTER(MsttsWaveGlowWeights), C.POINTER(MsttsWaveGlowGrads), _fp, C.c_int, C.c_int,
                                           C.c_float, _fp, _fp, C.c_size_t, _fp]),
    "mstts_upsample_mel_bwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mstts_upsample_mel_bwd": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "mstts_upsample_mel_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mstts_upsample_mel": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "mstts_stft_mel_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "mstts_stft_mel": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _fp, _fp,
                                 _fp, C.c_size_t, _fp]),
    "mstts_zlstm_fwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp, _fp, _fp, _fp]),
    "mstts_zlstm_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp]),
    "mstts_tc_gemm_tiled": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, _fp]),
    "mstts_tc_gemm_test": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "mstts_gemm_f32": (C.c_int, [C.c_int] * 5 + [_fp, C.c_int, C.c_longlong, _fp, C.c_int, C.c_longlong, _fp, C.c_int, C.c_longlong,
                                 C.c_float, C.c_int, C.c_int, _fp]),
    "mstts_release_scratch": (C.c_int, []),
    "mstts_conv1d_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "mstts_conv1d_fwd": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "mstts_conv1d_bwd": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "mstts_act_bn_dropout_workspace_bytes": (C.c_size_t, [C.c_int]),
    "mstts_act_bn_dropout_fwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                           C.c_float, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "mstts_act_bn_dropout_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_longlong, C.c_int, C.c_int, C.c_float, _fp, _fp, _fp, _fp,
                                           C.c_size_t, _fp]),
    "mstts_fill_mask": (C.c_int, [_fp, C.c_size_t, C.c_float, C.c_uint64, _fp]),
    "mstts_adam_tf": (C.c_int, [_fp, _fp, _fp, _fp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, _fp]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmstts_b200.so not found at %s -- build it with `python __graft_entry__.py build` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)  # raises AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class MsttsError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().mstts_last_error()
        raise MsttsError("%s failed (%d): %s" % (what or "libmstts_b200 call", rc, msg.decode() if msg else ""))


def stream_ptr(device):
    """raw handle of the current CUDA stream on ``device``.  Every library call needs it; building a torch Stream object for it
    (torch.cuda.current_stream) cost more host time than the launch itself."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


class on_device(object):
    """``with on_device(dev):`` -- torch.cuda.device(dev) that costs nothing when ``dev`` is already the current device"""
    __slots__ = ('ctx',)

    def __init__(self, device):
        import torch
        idx = device.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def ptr(t):
    """device pointer of a contiguous tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_contiguous(), "libmstts_b200 needs contiguous tensors"
    return C.c_void_p(t.data_ptr())
