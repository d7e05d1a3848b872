"""B200-native hot path of CODEJIN/multi_speaker_tts behind the reference's Python surface.

Host code is Python on PyTorch tensors (device memory, streams, torch.distributed); all compute on the
hot path is hand-written sm_100a CUDA reached through the C ABI in ``include/mstts_b200.h``
(``libmstts_b200.so``, loaded by ``_lib``).  There is no CPU fallback: importing the compute entry
points without the built library raises.
"""
__version__ = "0.1.0"
