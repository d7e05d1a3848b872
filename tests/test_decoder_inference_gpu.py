"""GPU parity of the free-running (inference) decode, Modules.py:212-237: the projected frame feeds the next prenet,
prenet dropout stays on, zoneout keeps its (1-r) factor without a mask, and the loop ends when every row has emitted
stop >= 0 or the step cap is reached.  Checked against the CPU oracle on identical inputs and dropout bits."""
import pytest
import torch

from multi_speaker_tts_b200 import synthetic as S

pytestmark = pytest.mark.gpu
TOL = 1e-3


MODES = ["fp32", "bf16x3"]


def _run(B, Te, cap, seed, dev, stop_bias=None, mode="fp32"):
    from oracle import decoder_oracle as O
    from multi_speaker_tts_b200.decoder import decoder_forward
    w = S.init_decoder_weights(0, bias_scale=0.05)
    if stop_bias is not None:
        w['projection/bias'][80] = stop_bias
    b = S.synthetic_decoder_batch(B, Te, cap, seed=seed, ragged=True)
    ref = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'], None,
                            is_training=False, max_steps=cap)
    wd = {k: v.to(dev) for k, v in w.items()}
    bd = {k: v.to(dev) for k, v in b.items()}
    lin, stop, align, _ = decoder_forward(wd, bd['memory'], bd['text_len'], None, None, bd['prenet_mask'], None,
                                          is_training=False, n_steps=cap + 1, mode=mode)
    return ref, (lin.cpu(), stop.cpu(), align.cpu())


def _check(ref, got):
    for r, g in zip(ref, got):
        assert g.shape == r.shape, "executed steps differ: %s vs oracle %s" % (tuple(g.shape), tuple(r.shape))
        assert torch.isfinite(g).all()
        assert (g - r).abs().max().item() < TOL
    assert torch.equal(got[1] >= 0, ref[1] >= 0), "stop decision differs"
    assert torch.equal(got[2].argmax(-1), ref[2].argmax(-1)), "alignment argmax differs"


@pytest.mark.parametrize("mode", MODES)
def test_stops_on_stop_token(cuda_dev, mode):
    ref, got = _run(2, 32, 30, 1, cuda_dev, mode=mode)
    assert ref[0].shape[1] < 31  # the oracle stopped before the cap
    _check(ref, got)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("B,Te,cap,seed", [(5, 33, 40, 4), (4, 64, 30, 5), (32, 128, 12, 6)])
def test_runs_to_step_cap(cuda_dev, B, Te, cap, seed, mode):
    ref, got = _run(B, Te, cap, seed, cuda_dev, mode=mode)
    assert ref[0].shape[1] == cap + 1
    _check(ref, got)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("B,Te,cap,seed", [(1, 7, 9, 8), (3, 9, 5, 2), (32, 100, 3, 11)])
def test_small_and_full_tiles(cuda_dev, B, Te, cap, seed, mode):
    """one row / short texts / a full 32-row tile; wherever the oracle stops, the kernel stops"""
    ref, got = _run(B, Te, cap, seed, cuda_dev, mode=mode)
    _check(ref, got)


@pytest.mark.parametrize("mode", MODES)
def test_delayed_stop(cuda_dev, mode):
    """rows finish at different steps (0 and 5) and the finished row keeps computing (impute_finished=False,
    Modules.py:116) until the last one emits stop >= 0"""
    ref, got = _run(2, 40, 60, 3, cuda_dev, stop_bias=-0.05, mode=mode)
    assert ref[0].shape[1] == 6
    _check(ref, got)


def test_large_batch_rows_wrap_clusters(cuda_dev):
    """B > 32: clusters process several batch rows per step"""
    ref, got = _run(36, 24, 6, 9, cuda_dev)
    _check(ref, got)


def test_bf16x3_free_running_refuses_shapes_beyond_its_tiling(cuda_dev):
    """B > 32 or Te > 128: explicit refusal (the caller picks mode fp32), never a silent fallback"""
    from multi_speaker_tts_b200._lib import MsttsError
    with pytest.raises(MsttsError):
        _run(36, 24, 6, 9, cuda_dev, mode="bf16x3")
    with pytest.raises(MsttsError):
        _run(2, 160, 6, 9, cuda_dev, mode="bf16x3")
