"""GPU: the persistent decoder loops spin on a grid barrier, so all 128 CTAs must be resident at once.  They are launched with
cudaLaunchAttributeCooperative (csrc/common.cuh: dec_cooperative_attr): the runtime gang-schedules the grid instead of letting a
concurrent kernel on another stream (an NCCL kernel, a user's side stream) hold SMs while half of the clusters spin.  Here a side
stream keeps every SM busy with large matmuls while a decoder train step (forward loop + reverse loop) runs: it must finish, and
its results must be bit-identical to the quiet run."""
import time

import pytest
import torch

from multi_speaker_tts_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _train_step(dev, mode):
    from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss
    w = {k: v.to(dev) for k, v in S.init_decoder_weights(0, bias_scale=0.05).items()}
    b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(8, 64, 40, seed=3, ragged=True).items()}
    T = int(b['mel_len'].max()) + 1
    lin, stop, align, st = decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'][:T].contiguous(),
                                           b['zone_mask'][:T].contiguous(), True, T, mode)
    loss2, dlin, dstop = decoder_loss(lin, stop, b['mel'], b['mel_len'])
    grads, dmem = decoder_backward(st, w, dlin, dstop)
    return lin, grads['cell_1/kernel'], dmem


@pytest.mark.timeout(180)
@pytest.mark.parametrize("mode", ["bf16x3", "fp32"])
def test_train_step_under_a_busy_side_stream(cuda_dev, mode):
    quiet = [t.clone() for t in _train_step(cuda_dev, mode)]
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=cuda_dev)
    a = torch.randn(8192, 8192, device=cuda_dev, dtype=torch.bfloat16)
    c = torch.empty_like(a)
    done = torch.cuda.Event()
    with torch.cuda.stream(side):
        for _ in range(60):           # ~1 ms each on every SM: the side stream stays busy for the whole step
            torch.matmul(a, a, out=c)
        done.record(side)
    t0 = time.perf_counter()
    busy = _train_step(cuda_dev, mode)
    torch.cuda.current_stream(cuda_dev).synchronize()
    dt = time.perf_counter() - t0
    overlapped = not done.query()     # the side stream was still running when the step finished (or just finished)
    torch.cuda.synchronize()
    print("%s train step under load: %.1f ms (side stream still busy at the end: %s)" % (mode, dt * 1e3, overlapped))
    for q, x in zip(quiet, busy):
        assert torch.equal(q, x), "results differ under a concurrent kernel"


def _all_grads(dev, T_mel=80):
    """one bf16x3 train step long enough (T >= 64) for the time-chunked overlap and the two post-loop chains to be active"""
    from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss
    w = {k: v.to(dev) for k, v in S.init_decoder_weights(0, bias_scale=0.05).items()}
    b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(6, 48, T_mel, seed=11, ragged=True).items()}
    T = int(b['mel_len'].max()) + 1
    assert T >= 64
    lin, stop, align, st = decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'][:T].contiguous(),
                                           b['zone_mask'][:T].contiguous(), True, T, "bf16x3")
    loss2, dlin, dstop = decoder_loss(lin, stop, b['mel'], b['mel_len'])
    grads, dmem = decoder_backward(st, w, dlin, dstop)
    out = {k: v.clone() for k, v in grads.items()}
    out['d_memory'] = dmem.clone()
    out['loss'] = torch.as_tensor(loss2, dtype=torch.float32).clone()  # per-block partials added in block order: no atomics
    return out


@pytest.mark.timeout(180)
def test_gradients_do_not_depend_on_the_stream_layout(cuda_dev, monkeypatch):
    """mstts_decoder_bwd runs its weight-gradient products on a side stream beside the loop and its post-loop work as two chains on
    two streams (csrc/decoder_bwd.cu).  None of that may change a bit: every gradient of the default layout equals the one-stream
    post-loop layout bit for bit, repeats are bit-identical, and the layout without any side stream agrees to accumulation-order
    accuracy (its time contraction is one product instead of eight accumulated chunks)."""
    base = _all_grads(cuda_dev)
    again = _all_grads(cuda_dev)
    for k in base:
        assert torch.equal(base[k], again[k]), "repeat differs: " + k
    monkeypatch.setenv("MSTTS_TAIL_STREAMS", "0")
    one = _all_grads(cuda_dev)
    for k in base:
        assert torch.equal(base[k], one[k]), "two-stream post-loop layout differs from the one-stream layout: " + k
    monkeypatch.delenv("MSTTS_TAIL_STREAMS")
    monkeypatch.setenv("MSTTS_NO_OVERLAP", "1")
    serial = _all_grads(cuda_dev)
    for k in base:
        ref = serial[k].double()
        err = (base[k].double() - ref).abs().max().item()
        assert err <= 2e-5 * max(ref.abs().max().item(), 1e-6), "%s: %.3e" % (k, err)
