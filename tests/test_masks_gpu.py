"""GPU: the counter-based dropout / zoneout mask generator (mstts_fill_mask) and the decoder workspace guard.
The reference draws fresh tf.random_uniform noise every step (Modules.py:252, ZoneoutLSTMCell.py:266-271): masks of different
steps must be independent, not permutations of one another (callers pass consecutive small seeds)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_masks_of_nearby_seeds_are_not_permutations(cuda_dev):
    from multi_speaker_tts_b200.decoder import fill_mask
    n = 1 << 16
    masks = {}
    for seed in (1, 2, 5, 9, 4000013):
        m = torch.empty(n, device=cuda_dev, dtype=torch.uint8)
        fill_mask(m, 0.5, seed)
        masks[seed] = m.cpu()
    counts = {s: int(m.sum()) for s, m in masks.items()}
    for s, c in counts.items():
        assert abs(c - n / 2) < 6 * (n ** 0.5) / 2, (s, c)               # Bernoulli(0.5): 6 sigma
    assert len(set(counts.values())) >= 4, counts                        # XOR-permuted masks would all keep the same count
    groups = {s: sorted(m.view(-1, 4).view(torch.int32).view(-1).tolist()) for s, m in masks.items()}
    base = masks[1]
    for s in (5, 9):
        agree = float((masks[s] == base).float().mean())
        assert 0.47 < agree < 0.53, (s, agree)                            # independent bits agree half of the time
    # the same seed reproduces the same mask (checkpoint resume replays the stream)
    again = torch.empty(n, device=cuda_dev, dtype=torch.uint8)
    fill_mask(again, 0.5, 5)
    assert torch.equal(again.cpu(), masks[5])
    # keep probability is honoured
    z = torch.empty(1 << 18, device=cuda_dev, dtype=torch.uint8)
    fill_mask(z, 0.9, 123)
    assert abs(float(z.float().mean()) - 0.9) < 0.005
    assert groups[1] != groups[5]


def test_second_forward_before_backward_raises(cuda_dev):
    """one decoder workspace per device holds the saved activations: a stale backward must raise, not return wrong gradients"""
    from multi_speaker_tts_b200 import Modules, synthetic
    from multi_speaker_tts_b200.Location_Sensitive_Attention import Location_Sensitive_Attention
    B, Te, L, D = 2, 10, 6, 768
    w = {k: v.to(cuda_dev).requires_grad_(True) for k, v in synthetic.init_decoder_weights(seed=1, mem_dim=D).items()}
    g = torch.Generator().manual_seed(0)
    memory = torch.randn(B, Te, D, generator=g).to(cuda_dev)
    mel = torch.randn(B, L, 80, generator=g).to(cuda_dev)
    tl = torch.full((B,), Te, dtype=torch.int32, device=cuda_dev)
    ml = torch.full((B,), L, dtype=torch.int32, device=cuda_dev)

    def fwd():
        att = Location_Sensitive_Attention(128, memory, tl, 31, 1, 32, 0.0, True, variables=w)
        out, _ = Modules.Decoder_LSTM(mel, ml, att, True, variables=w, seed=3)
        return out.linear.sum()
    l1 = fwd()
    l2 = fwd()
    with pytest.raises(RuntimeError, match="overwritten by a later training forward"):
        l1.backward()
    l2.backward()        # the latest forward still differentiates
    assert w['cell_0/kernel'].grad is not None
