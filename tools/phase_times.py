"""Per-phase cycle breakdown of the bf16x3 persistent decoder kernels (CTA 0's clock64 stamps):
python tools/phase_times.py B Te L [fwd|bwd]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_speaker_tts_b200 import synthetic as S, _lib
from multi_speaker_tts_b200.decoder import decoder_forward, decoder_backward, decoder_loss

B, Te, L = (int(x) for x in sys.argv[1:4])
which = sys.argv[4] if len(sys.argv) > 4 else "fwd"
dev = torch.device("cuda:0")
w = {k: v.to(dev) for k, v in S.init_decoder_weights(0).items()}
b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(B, Te, L).items()}
T = L + 1
ws = None
for it in range(3):
    lin, stop, al, st = decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'],
                                        b['zone_mask'], True, T, "bf16x3", workspace=ws)
    ws = st.ws
    if which == "bwd":
        loss2, dlin, dstop = decoder_loss(lin, stop, b['mel'], b['mel_len'])
        decoder_backward(st, w, dlin, dstop)
torch.cuda.synchronize()
off = _lib.lib().mstts_decoder_ws_offset(b"dbg" if which == "fwd" else b"dbg_b", B, Te, L, 768, T, 1)
stamps = ws[off:off + T * 32 * 8].view(torch.int64).view(T, 32).cpu().double()
mhz = 1.0
names_b = ["start", "attn bwd done", "bar1", "B'e done", "bar2", "drain B'g done", "bar3", "A'e done", "bar4", "drain A'g done", "bar5"]
names_f = ["start", "J0 done", "reduceA done", "epiA done", "barA done", "J1 done", "reduceB done", "epiB done", "barB done",
           "attn done", "barC done", "e pushed", "e gathered", "softmax done", "q ready", "energies done"]
order_f = [0, 1, 2, 3, 4, 5, 6, 7, 8, 14, 15, 11, 12, 13, 9, 10]
sel = stamps[T // 4: 3 * T // 4]
if which == "bwd":
    # rows 0..127: global-timer stamps (ns) of every CTA at the middle step -> arrival spread at the phase ends
    allc = stamps[:128, :11]
    if (allc > 0).all():
        base = allc[:, 0].min()
        print("middle step, all 128 CTAs, global timer (ns since the first CTA started the step): min / median / max over CTAs")
        for k in range(11):
            col = allc[:, k] - base
            print("  %-16s %8.0f %8.0f %8.0f   slowest CTA %d" % (names_b[k], col.min().item(), col.median().item(), col.max().item(),
                                                                   int(col.argmax())))
    prev = sel[:, 0]
    print("reverse kernel, cycles per step (mean over the middle half of the steps), CTA 0:")
    for k in range(11):
        print("  %-16s +%8.0f" % (names_b[k], (sel[:, k] - prev).mean().item()))
        prev = sel[:, k]
    print("  total/step     %8.0f cycles" % (sel[:, 10] - sel[:, 0]).mean().item())
    sys.exit(0)
allc = stamps[:128, :16]
if T // 2 >= 128 and (allc[:, :11] > 0).all():
    base = allc[:, 0].min()
    print("middle step, all 128 CTAs, global timer (ns since the first CTA started the step): min / median / max over CTAs")
    for k in order_f:
        col = allc[:, k] - base
        if (col < 0).any():   # attention stamps exist only in the CTAs that own a batch row
            continue
        print("  %-14s %8.0f %8.0f %8.0f   slowest CTA %d" % (names_f[k], col.min().item(), col.median().item(), col.max().item(),
                                                               int(col.argmax())))
prev = sel[:, 0]
print("cycles per step (mean over the middle half of the steps), CTA 0:")
for k in order_f:
    d = (sel[:, k] - prev).mean().item()
    print("  %-14s +%8.0f" % (names_f[k], d))
    prev = sel[:, k]
print("  total/step   %8.0f cycles" % (sel[:, 10] - sel[:, 0]).mean().item())
if which == "fwd":
    print("MMA thread: J0 committed at +%.0f after step start, J1 committed at +%.0f after barA done" % (
        (sel[:, 20] - sel[:, 0]).mean().item(), (sel[:, 25] - sel[:, 4]).mean().item()))
