"""Quick device timing of the decoder forward (not the bench): python tools/time_fwd.py B Te L [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_speaker_tts_b200 import synthetic as S
from multi_speaker_tts_b200.decoder import decoder_forward

B, Te, L = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
dev = torch.device("cuda:0")
w = {k: v.to(dev) for k, v in S.init_decoder_weights(0).items()}
b = {k: v.to(dev) for k, v in S.synthetic_decoder_batch(B, Te, L).items()}
ws = None
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lin, stop, al, st = decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'],
                                        b['zone_mask'], True, L + 1, mode, workspace=ws)
    ws = st.ws
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("iter %d: %.3f ms  -> %.1f frames/s  (%.2f us/step) finite=%s" % (it, ms, B * L / ms * 1e3, ms * 1e3 / (L + 1),
                                                                   bool(torch.isfinite(lin).all())))
