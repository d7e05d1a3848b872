import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
which = sys.argv[1]
dev = torch.device("cuda:0")
if which == "stft":
    from multi_speaker_tts_b200 import Audio as G
    wav = (torch.rand(64, 220500, device=dev) * 1.98 - 0.99)
    for _ in range(3):
        G.stft_features(wav, 1024, 256, 1024, 22050, 80, 4.0)
else:
    from oracle import waveglow_oracle as W
    from multi_speaker_tts_b200.WaveGlow import Modules as M
    raws, upk, upb = W.init_waveglow(0, end_scale=0.01, g_mode="unit", inv_mode="orthogonal")
    params = M.WaveGlowParams(raws, upk, upb, dev)
    audio, mel = W.synthetic_batch(8, 16000, 64)
    ad, md = audio.to(dev), mel.to(dev)
    for _ in range(2):
        a, m = M.Restructure_Train_Data(ad, md, params)
        z, ls, ld, ss = M.Glow_Train(a, m, params)
torch.cuda.synchronize()
