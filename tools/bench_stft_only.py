"""STFT + mel (BASELINE config 4) alone: python tools/bench_stft_only.py  -> ms per call and the HBM-roofline fraction"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_secondary as BS  # noqa: E402

a = argparse.Namespace(steps=50, no_cpu=True)
pk, src = BS.peaks()
o = BS.bench_stft(a, pk, src, emit=False)
print("stft NT=%s: %.4f ms  frac %.4f" % (os.environ.get("MSTTS_STFT_NT", "512"), o["ms_per_step"], o["roofline"]["frac"]))
