"""Per-kernel device time of one full MSTTS train step (BASELINE config 5, B=16, one GPU), from CUPTI (torch.profiler).
Usage: python tools/profile_full_model.py [B]   -> table on stdout (kernel, launches, total us, share)"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_speaker_tts_b200 import MSTTS_SV, Feeder  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda:0")
    feeder = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(B, 128, 800), rank=0)
    model = MSTTS_SV.Tacotron2(is_Training=True, device=dev, feeder=feeder, process_group=None)
    pat = feeder.Get_Train_Pattern()
    for _ in range(3):
        model.Run_Train_Step(pat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model.Run_Train_Step(pat)
    e1.record()
    torch.cuda.synchronize()
    print("step %.3f ms (CUDA events, 5 steps)" % (e0.elapsed_time(e1) / 5))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.Run_Train_Step(pat)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split("(")[0][:70]
            agg[name][0] += 1
            agg[name][1] += ev.device_time
            t0 = ev.time_range.start
            t1 = ev.time_range.end
            t_first = t0 if t_first is None else min(t_first, t0)
            t_last = t1 if t_last is None else max(t_last, t1)
    # coarse timeline: kernels >= 60 us by start time, the short ones between them summed per stream
    evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    print("timeline (start ms, duration us, stream-ish id, kernel); '..' = short launches summed")
    small = [0, 0.0]
    for ev in evs:
        if ev.device_time >= 60:
            if small[0]:
                print("            .. %d short launches, %.0f us" % (small[0], small[1]))
                small = [0, 0.0]
            print("  %8.3f %9.1f  s%-3s %s" % ((ev.time_range.start - t_first) / 1e3, ev.device_time, getattr(ev, "device_resource_id", "?"),
                                            ev.name.split("(")[0][:60]))
        else:
            small[0] += 1
            small[1] += ev.device_time
    if small[0]:
        print("            .. %d short launches, %.0f us" % (small[0], small[1]))
    tot = sum(v for _, v in agg.values())
    print("device span %.3f ms, summed kernel time %.3f ms, %d launches" % ((t_last - t_first) / 1e3, tot / 1e3, sum(c for c, _ in agg.values())))
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-72s %5d %10.1f us %5.1f %%" % (name, c, v, 100 * v / tot))


if __name__ == "__main__":
    main()
