#!/usr/bin/env python
"""Full Tacotron2 train step through the reference surface (MSTTS_SV.Tacotron2.Run_Train_Step) at BASELINE config 2 shapes:
speaker-embedding net + encoder + fused decoder + postnet + losses + backward + TF Adam, host feed dict in, losses out.
python tools/bench_full_model.py [--steps K] [--B 32 --Te 128 --L 800]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--Te", type=int, default=128)
ap.add_argument("--L", type=int, default=800)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
from multi_speaker_tts_b200 import MSTTS_SV, Feeder
from multi_speaker_tts_b200.decoder import set_profiling, kernel_ms

# data parallel under torchrun (one process per GPU, one all-reduce of the flat gradient buffer per step)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda:%d" % local)
torch.cuda.set_device(dev)
pg = None
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
    pg = torch.distributed.group.WORLD
feeder = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(args.B, args.Te, args.L), rank=rank)
model = MSTTS_SV.Tacotron2(is_Training=True, device=dev, feeder=feeder, process_group=pg)
pat = feeder.Get_Train_Pattern()
for _ in range(3):
    model.Run_Train_Step(pat)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
set_profiling(True)
t0 = time.perf_counter()
for _ in range(args.steps):
    r = model.Run_Train_Step(pat)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / args.steps * 1e3
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = t.item()
    # every rank must hold identical parameters after identical all-reduced updates
    chk = model.flat_p.double().sum().reshape(1)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    torch.distributed.all_gather(gathered, chk)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
f_ms, f_n = kernel_ms(0)
b_ms, b_n = kernel_ms(1)
out = {"metric": "Tacotron2 full-model train step (reference surface)", "n_gpus": world, "ms_per_step": ms,
       "frames_per_s": world * args.B * args.L / (ms * 1e-3),
       "decoder_fwd_loop_ms": f_ms / max(f_n, 1), "decoder_bwd_loop_ms": b_ms / max(b_n, 1),
       "config": {"B": args.B, "Te": args.Te, "L": args.L}, "losses": {k: r[k] for k in ("Linear_Loss", "Postnet_Loss", "Stop_Loss")}}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    torch.distributed.destroy_process_group()
if args.profile and world == 1:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model.Run_Train_Step(pat)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
