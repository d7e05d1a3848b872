#!/usr/bin/env python
"""Headline benchmark: decoder mel-frames/s of one Tacotron2-decoder TRAIN step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode fp32]

A step = masks -> decoder forward loop -> loss -> reverse loop -> weight-gradient GEMMs -> (one NCCL
all-reduce of the flat gradient buffer when N > 1) -> TF-style Adam, on synthetic inputs of BASELINE config 2
(B=32 per GPU, text_len=128, mel_len=800, 80-mel).  Rank 0 prints ONE JSON line.  `value` has the inputs
resident in HBM; `e2e` adds the host->device copy of each step's inputs (pinned) and a device->host read of
the loss.  `--impl reference` times the CPU restatement of the reference (oracle/, the reference itself is
TF1 and cannot run here) on a bounded sample of the same workload.  The same line carries `full_model` (BASELINE config 5:
MSTTS_SV.Tacotron2.Run_Train_Step at B=16/GPU with the 121 MB all-reduce, at every N; `--workload full` makes it the
headline) and, at N=1, `secondary` (configs 3 and 4: WaveGlow forward+NLL / train step, STFT+mel, each with its own
roofline and cpu_baseline).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "decoder mel-frames/s (train step)"
UNIT = "frames/s"
B_PER_GPU, TE, L, D = 32, 128, 800, 768
W_STEP = 20299345  # parameters touched per decoder step
MY_KERNELS_PER_STEP = 110  # fallback when the CUPTI profiler is unavailable (count of profiles/r2_launches_bf16x3.csv per step); the
# line normally carries the count measured live on one extra untimed step (count_launches)


def workload_config(world):
    """identical in both arms (`--impl b200` and `--impl reference`): the driver compares the two dicts"""
    return {
        "workload": "Tacotron2 decoder train step (BASELINE config 2): B=%d/GPU text_len=%d mel_len=%d 80-mel, %d decoder steps, "
                    "fwd+loss+bwd+TF-Adam (+1 gradient all-reduce when n_gpus > 1)" % (B_PER_GPU, TE, L, L + 1),
        "batch_per_gpu": B_PER_GPU, "text_len": TE, "mel_len": L, "decoder_steps": L + 1, "parallelism": "dp%d" % world,
        "l2_policy": "per-step working set (saved activations + weights, ~2.4 GB) exceeds the 126 MB L2; no flush",
    }


LIBRARY_GEMM_RE = r"nvjet|cutlass|cublas|xmma|gemv|gemm_e|cudnn|sm[0-9]+_.*(gemm|mma)"


def count_launches(fn):
    """Kernel launches of one call of `fn`, measured with the CUPTI-backed torch profiler (outside any timed region).
    Returns {"own": kernels of libmstts_b200.so, "aten": torch elementwise / reduction kernels, "library_gemm": vendor GEMM
    kernels (must be 0), "nccl": ..., "names": {kernel: count}} or None when the profiler is unavailable."""
    import re
    try:
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = {}
        for ev in prof.events():
            if str(getattr(ev, "device_type", "")).endswith("CUDA") and ev.name and not ev.name.startswith(("Memcpy", "Memset")):
                names[ev.name] = names.get(ev.name, 0) + 1
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}
    out = {"own": 0, "aten": 0, "library_gemm": 0, "nccl": 0}
    for n, c in names.items():
        if re.search(r"nccl", n, re.I):
            out["nccl"] += c
        elif re.search(LIBRARY_GEMM_RE, n, re.I) and "tc_gemm_kernel" not in n:
            out["library_gemm"] += c
        elif n.startswith(("void at::", "at::", "void at_cuda", "void (anonymous namespace)::elementwise")) or "at::native" in n:
            out["aten"] += c
        else:
            out["own"] += c
    out["names"] = {k: v for k, v in sorted(names.items(), key=lambda kv: -kv[1])[:40]}
    return out


def fwd_bytes_per_step(B, Te, s_w=4, s_kv=4):
    """algorithmic bytes of one forward decoder step (step-streaming model)."""
    return W_STEP * s_w + B * Te * 896 * s_kv + 12 * B * Te + 32768 * B + 644 * B


def saved_bytes_per_step(B, Te):
    """activations written once by the forward loop and read once by the reverse loop"""
    return B * (2 * 4 * 1024 + 2 * 2 * 1024 + Te + 1792) * 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML every 200 ms while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.ok = index, False, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_train_step_oracle(B, Te, Ls, threads, n_timed=1, n_warm=0):
    """Reference CPU path = the oracle restatement, op for op per step, torch.autograd backward, TF Adam.
    Returns seconds per step on a (B, Te, Ls) sample."""
    from oracle import decoder_oracle as O  # the one place bench.py executes the oracle
    from multi_speaker_tts_b200 import synthetic as S
    torch.set_num_threads(threads)
    w = {k: v.requires_grad_(True) for k, v in S.init_decoder_weights(0).items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v2 = {k: torch.zeros_like(v) for k, v in w.items()}
    b = S.synthetic_decoder_batch(B, Te, Ls)
    times = []
    for it in range(n_warm + n_timed):
        t0 = time.perf_counter()
        lin, stop, al = O.decoder_forward(w, b['memory'], b['text_len'], b['mel'], b['mel_len'], b['prenet_mask'],
                                          b['zone_mask'])
        ll, sl = O.decoder_loss(lin, stop, b['mel'], b['mel_len'])
        (ll + sl).backward()
        with torch.no_grad():
            for k in w:
                O.tf_adam_step(w[k], m[k], v2[k], w[k].grad, it + 1, O.learning_rate(it))
                w[k].grad = None
        dt = time.perf_counter() - t0
        if it >= n_warm:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # Bounded sample, chosen so that the reference is not charged its per-step fixed costs (Adam over 20.3 M weights, the
    # batched prenet / projection products) on a short sequence: two probes fit sec(L) = c0 + c1 * (L + 1), then the largest
    # mel_len <= 800 whose (warmup + steps) train steps fit the time budget is timed -- the full workload on the GPU box's
    # 16 cores at the default K/W.
    budget_s = float(os.environ.get("MSTTS_REF_BUDGET_S", "200"))
    s_a = cpu_train_step_oracle(B_PER_GPU, TE, 16, threads)
    s_b = cpu_train_step_oracle(B_PER_GPU, TE, 48, threads)
    c1 = max((s_b - s_a) / 32.0, 1e-6)
    c0 = max(s_a - 17 * c1, 0.0)
    n_runs = max(args.steps + args.warmup, 1)
    Ls = int(min(L, max(24, (budget_s / n_runs - c0) / c1 - 1)))
    sec = cpu_train_step_oracle(B_PER_GPU, TE, Ls, threads, n_timed=args.steps, n_warm=args.warmup)
    val = B_PER_GPU * Ls / sec
    sample = "B=%d Te=%d L=%d (%d of 801 decoder steps per train step), fwd+loss+autograd bwd+TF Adam, %.1f s/step" % (
        B_PER_GPU, TE, Ls, Ls + 1, sec)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is TF1 (cannot run in this image); this is the PyTorch-CPU restatement in oracle/",
    }))


def time_region(fn, steps, barrier):
    """CUDA-event time of `steps` calls of fn on the current stream, barrier + synchronize on both sides (ms, wall ms)"""
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(steps):
        last = fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, last


def gather_floats(x, world, dev):
    if world == 1:
        return [float(x)]
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(out, t)
    return [float(o.item()) for o in out]


def run_full_model(args, dev, pg, world, rank, barrier, steps):
    """BASELINE config 5: full MSTTS (frozen speaker-embedding net + Tacotron2 encoder / decoder / postnet, losses, backward,
    ONE all-reduce of the 30.3 M-float gradient buffer, TF Adam) through MSTTS_SV.Tacotron2.Run_Train_Step, B=16 per GPU.
    Host feed dict in (pinned copies inside), losses read back: end to end by construction."""
    from multi_speaker_tts_b200 import MSTTS_SV, Feeder
    Bf = 16
    feeder = Feeder.Feeder(is_Training=True, synthetic=True, synthetic_shape=(Bf, TE, L), rank=rank)
    model = MSTTS_SV.Tacotron2(is_Training=True, device=dev, feeder=feeder, process_group=pg)
    pat = feeder.Get_Train_Pattern()
    for _ in range(3):
        model.Run_Train_Step(pat)
    model.allreduce_events = []
    ms, wall, r = time_region(lambda: model.Run_Train_Step(pat), steps, barrier)
    ar = [a.elapsed_time(b) for a, b in model.allreduce_events]
    model.allreduce_events = None
    own = max(ms, wall)
    per_rank = gather_floats(own / steps, world, dev)
    ar_rank = gather_floats(sum(ar) / max(len(ar), 1), world, dev)
    if world > 1:
        t = torch.tensor([own], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        own = t.item()
        chk = model.flat_p.double().sum().reshape(1)
        g = [torch.zeros_like(chk) for _ in range(world)]
        torch.distributed.all_gather(g, chk)
        same = all(torch.equal(x, g[0]) for x in g)
    else:
        same = True
    out = {"metric": "decoder mel-frames/s (full MSTTS train step)", "value": world * Bf * L * steps / (own * 1e-3), "unit": UNIT,
           "n_gpus": world, "steps": steps, "ms_per_step": own / steps, "scaling": "weak",
           "config": {"workload": "Full MSTTS data-parallel train step (BASELINE config 5): frozen speaker-embedding net + Tacotron2 "
                                  "(encoder, decoder, postnet) fwd+loss+bwd+TF-Adam through MSTTS_SV.Tacotron2.Run_Train_Step, "
                                  "B=%d/GPU text_len=%d mel_len=%d, host feed dict in, losses out" % (Bf, TE, L),
                      "batch_per_gpu": Bf, "parallelism": "dp%d" % world},
           "allreduce_bytes": int(model.flat_g.numel()) * 4 if world > 1 else 0,
           "allreduce_ms_per_rank": ar_rank, "ms_per_step_per_rank": per_rank, "replicas_identical": bool(same),
           "losses": {k: r[k] for k in ("Linear_Loss", "Postnet_Loss", "Stop_Loss")}}
    del model
    torch.cuda.empty_cache()
    return out


def run_secondary(args, dev, peaks, peak_kind):
    """BASELINE configs 3 and 4 on one B200, each with the roofline that bounds it and the oracle timed on the host cores"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_secondary as BS
    out = {}
    for name, fn in (("waveglow", BS.bench_waveglow), ("stft_mel", BS.bench_stft)):
        try:
            out[name] = fn(args, peaks, peak_kind, emit=False)
        except Exception as e:  # the headline line must survive a secondary failure
            out[name] = {"error": repr(e)}
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--workload", default="decoder", choices=["decoder", "full"],
                    help="decoder = BASELINE config 2 (the headline metric); full = config 5 as the headline line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-secondary", action="store_true", help="skip configs 3 / 4 (WaveGlow, STFT) and the config-5 section")
    args = ap.parse_args()
    args.no_cpu = args.no_cpu or args.no_cpu_baseline
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: library chatter (NCCL prints its version banner to stdout) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD

    from multi_speaker_tts_b200 import synthetic as S
    from multi_speaker_tts_b200.trainer import DecoderTrainer
    from multi_speaker_tts_b200.decoder import set_profiling, kernel_ms

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def emit(out):
        if rank == 0:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(out), flush=True)
            os.dup2(2, 1)
        if world > 1:
            torch.distributed.destroy_process_group()

    peaks, peak_kind = measured_peaks()
    if args.workload == "full":
        out = run_full_model(args, dev, pg, world, rank, barrier, args.steps)
        out.update({"warmup": 3, "higher_is_better": True, "vs_baseline": None, "dtype": "bf16x3", "data": "synthetic",
                    "e2e": {"value": out["value"], "unit": UNIT, "note": "Run_Train_Step takes host feed dicts and returns host losses"}})
        emit(out)
        return

    B, T = B_PER_GPU, L + 1
    tr = DecoderTrainer(dev, mem_dim=D, mode=args.mode, seed=0, process_group=pg)
    host = S.synthetic_decoder_batch(B, TE, L, seed=1234, rank=rank)
    pinned = {k: host[k].pin_memory() for k in ('memory', 'text_len', 'mel', 'mel_len')}
    dev_in = {k: v.to(dev) for k, v in pinned.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())

    def step_resident():
        return tr.train_step(dev_in['memory'], dev_in['text_len'], dev_in['mel'], dev_in['mel_len'], T)

    def step_e2e():
        cur = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        loss2 = tr.train_step(cur['memory'], cur['text_len'], cur['mel'], cur['mel_len'], T)
        return loss2.cpu()  # device->host read of the step's result (synchronises)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    # one extra untimed step on EVERY rank (it contains the all-reduce); rank 0 runs it under the CUPTI profiler
    launches = None
    if rank == 0:
        launches = count_launches(step_resident)
    else:
        step_resident()
    step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    set_profiling(True)
    tr.allreduce_events = []
    ms, _, _ = time_region(step_resident, args.steps, barrier)
    fwd_ms, nf = kernel_ms(0)
    bwd_ms, nb = kernel_ms(1)
    set_profiling(False)
    ar = [a.elapsed_time(b) for a, b in tr.allreduce_events]
    tr.allreduce_events = None
    # the same step with the weight-gradient products AFTER the reverse loop instead of beside it (MSTTS_NO_OVERLAP=1): shows what
    # the overlap buys and what the contention costs the loop kernel (short run, not part of `value`)
    os.environ["MSTTS_NO_OVERLAP"] = "1"
    step_resident()
    set_profiling(True)
    ms_no, _, _ = time_region(step_resident, min(args.steps, 5), barrier)
    f2, n2 = kernel_ms(0)
    b2, m2 = kernel_ms(1)
    set_profiling(False)
    del os.environ["MSTTS_NO_OVERLAP"]
    no_overlap = {"step_ms": ms_no / min(args.steps, 5), "fwd_loop_ms": f2 / max(n2, 1), "bwd_loop_ms": b2 / max(m2, 1)}
    # end-to-end arm (host buffers, copies inside the timed region)
    step_e2e()
    ms_e2e, wall_e2e, last = time_region(step_e2e, args.steps, barrier)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    per_rank_ms = gather_floats(ms / args.steps, world, dev)
    ar_rank = gather_floats(sum(ar) / max(len(ar), 1), world, dev)
    if world > 1:
        tmax = torch.tensor([ms, ms_e2e, wall_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e, wall_e2e = tmax.tolist()
    frames = world * B * L * args.steps
    value = frames / (ms * 1e-3)
    e2e_val = frames / (max(ms_e2e, wall_e2e) * 1e-3)

    hbm = float(peaks["hbm_gbs"])
    s_w = 4  # fp32 weights, or bf16 hi + lo = the same 4 bytes per stored weight
    fwd_alg = fwd_bytes_per_step(B, TE, s_w, 4) * T
    bwd_alg = (fwd_bytes_per_step(B, TE, s_w, 4) + saved_bytes_per_step(B, TE)) * T
    fwd_avg, bwd_avg = fwd_ms / max(nf, 1), bwd_ms / max(nb, 1)

    # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture of this command at the
    # default workload (profiles/r2_traffic.json: ncu cannot run inside a timed bench); null for any other shape / mode
    traffic, traffic_src = {}, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath) and args.mode == "bf16x3" and (B, TE, L) == (32, 128, 800):
            tj = json.load(open(tpath))
            traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}
            traffic_src = "profiles/" + name
            break

    def roof(alg, avg_ms, name, tkey):
        ach = alg / (avg_ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": traffic.get(tkey), "traffic_source": traffic_src, "peak_source": peak_kind, "avg_launch_ms": avg_ms,
                "algorithmic_bytes": alg}

    suffix = "_tc_kernel" if args.mode == "bf16x3" else "_kernel"
    r_f = roof(fwd_alg, fwd_avg, "decoder_fwd%s (forward loop, %d steps/launch)" % (suffix, T), "decoder_fwd_tc_kernel")
    r_b = roof(bwd_alg, bwd_avg, "decoder_bwd%s (reverse loop, %d steps/launch)" % (suffix, T), "decoder_bwd_tc_kernel")
    dominant, other = (r_b, r_f) if bwd_avg >= fwd_avg else (r_f, r_b)

    out = None
    if rank == 0:
        own = launches.get("own") if isinstance(launches, dict) else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mode == "fp32" else args.mode, "data": "synthetic",
            "config": workload_config(world), "precision_mode": args.mode,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                    "ms_per_step": max(ms_e2e, wall_e2e) / args.steps},
            "gpu_launches": (own if own else MY_KERNELS_PER_STEP) * args.steps,
            "launches_per_step": launches if launches is not None else {"note": "static count"},
            "clocks": sampler.summary(),
            "roofline": dominant, "roofline_other": other,
            "kernel_share": {"fwd_loop_ms": fwd_avg, "bwd_loop_ms": bwd_avg, "step_ms": ms / args.steps,
                             "outside_loops_ms": ms / args.steps - fwd_avg - bwd_avg,
                             "note": "the weight-gradient products of finished time chunks run BESIDE the reverse loop on the 20 SMs it "
                                     "leaves idle; the loop's duration (and roofline.frac) includes that contention",
                             "without_overlap": no_overlap},
            "ms_per_step_per_rank": per_rank_ms, "allreduce_ms_per_rank": ar_rank,
            "allreduce_bytes": int(tr.flat_g.numel()) * 4 if world > 1 else 0,
            "loss": [float(x) for x in last.tolist()],
        }
    del tr, dev_in
    torch.cuda.empty_cache()
    if not args.no_secondary:
        # BASELINE config 5 beside the decoder number, at every N (short run); configs 3 / 4 on one GPU
        try:
            full = run_full_model(args, dev, pg, world, rank, barrier, min(args.steps, 10))
        except Exception as e:
            full = {"error": repr(e)}
        if rank == 0:
            out["full_model"] = full
        if world == 1:
            out["secondary"] = run_secondary(args, dev, peaks, peak_kind)
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        Ls = 100
        sec = cpu_train_step_oracle(B, TE, Ls, threads, n_timed=1, n_warm=0)
        if sec < 3.5:  # fast host: time the full workload instead (about 8x the sample)
            Ls = L
            sec = cpu_train_step_oracle(B, TE, Ls, threads, n_timed=1, n_warm=0)
        out["cpu_baseline"] = {
            "value": B * Ls / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "B=%d Te=%d L=%d (%d of 801 decoder steps), 1 train step, %.1f s" % (B, TE, Ls, Ls + 1, sec)}
    emit(out)


if __name__ == "__main__":
    main()
