#!/usr/bin/env python
"""Headline benchmark: decoder mel-frames/s of one Tacotron2-decoder TRAIN step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode fp32]

A step = masks -> decoder forward loop -> loss -> reverse loop -> weight-gradient GEMMs -> (one NCCL
all-reduce of the flat gradient buffer when N > 1) -> TF-style Adam, on synthetic inputs of BASELINE config 2
(B=32 per GPU, text_len=128, mel_len=800, 80-mel).  Rank 0 prints ONE JSON line.  `value` has the inputs
resident in HBM; `e2e` adds the host->device copy of each step's inputs (pinned) and a device->host read of
the loss.  `--impl reference` times the CPU restatement of the reference (oracle/, the reference itself is
TF1 and cannot run here) on a bounded sample of the same workload.
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
MY_KERNELS_PER_STEP = 44  # counted in profiles/r1_launches_bf16x3.csv between two forward loops: the 2 persistent loops, 2 mask fills,
# 9 hi/lo split kernels, 12 column-sum kernels, 4 prenet activation kernels, loss, Adam x2 and 12 prep / epilogue kernels
# (library GEMMs -- 48 launches per step -- are not counted)


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
        "config": {"workload": "Tacotron2 decoder train step B=32 text_len=128 mel_len=800 80-mel (BASELINE config 2)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is TF1 (cannot run in this image); this is the PyTorch-CPU restatement in oracle/",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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

    B, T = B_PER_GPU, L + 1
    tr = DecoderTrainer(dev, mem_dim=D, mode=args.mode, seed=0, process_group=pg)
    host = S.synthetic_decoder_batch(B, TE, L, seed=1234, rank=rank)
    pinned = {k: host[k].pin_memory() for k in ('memory', 'text_len', 'mel', 'mel_len')}
    dev_in = {k: v.to(dev) for k, v in pinned.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return tr.train_step(dev_in['memory'], dev_in['text_len'], dev_in['mel'], dev_in['mel_len'], T)

    def step_e2e():
        cur = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        loss2 = tr.train_step(cur['memory'], cur['text_len'], cur['mel'], cur['mel_len'], T)
        return loss2.cpu()  # device->host read of the step's result (synchronises)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss2 = step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    fwd_ms, nf = kernel_ms(0)
    bwd_ms, nb = kernel_ms(1)
    set_profiling(False)
    # end-to-end arm (host buffers, copies inside the timed region)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        last = step_e2e()
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        tmax = torch.tensor([ms, ms_e2e, wall_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e, wall_e2e = tmax.tolist()
    frames = world * B * L * args.steps
    value = frames / (ms * 1e-3)
    e2e_val = frames / (max(ms_e2e, wall_e2e) * 1e-3)

    peaks, peak_kind = measured_peaks()
    hbm = float(peaks["hbm_gbs"])
    s_w = 4 if args.mode == "fp32" else (4 if args.mode == "bf16x3" else 2)
    fwd_alg = fwd_bytes_per_step(B, TE, s_w, 4) * T
    bwd_alg = (fwd_bytes_per_step(B, TE, s_w, 4) + saved_bytes_per_step(B, TE)) * T
    fwd_avg, bwd_avg = fwd_ms / max(nf, 1), bwd_ms / max(nb, 1)

    # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture of this command at the
    # default workload (profiles/r1_traffic.json); null for any other shape / mode
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and args.mode == "bf16x3" and (B, TE, L) == (32, 128, 800):
        tj = json.load(open(tpath))
        traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}

    def roof(alg, avg_ms, name, tkey):
        ach = alg / (avg_ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": traffic.get(tkey), "peak_source": peak_kind, "avg_launch_ms": avg_ms, "algorithmic_bytes": alg}

    r_f = roof(fwd_alg, fwd_avg, "decoder_fwd_kernel (forward loop, %d steps/launch)" % T, "decoder_fwd_tc_kernel")
    r_b = roof(bwd_alg, bwd_avg, "decoder_bwd_kernel (reverse loop, %d steps/launch)" % T, "decoder_bwd_tc_kernel")
    dominant, other = (r_b, r_f) if bwd_avg >= fwd_avg else (r_f, r_b)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mode == "fp32" else args.mode, "data": "synthetic",
            "config": {
                "workload": "Tacotron2 decoder train step (BASELINE config 2): B=%d/GPU text_len=%d mel_len=%d 80-mel, "
                            "%d decoder steps, fwd+loss+bwd+TF-Adam%s" % (B, TE, L, T, "+1 NCCL allreduce" if world > 1 else ""),
                "precision_mode": args.mode, "parallelism": "dp%d" % world,
                "l2_policy": "per-step working set (saved activations + weights, ~2.4 GB) exceeds the 126 MB L2; no flush",
            },
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                    "ms_per_step": max(ms_e2e, wall_e2e) / args.steps},
            "gpu_launches": MY_KERNELS_PER_STEP * args.steps,
            "clocks": sampler.summary(),
            "roofline": dominant, "roofline_other": other,
            "kernel_share": {"fwd_loop_ms": fwd_avg, "bwd_loop_ms": bwd_avg, "step_ms": ms / args.steps},
            "loss": [float(x) for x in last.tolist()],
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            Ls = 100
            sec = cpu_train_step_oracle(B, TE, Ls, threads, n_timed=1, n_warm=0)
            if sec < 3.5:  # fast host: time the full workload instead (about 8x the sample)
                Ls = L
                sec = cpu_train_step_oracle(B, TE, Ls, threads, n_timed=1, n_warm=0)
            out["cpu_baseline"] = {
                "value": B * Ls / sec, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "B=%d Te=%d L=%d (%d of 801 decoder steps), 1 train step, %.1f s" % (B, TE, Ls, Ls + 1, sec)}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
