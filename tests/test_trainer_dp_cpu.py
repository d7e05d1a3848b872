"""World-size-2 gloo test (CPU) of the data-parallel host logic used by trainer.DecoderTrainer: one all-reduce of the flat
gradient buffer, 1/world scale, identical TF-style Adam on every rank.  The CUDA kernels are replaced by the
oracle's Adam here -- this test covers the collective plumbing, not the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import decoder_oracle as O
    from multi_speaker_tts_b200 import trainer
    torch.manual_seed(0)
    p = torch.randn(1000)            # identical replica on every rank
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    g_local = torch.randn(1000, generator=torch.Generator().manual_seed(100 + rank))  # per-shard gradient
    flat_g = g_local.clone()
    dist.all_reduce(flat_g)          # the single gradient all-reduce of the step
    lr = trainer.learning_rate(0)
    O.tf_adam_step(p, m, v, flat_g / world, 1, lr, eps=1e-6)
    gathered = [torch.zeros_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    if rank == 0:
        g_mean = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        p_ref = torch.randn(1000, generator=torch.manual_seed(0))
        torch.save({"same": all(torch.equal(gathered[0], x) for x in gathered), "p": p, "g_mean": g_mean, "flat": flat_g / world}, out)
    dist.destroy_process_group()


def test_allreduce_mean_and_identical_replicas(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same"], "replicas diverged after the step"
    assert torch.allclose(r["flat"], r["g_mean"], atol=1e-6), "all-reduce / world != mean of the per-shard gradients"


def test_learning_rate_schedule_matches_reference_formula():
    from multi_speaker_tts_b200 import trainer
    assert trainer.learning_rate(0) == 1e-3
    assert abs(trainer.learning_rate(10000) - 5e-4) < 1e-12      # 0.5 ** (step / 10000), MSTTS_SV.py:163-169
    assert trainer.learning_rate(10 ** 7) == 1e-5                 # clipped at Min
