"""CPU: tf.train.Saver semantics of the trainers' checkpoints (MSTTS_SV.py:30-40,244-251,287-289): CHECKPOINT-<step> files,
max_to_keep = 5 rotation, the `checkpoint` state file and latest_checkpoint; under world size 2 (gloo) only rank 0 writes and
every rank leaves Save() after the file exists."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_speaker_tts_b200 import checkpoint


def test_rotation_keeps_five_and_state_file_names_the_latest(tmp_path):
    d = str(tmp_path / "ckpt")
    assert checkpoint.latest_checkpoint(d) is None
    for step in (1000, 2000, 3000, 4000, 5000, 6000, 7000):
        path = checkpoint.save(d, {'global_step': step, 'x': torch.full((3,), float(step))}, step, max_to_keep=5)
        assert os.path.basename(path) == 'CHECKPOINT-%d.pt' % step
        assert checkpoint.latest_checkpoint(d) == path
    files = sorted(f for f in os.listdir(d) if f.endswith('.pt'))
    assert files == ['CHECKPOINT-%d.pt' % s for s in (3000, 4000, 5000, 6000, 7000)]
    state = open(os.path.join(d, 'checkpoint')).read().splitlines()
    assert state[0] == 'model_checkpoint_path: "CHECKPOINT-7000.pt"'
    assert state[1:] == ['all_model_checkpoint_paths: "CHECKPOINT-%d.pt"' % s for s in (3000, 4000, 5000, 6000, 7000)]
    blob = torch.load(checkpoint.latest_checkpoint(d))
    assert blob['global_step'] == 7000 and torch.equal(blob['x'], torch.full((3,), 7000.0))
    assert not [f for f in os.listdir(d) if f.endswith('.tmp')]


def test_latest_checkpoint_without_state_file_and_legacy_name(tmp_path):
    d = str(tmp_path / "c2")
    os.makedirs(d)
    torch.save({'global_step': 1}, os.path.join(d, 'CHECKPOINT.pt'))           # round-1 layout
    assert checkpoint.latest_checkpoint(d).endswith('CHECKPOINT.pt')
    torch.save({'global_step': 20}, os.path.join(d, 'CHECKPOINT-20.pt'))
    torch.save({'global_step': 100}, os.path.join(d, 'CHECKPOINT-100.pt'))      # numeric, not lexicographic, order
    assert checkpoint.latest_checkpoint(d).endswith('CHECKPOINT-100.pt')
    # re-saving an existing step does not duplicate it in the state file
    checkpoint.save(d, {'global_step': 100}, 100)
    checkpoint.save(d, {'global_step': 100}, 100)
    state = open(os.path.join(d, 'checkpoint')).read()
    assert state.count('CHECKPOINT-100.pt') == 2    # once as latest, once in the list


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, d):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = dist.group.WORLD
    # every rank calls Save with ITS OWN tensor; only rank 0's may land on disk, and nobody returns before it exists
    path = checkpoint.save(d, {'rank': rank, 'x': torch.full((4,), float(rank))}, 42, process_group=pg)
    assert os.path.exists(path)
    blob = torch.load(path)
    assert blob['rank'] == 0
    dist.destroy_process_group()


def test_only_rank0_writes_under_data_parallel(tmp_path):
    d = str(tmp_path / "dp")
    mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
    assert sorted(os.listdir(d)) == ['CHECKPOINT-42.pt', 'checkpoint']
