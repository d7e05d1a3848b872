"""tf.train.Saver semantics for the trainers (MSTTS_SV.py:30-40,244-251,287-289; WaveGlow/WaveGlow.py):
``Saver.save(sess, '<dir>/CHECKPOINT', global_step=n)`` writes ``CHECKPOINT-<n>`` files, keeps the newest ``max_to_keep`` (5)
and records them in the text state file ``<dir>/checkpoint`` (``model_checkpoint_path`` / ``all_model_checkpoint_paths``
lines); ``tf.train.latest_checkpoint(dir)`` returns what that state file names.  Here a checkpoint is one ``torch.save``
blob ``CHECKPOINT-<n>.pt`` (variables keyed by their TF names, SURVEY A-8), written to a temporary file and renamed so that a
reader never sees a torn file.  Under data parallel only rank 0 writes; the others wait at a barrier."""
import os
import re

import torch

STATE_FILE = 'checkpoint'
PREFIX = 'CHECKPOINT'


def _norm(path):
    return path.replace("\\", "/")


def _read_state(directory):
    path = os.path.join(_norm(directory), STATE_FILE)
    if not os.path.exists(path):
        return None, []
    latest, every = None, []
    with open(path, 'r') as f:
        for line in f:
            m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths)\s*:\s*"(.*)"\s*$', line)
            if not m:
                continue
            if m.group(1) == 'model_checkpoint_path':
                latest = m.group(2)
            else:
                every.append(m.group(2))
    return latest, every


def _write_state(directory, latest, every):
    path = os.path.join(_norm(directory), STATE_FILE)
    tmp = path + '.tmp'
    with open(tmp, 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % latest)
        for name in every:
            f.write('all_model_checkpoint_paths: "%s"\n' % name)
    os.replace(tmp, path)


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the path named by the state file (None when there is none).  Directories written before the
    state file existed (a bare CHECKPOINT.pt) and directories whose state file was lost are still found."""
    directory = _norm(directory)
    latest, _ = _read_state(directory)
    if latest is not None:
        path = os.path.join(directory, latest)
        if os.path.exists(path):
            return path
    if os.path.isdir(directory):
        steps = []
        for name in os.listdir(directory):
            m = re.match(r'^%s-(\d+)\.pt$' % PREFIX, name)
            if m:
                steps.append((int(m.group(1)), name))
        if steps:
            return os.path.join(directory, max(steps)[1])
        legacy = os.path.join(directory, PREFIX + '.pt')
        if os.path.exists(legacy):
            return legacy
    return None


def save(directory, blob, global_step, max_to_keep=5, process_group=None):
    """Saver.save(..., global_step=global_step) with max_to_keep rotation.  Returns the checkpoint path (on every rank)."""
    directory = _norm(directory)
    name = '%s-%d.pt' % (PREFIX, int(global_step))
    path = os.path.join(directory, name)
    rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
    if rank == 0:
        os.makedirs(directory, exist_ok=True)
        tmp = path + '.tmp'
        torch.save(blob, tmp)
        os.replace(tmp, path)
        _, every = _read_state(directory)
        every = [n for n in every if n != name and os.path.exists(os.path.join(directory, n))] + [name]
        while max_to_keep and len(every) > max_to_keep:
            old = every.pop(0)
            try:
                os.remove(os.path.join(directory, old))
            except OSError:
                pass
        _write_state(directory, name, every)
    if process_group is not None:
        torch.distributed.barrier(group=process_group)
    return path
