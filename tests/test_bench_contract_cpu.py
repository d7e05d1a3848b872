"""Driver contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`) under torchrun
prints exactly ONE JSON line on stdout, from rank 0 only, with the keys the driver reads; the other rank exits 0 silently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def _check_line(stdout, n_gpus):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["higher_is_better"] is True
    assert d["metric"] == "decoder mel-frames/s (train step)" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["cpu_baseline"] and d["vs_baseline"] is None
    # both arms print the same config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(n_gpus)
    return d


def test_reference_arm_single_process():
    r = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], {"MSTTS_REF_BUDGET_S": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_line(r.stdout, 1)
    import re
    L = int(re.search(r" L=(\d+) ", d["cpu_baseline"]["sample"]).group(1))
    assert 24 <= L <= 200                           # a 2 s budget selects a short sample, never the full 800 frames


def test_reference_arm_under_torchrun_world_2():
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", "29631", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             {"MSTTS_REF_BUDGET_S": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 2)
