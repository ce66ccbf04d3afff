"""CPU, gloo, world_size 2: host-side logic of the multi-GPU path (no CUDA involved)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gloo_world2_plumbing():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29517',
           os.path.join(ROOT, 'tests', 'mp_cpu_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert r.stdout.count(' ok') == 2, tail
