"""The multi-process NCCL path that bench.py times at N > 1, checked against the CPU oracle
(tests/mp_worker_nccl.py has the protocol).  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_workers(n, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mp_worker_nccl.py"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK:" in r.stdout
    return r.stdout


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_stream_equals_cpu_reference():
    run_workers(2, ["--frames", "32", "--res", "0.005"])


@pytest.mark.skipif(_gpus() < 4, reason="needs 4 GPUs")
def test_four_rank_nccl_stream_equals_cpu_reference():
    run_workers(4, ["--frames", "22", "--res", "0.005", "--first-frame", "100"])
