"""Multi-GPU parity of the frame pipeline (b200r_pipeline_*): tools/dist_check.py under torch.distributed.run, one rank per GPU.
Every rank compares every assembled frame - NCCL all-gather assembly and peer-push assembly, 3 frames in flight - bit for bit
with the same frame rendered whole on its own GPU. Needs at least 2 GPUs on the box (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(n, *args):
    port = 29500 + (os.getpid() % 400) + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), *args]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "on all ranks: 0" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("args", [("c2",), ("c2", "mlaa")])
def test_assembled_frames_equal_whole_frames(n, args):
    if _gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    _run(n, *args)


def test_c5_assembled_frames_equal_whole_frames():
    n = min(8, _gpus())
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _run(n, "c5")
