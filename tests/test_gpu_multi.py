"""Row-sharded online solve: the stepping path on one GPU, and over NCCL when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("eps", ["0.05", "0.01"])
def test_stepping_path_single_gpu(eps):
    """The stepping entry points + one CUDA graph per batch on one GPU (no collective); eps 0.01 = precise operands."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_sharded_pair.py"), "1500", "1637", eps],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


@pytest.mark.parametrize("eps", ["0.05", "0.01"])
def test_row_sharded_two_gpus_nccl(eps):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if eps == "0.05" else "29534",
           os.path.join(ROOT, "tests", "run_sharded_pair.py"), "3000", "3301", eps]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("OK") == 2


def test_sharded_median_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tests", "run_sharded_median.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("OK") == 6 and "FAIL" not in out.stdout
