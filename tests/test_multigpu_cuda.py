"""The multi-GPU steppers against the single-GPU stepper on a box with >= 2 GPUs (skipped otherwise): spawns
torchrun with 2 ranks on tools/check_slab.py, which asserts rel Linf < 1e-10 on the vorticity after a few steps."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["rows", "rows-eager", "rows-periodic", "z"])
def test_two_rank_slab_stepper_matches_one_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", "tools/check_slab.py", "1024", "4", mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "vs single GPU" in r.stdout
