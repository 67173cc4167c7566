"""Multi-rank GPU tests (self-skipping below 2 visible GPUs): the long grid decomposed over ranks with NCCL ghost exchange
must equal the undecomposed single-GPU run bit for bit, and a sweep sharded over ranks must equal the unsharded one.
Runs tools/longgrid_multigpu_check.py under torchrun, the way bench.py --gpus N is launched."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, port=29533, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("mode,overlap", [("lorentz", "0"), ("lorentz_nl", "0"), ("free", "0"), ("lorentz", "1")])
def test_decomposed_long_grid_is_bit_identical_to_one_gpu(mode, overlap):
    """overlap = 1: LongGrid.overlap, every block as inner tiles (while the NCCL messages are in flight) + edge tiles."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    r = _torchrun(world, "tools/longgrid_multigpu_check.py", "--cells", "300000", "--steps", "256", "--mode", mode,
                  env={"PF_LONGGRID_OVERLAP": overlap})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "== single GPU: True" in r.stdout and "ok=True" in r.stdout and "False" not in r.stdout


def test_sharded_reflection_sweep_equals_the_unsharded_one():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, "tools/sweep_multigpu_check.py", port=29534)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
    assert "sharded == unsharded: True" in r.stdout
