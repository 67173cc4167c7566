#!/usr/bin/env python
"""Run under torchrun on N GPUs: sweep.reflection_sweep sharded over the ranks (member % N == rank, no data-path
collective) must give the numbers of the unsharded sweep (to an ulp of the cuFFT reduction) (rank 0 runs that too and compares)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import sweep  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
freqs = np.linspace(6e9, 10.5e9, 13)
mine = sweep.reflection_sweep(freqs, 0.3, 2000, 2200, periods=1.0, rank=rank, world_size=world)
parts = [None] * world
dist.all_gather_object(parts, (mine["index"], mine["measured"]))
ok = True
if rank == 0:
    full = sweep.reflection_sweep(freqs, 0.3, 2000, 2200, periods=1.0)
    got = np.empty(len(freqs))
    for idx, val in parts:
        got[idx] = val
    # (cuFFT's plan depends on the batch size: the reflection figure agrees to an ulp, the time stepping is bit-identical)
    ok = bool(np.allclose(got, full["measured"], rtol=1e-13, atol=0))
    print(f"world={world} sharded == unsharded: {ok}  max |diff| = {np.max(np.abs(got - full['measured'])):.3e}", flush=True)
    if not ok:
        print("sharded  ", got, "\nunsharded", full["measured"], flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
