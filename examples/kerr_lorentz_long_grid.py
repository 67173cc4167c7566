#!/usr/bin/env python
"""BASELINE config 5: one long grid (default 1e8 cells) of dispersive AND nonlinear material -- the Lorentz ADE plus the
cubic Kerr law on Dx - P (PF_LORENTZ_NL) -- streamed with 64-step temporal blocking; under torchrun the grid is
decomposed along z over the ranks (ghost exchange every 64 steps).
  python examples/kerr_lorentz_long_grid.py [cells] [steps]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/kerr_lorentz_long_grid.py 800000000"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import longgrid  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid, info = longgrid.lorentz_long_grid(cells, T=steps, k=64, mode="lorentz_nl", rank=rank, world_size=world)
grid.run(64, do_pol=True)                       # warm-up block
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
grid.run(steps - 64, do_pol=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
probe = grid.probe_out[0, :steps].abs().max().item()
if rank == 0:
    print(f"{cells} cells x {steps - 64} steps on {world} GPU(s), {len(grid.pieces)} pieces: {dt * 1e3:.1f} ms = "
          f"{cells * (steps - 64) / dt / 1e9:.0f} Gcell-updates/s (wave front still in vacuum: quiescent slab cells skip the "
          f"cubic root); max |Ex| at the probe {probe:.4f}")
if world > 1:
    dist.destroy_process_group()
