#!/usr/bin/env python
"""BASELINE config 3: frequency x amplitude sweep of the cubic nonlinear slab, every member in one batch on the GPU
(under torchrun the members are dealt round-robin over the ranks, no collective); prints the fundamental and
third-harmonic amplitude at the slab's front face for every member.
  python examples/nonlinear_sweep.py [n_freq] [n_amp] [steps]     (SE.CUBIC = "newton" roughly doubles the speed)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import sweep  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
na = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
freqs, amps = np.linspace(6e9, 10.5e9, nf), np.linspace(0.1, 10.0, na)
t0 = time.perf_counter()
res = sweep.nonlinear_sweep(freqs, amps, 0.7, 7000, 8000, nsteps=steps, rank=rank, world_size=world)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"rank {rank}: {len(res['index'])} of {nf * na} members x {steps} steps in {dt:.2f} s "
      f"({res.get('cell_steps', 0) / dt / 1e9:.1f} Gcell-updates/s incl. host setup)")
if rank == 0:
    print("  f [GHz]   amp   |E(f)| front   |E(3f)| front")
    for j in range(min(len(res["index"]), 12)):
        h = res["harmonic_amplitude"][j]
        print(f"  {res['freq'][j] / 1e9:7.2f} {res['amp'][j]:6.2f} {h[0, 0]:14.6e} {h[1, 0]:14.6e}")
