#!/usr/bin/env python
"""Three PIC steps (fused push+re-sort, deposit) on 2e7 particles, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyfdtd_b200  # noqa
from pyfdtd_b200 import pic
L, dz, dt = 13194, 8.3276e-5, 2.6389e-13
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
z, ux, uz, w = pic.make_beam(n, L, dz, seed=1)
ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
Ex = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 1e5
Hy = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 3e2
ps.sort()
for _ in range(3):
    ps.push_sorted(Ex, Hy)
    ps.deposit()
for _ in range(2):
    ps.push(Ex, Hy)
    ps.deposit()
torch.cuda.synchronize()
print("ok")
