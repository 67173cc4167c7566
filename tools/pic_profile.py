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


def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ps.sort()
t_ps = timed(lambda: ps.push_sorted(Ex, Hy))
t_dep = timed(ps.deposit)
t_push = timed(lambda: ps.push(Ex, Hy))
ps.sort()
t_fused = timed(lambda: ps.step_sorted(Ex, Hy))
print("fused step (push + re-sort + deposit in the move pass) %.3f ms -> %.3e particle-steps/s" % (t_fused, n / t_fused * 1e3))
ps.sort()
print("n=%d  push_sorted %.3f ms  deposit %.3f ms  -> %.3e particle-steps/s;  plain push %.3f ms (%.3e /s)"
      % (n, t_ps, t_dep, n / (t_ps + t_dep) * 1e3, t_push, n / t_push * 1e3))
