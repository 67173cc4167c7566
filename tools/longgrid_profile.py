#!/usr/bin/env python
"""A few blocks of one long grid (BASELINE configs 1 / 5) from a synthetic non-zero state, for ncu captures of the
long-grid variants of k_tile.  Usage: python tools/longgrid_profile.py [free|lorentz|lorentz_nl|nl] [cells]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pyfdtd_b200 import longgrid, sweep  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "lorentz"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
steps = 128
grid, info = longgrid.lorentz_long_grid(cells, T=steps * 6, k=64, mode=mode)
gen = torch.Generator(device="cuda").manual_seed(1234)
for arrs0, arrs1 in zip(*grid.bufs):
    for n, t in arrs0.items():
        if t is not None:
            t.copy_((torch.rand(t.shape, dtype=t.dtype, device=t.device, generator=gen) * 2 - 1) * sweep.MemberBatch.STATE_SCALE[n])
            arrs1[n].copy_(t)
sec = bench._time_cuda(torch, lambda: grid.run(steps, do_pol=(mode in ("lorentz", "lorentz_nl"))), 2)
print({"mode": mode, "cells": cells, "steps": steps, "Gcell_updates_per_s": cells * steps / sec / 1e9})
