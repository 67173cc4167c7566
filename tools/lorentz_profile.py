#!/usr/bin/env python
"""A few launches of the bench's Lorentz sweep batch in one arithmetic mode, for ncu captures of k_tile.
Usage: python tools/lorentz_profile.py [exact|fma|fp32] [members] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pyfdtd_b200 import Solver_Engine as SE  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "exact"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
b, table = bench.lorentz_sweep_batch(M, S, 64, fma=variant == "fma", fp32=variant == "fp32")
b.upload()
b.randomize_state(seed=1234)


def step():
    b.reset_state(template=True)
    b.run(do_pol=True)


sec = bench.time_cuda(torch, step, 2)
print({"variant": variant, "members": M, "steps": S, "Gcell_updates_per_s": b.cell_steps / sec / 1e9})
