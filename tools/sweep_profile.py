#!/usr/bin/env python
"""Host-side profile of a reference-style frequency sweep (LoopedSim, default geometry) -- where a real
sweep spends its wall time once stepping is on the GPU.  Run on the GPU box."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import pyfdtd_b200  # noqa: F401,E402
from pyfdtd_b200 import sweep  # noqa: E402
from test_host_layer import build_objects  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 20
spec = dict(mode="lorentz", freq=6e9, dom=0.7, win=[7000, 8000], source="sine", periods=1.0, epsRe=1.0)


def once():
    V, P, C_V, C_P = build_objects(spec)
    t0 = time.perf_counter()
    out = sweep.frequency_sweep(V, P, 0.7, 7000, 8000, Low=6e9, Interval=2e8, points=points)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


once()
dt, out = once()
print("points", points, "sweep seconds", round(dt, 4))
print("measured R[:5]", out[1][:5], "analytical R[:5]", out[2][:5])
pr = cProfile.Profile()
pr.enable()
once()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(32)
