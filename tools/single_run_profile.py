#!/usr/bin/env python
"""Host-side profile of one default-geometry Controller() run (run on the GPU box)."""
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
from pyfdtd_b200 import MasterController as MC, Solver_Engine as SE  # noqa: E402
from test_host_layer import build_objects  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "lorentz"
spec = dict(mode=mode, freq=9e9, dom=0.7, win=[7000, 8000], source="sine", periods=1000, epsRe=4.0 if mode == "free" else 1.0)


def once():
    V, P, C_V, C_P = build_objects(spec)
    t0 = time.perf_counter()
    MC.Controller(V, P, C_V, C_P)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, P


for _ in range(2):
    dt, P = once()
print("Nz", P.Nz, "T", P.timeSteps, "Controller seconds", dt, SE.LAST_RUN_INFO)
V, P, C_V, C_P = build_objects(spec)
pr = cProfile.Profile()
pr.enable()
MC.Controller(V, P, C_V, C_P)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

# ---- device time of the tile kernel alone -----------------------------------------------------
from pyfdtd_b200 import _native as nat  # noqa: E402
lib = nat.lib()
V, P, C_V, C_P = build_objects(spec)
lib.pf_profile_enable(1)
t0 = time.perf_counter()
MC.Controller(V, P, C_V, C_P)
wall = time.perf_counter() - t0
ms, n = nat.c_double(), nat.c_int()
lib.pf_profile_collect(ms, n)
lib.pf_profile_enable(0)
steps = 2 * P.timeSteps
print("k_tile: %d launches, %.3f ms total, %.2f us/launch, %.3f us/step (%.0f cycles at 1.9 GHz); wall %.1f ms"
      % (n.value, ms.value, 1e3 * ms.value / n.value, 1e3 * ms.value / steps, 1.9e3 * 1e3 * ms.value / steps / 1e3, wall * 1e3))
