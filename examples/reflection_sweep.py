#!/usr/bin/env python
"""The reference's headline experiment -- reflection coefficient of the Lorentz half-space vs frequency
(`Reflection vs frequency.png`, produced by MasterController.LoopedSim(loop=True), :530-569) -- run
through the drop-in modules.  All sweep members are advanced together on the GPU.

    python examples/reflection_sweep.py [--points 20] [--low 6e9] [--interval 5e8]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import Environment_Setup as envDef, MasterController as MC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=10)
ap.add_argument("--low", type=float, default=6e9)
ap.add_argument("--interval", type=float, default=5e8)
ap.add_argument("--domain", type=float, default=0.7)
a = ap.parse_args()

# MasterController.__Main__ (:620-663) with LorMed = True
setup = envDef.envSetup(a.low, a.domain, 7000, 8000, LorMed=True)
P = MC.Params(*setup, False, a.domain, a.low, 20)
V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 10)
C_P = MC.CPML_Params(P.dz)
C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
P.TFSF, P.SineCont, P.Gaussian, P.Periods = True, True, False, 1000
P.LorentzMed, P.FreeSpace, P.nonLinMed = True, False, False

t0 = time.perf_counter()
MC.LoopedSim(MC.Reporter(), V, P, C_V, C_P, False, a.domain, 7000, 8000, loop=True, Low=a.low, Interval=a.interval,
             points=a.points)
dt = time.perf_counter() - t0
freqs, measured, analytical = MC.LoopedSim.last_sweep
print(f"{a.points} sweep members in {dt:.2f} s")
print("  f [GHz]   measured   analytical (Fresnel)")
for f, m, an in zip(freqs, measured, analytical):
    print(f"  {f / 1e9:6.2f}    {m:.4f}     {an:.4f}")

# The same physics for MANY frequencies at once -- the product call behind bench.py's `e2e_full_sweep`: every grid is
# envSetup(f, ...) on its own (no member-to-member sizing chain), setup vectorised over members, inputs built natively,
# reflection extracted on the device.
import numpy as np  # noqa: E402
from pyfdtd_b200 import sweep  # noqa: E402

fr = np.linspace(a.low, a.low + a.interval * (a.points - 1), 256)
t0 = time.perf_counter()
res = sweep.reflection_sweep(fr, a.domain, 7000, 8000, periods=1000)
dt = time.perf_counter() - t0
print(f"sweep.reflection_sweep: {len(fr)} distinct grids, {res['cell_steps'] / 1e9:.0f} Gcell-updates in {dt:.2f} s "
      f"(host setup {res['timing']['setup_s'] * 1e3:.0f} ms + input building {res['timing']['build_inputs_s'] * 1e3:.0f} ms)")
print("  R(f) first / last:", res["measured"][0], res["measured"][-1], " Fresnel:", res["analytical"][0], res["analytical"][-1])
