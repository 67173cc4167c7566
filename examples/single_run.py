#!/usr/bin/env python
"""One run of each integrator at the reference's default geometry (9 GHz, 0.7 m), as MasterController.__Main__
(:620-667) sets it up, through the drop-in modules; prints wall time and the reflection figures."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import Environment_Setup as envDef, MasterController as MC  # noqa: E402

for name, lor, nl in (("free space / dielectric", False, False), ("Lorentz medium", True, False), ("cubic nonlinear", False, True)):
    setup = envDef.envSetup(9e9, 0.7, 7000, 8000, nonLinMed=nl, LorMed=lor)
    P = MC.Params(*setup, False, 0.7, 9e9, 20)
    V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 10)
    C_P = MC.CPML_Params(P.dz)
    C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
    P.TFSF, P.SineCont, P.Periods = True, True, 1000
    P.LorentzMed, P.nonLinMed, P.FreeSpace = lor, nl, not (lor or nl)
    t0 = time.perf_counter()
    V, P, C_V, C_P, Exs, Hys = MC.Controller(V, P, C_V, C_P)
    dt = time.perf_counter() - t0
    passes = 1 if nl else 2
    print(f"{name:26s} Nz={P.Nz} T={P.timeSteps} passes={passes}: {dt:.3f} s "
          f"({passes * P.timeSteps * (P.Nz + 1) / dt / 1e9:.2f} Gcell-updates/s), max|Ex|={np.max(np.abs(V.Ex)):.4f}")
    if lor:
        t = np.arange(len(V.x1ColBe)) * P.delT
        print(f"{'':26s} reflection measured {MC.results(V, P, C_V, C_P, t, RefCo=True):.4f}, "
              f"analytical {MC.results(V, P, C_V, C_P, t, AnalRefCo=True):.4f}")
