#!/usr/bin/env python
"""BASELINE config 4: an electron beam of macro-particles in the 1-D waveguide grid, coupled to the FDTD fields through
the reference's Jx slot (BaseFDTD11.py:667): every step = one field step with the deposited current subtracted, then the
fused particle step (Boris push + cell re-sort + deterministic deposit, pf_pic_step_sorted).
  python examples/pic_beam.py [particles] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import BaseFDTD11, Environment_Setup as envDef, MasterController as MC, Solver_Engine as SE  # noqa: E402
from pyfdtd_b200 import _device as dev, pic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
P = MC.Params(*envDef.envSetup(9e9, 0.7, 7000, 8000), False, 0.7, 9e9, 20)
P.TFSF, P.SineCont, P.Periods, P.FreeSpace = True, True, 1000, True
V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
C_P, C_V = MC.CPML_Params(P.dz), MC.CPML_Variables(P.Nz, P.timeSteps)
C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False)
L = len(V.Ex)
z, ux, uz, w = pic.make_beam(n, L, P.dz, gamma=1.2, thermal=0.01, seed=1)
beam = pic.ParticleSet(z, ux, uz, w, L, P.dz, P.delT)
grid = dev.DeviceGrid(L=L, T=P.timeSteps, arrays=BaseFDTD11._host_arrays(V, C_V, V.tempVarPol),
                      scalars=BaseFDTD11.grid_scalars(V, P), srcE=np.asarray(Exs) / P.courantNo,
                      srcH=np.asarray(Hys) / P.courantNo, probe_idx=[], flags=BaseFDTD11.grid_flags(P))
sim = pic.CoupledPIC(grid, beam, mode="free", fused=True)
for _ in range(5):
    sim.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    sim.step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
f = grid.fetch(["Ex"], probes=False)
print(f"{n} macro-particles on {L} cells, {steps} coupled steps: {dt * 1e3:.1f} ms = {n * steps / dt:.3e} particle-steps/s; "
      f"max |Ex| {np.max(np.abs(f['Ex'])):.3e} V/m, max |Jx| {beam.Jx.abs().max().item():.3e}; CFL violated: {beam.cfl_violated()}")
