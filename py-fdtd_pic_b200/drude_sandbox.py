"""The reference's Drude scratch script (TESTBOXDIPSERSE.py) as a function.

The reference steps a Drude medium only here: a current-form (J) update with a hard source and no PML, written as plain
Python loops that run at import and plot the final field (TESTBOXDIPSERSE.py:22-120).  ``run()`` reproduces the script's
numbers -- same constants, same source table, same loop order, including Python's ``Hy[-1]`` wrap in the E update of cell 0
-- with the time loop (:79-94) executed by ``pf_drude_j_run`` (csrc/pf_dormant.cu) on the GPU.  Bit-identical to the script
(tests/golden/drude_sandbox*.npz, made by executing the unmodified script text).  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat


def constants(domain=14000, tim=5000, freq=0.08e9, nl=700, src=2100, matFront=8000):
    """TESTBOXDIPSERSE.py:22-56, evaluated in the script's order."""
    dz = (3e8 / (freq)) * (1 / nl)
    dt = (dz / 3e8)
    cour = (3e8 * dt) / dz
    matRear = domain - 1
    gammaE = 0
    plasmaE = 2 * np.pi * 1e9
    perm0 = 8.854e-12
    betaE = (0.5 * plasmaE ** 2 * perm0 * dt) / (1 + 0.5 * gammaE * dt)
    kapE = (1 - 0.5 * gammaE * dt) / (1 + 0.5 * gammaE * dt)
    ppw = 3e8 / (freq * dz)
    Exs, Hys = [], []
    for timer in range(tim):
        if (timer * dt < 1 / freq):
            Exs.append(float(np.sin(2.0 * np.pi / ppw * (cour * timer))))
            Hys.append(float(np.sin(2.0 * np.pi / ppw * (cour * (timer + 1)))))
        elif (timer * dt >= 1 / freq):
            Exs.append(0)
            Hys.append(0)
    for boo in range(tim):
        if (Hys[boo] == 0):
            Hys[boo - 1] = 0
    return dict(domain=domain, tim=tim, dz=dz, dt=dt, cour=cour, src=src, matFront=matFront, matRear=matRear, perm0=perm0,
                betaE=float(betaE), kapE=float(kapE), Exs=np.asarray(Exs, dtype=np.float64), Hys=np.asarray(Hys, dtype=np.float64))


def run(domain=14000, tim=5000, freq=0.08e9, nl=700, src=2100, matFront=8000):
    """Final (Ex, Hy, Jx) of the script for the given sizes (defaults = the script's own)."""
    torch = nat.require_cuda()
    k = constants(domain, tim, freq, nl, src, matFront)
    z = lambda: torch.zeros(domain, dtype=torch.float64, device="cuda")
    Ex, Hy, Jx, tempE, tempEOld = z(), z(), z(), z(), z()
    Hys = torch.as_tensor(k["Hys"], device="cuda")
    d = nat.PfDrudeJ()
    d.n, d.src, d.mat_front, d.mat_rear, d.n_src = domain, src, matFront, k["matRear"], tim
    perm0, betaE, dt, cour, kapE = k["perm0"], k["betaE"], k["dt"], k["cour"], k["kapE"]
    d.inv_cour = 1 / cour
    d.kapE, d.betaE = kapE, betaE
    d.c_self = (2 * perm0 - betaE * dt) / (2 * perm0 + betaE * dt)
    d.c_curl = (2 * dt) / (2 * perm0 + betaE * dt)
    d.half_one_plus_kap = 0.5 * (1 + kapE)
    d.Ex, d.Hy, d.Jx, d.tempE, d.tempEOld, d.Hys = (t.data_ptr() for t in (Ex, Hy, Jx, tempE, tempEOld, Hys))
    nat.check(nat.lib().pf_drude_j_run(ctypes.byref(d), 0, tim, nat.current_stream_ptr()), "pf_drude_j_run")
    return Ex.cpu().numpy(), Hy.cpu().numpy(), Jx.cpu().numpy()
