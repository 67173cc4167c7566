"""Grid / time-step / probe / CPML sizing from (frequency, domain size).

Host-side mirror of the reference's ``Environment_Setup.envSetup`` (Environment_Setup.py:19-166):
same call signature, same 20-tuple, same integer/float arithmetic, so every downstream index and
step count is identical.  Where the reference prints a message and calls ``sys.exit()`` this raises
``ValueError`` carrying the same message.
"""
from __future__ import annotations

import numpy as np

C0 = 299792458.0  # hard-coded in the reference too (Environment_Setup.py:34)

_TUPLE_FIELDS = ("Nz", "timeSteps", "eLoss", "mLoss", "eSelfCo", "eHcompsCo", "hSelfCo", "hEcompsCo",
                 "x1Loc", "x2Loc", "materialFrontEdge", "materialRearEdge", "pmlWidth", "nzsrc", "lamMin",
                 "dz", "delT", "courantNo", "period", "Nlam")


def _points_per_wavelength(VExists, V, P, nonLinMed):
    """Nlam rule, Environment_Setup.py:23-46."""
    if not VExists:
        return 350 if nonLinMed else 400
    w = 2 * np.pi * P.freq_in
    eps = 1 + (V.plasmaFreqE * V.plasmaFreqE) / (V.omega_0E * V.omega_0E - (w * w) + 1j * V.gammaE * w)
    scale = 200 if nonLinMed else 60
    return int(scale * (np.real(eps)) ** 1.05)


def _pick_time_steps(freq_in, delT, minim, maxim):
    """First N in [minim, maxim) that puts freq_in on an FFT bin, else maxim-1 (:77-116)."""
    if minim == maxim:
        minim -= 1
    chosen = None
    for N in range(minim, maxim):
        chosen = N
        bin_pos = (freq_in * N) / (1 / delT)
        if int(bin_pos) - bin_pos == 0:
            break
    if chosen is None or chosen < minim:
        chosen = minim
    return chosen


def envSetup(newFreq_in, domainSize, minim=400, maxim=600, VExists=False, V=[], P=[], nonLinMed=False,
             LorMed=False):
    freq_in = newFreq_in
    lamMin = C0 / freq_in
    Nlam = _points_per_wavelength(VExists, V, P, nonLinMed)
    dz = lamMin / Nlam
    delT = (dz / C0) * 0.95
    period = 1 / freq_in
    courantNo = (C0 * delT) / dz
    if courantNo > 3 or courantNo < 0:
        raise ValueError(f"{courantNo} courantNo is unstable")
    pmlWidth = 6 * int(lamMin / dz)
    if pmlWidth >= 12000:
        raise ValueError(f"pmlWidth too big {pmlWidth}")
    Nz = int(domainSize / dz) + 2 * pmlWidth

    timeSteps = _pick_time_steps(freq_in, delT, minim, maxim)
    timeSteps += int(timeSteps * (Nlam / 200))
    if timeSteps >= 2 ** 15:
        raise ValueError("timeSteps too large")

    srcOffset = int(0.05 / dz)
    if srcOffset >= Nz * 0.65:
        raise ValueError(f"{srcOffset} src is too far into domain")
    nzsrc = srcOffset + pmlWidth
    if nzsrc - 10 <= pmlWidth:
        raise ValueError("The probe for fft is in the PML region")

    slabOffset = int(0.1 / dz)
    materialFrontEdge = slabOffset + pmlWidth
    materialRearEdge = Nz - 1           # half-space: the slab runs into the right CPML
    if materialRearEdge - materialFrontEdge < 10:
        raise ValueError(f"{materialRearEdge - materialFrontEdge} width is too small or negative")
    if slabOffset >= domainSize / dz:
        raise ValueError("Material starts in CPML region")
    if materialFrontEdge <= nzsrc:
        raise ValueError("Source is inside material")

    eLoss = 0
    mLoss = 0
    out = dict(Nz=Nz, timeSteps=timeSteps, eLoss=eLoss, mLoss=mLoss, eSelfCo=(1 - eLoss) / (1 + eLoss),
               eHcompsCo=1 / (1 + eLoss), hSelfCo=(1 - mLoss) / (1 + mLoss), hEcompsCo=1 / (1 + mLoss),
               x1Loc=materialFrontEdge - 20, x2Loc=nzsrc - 100, materialFrontEdge=materialFrontEdge,
               materialRearEdge=materialRearEdge, pmlWidth=pmlWidth, nzsrc=nzsrc, lamMin=lamMin, dz=dz,
               delT=delT, courantNo=courantNo, period=period, Nlam=Nlam)
    return tuple(out[k] for k in _TUPLE_FIELDS)
