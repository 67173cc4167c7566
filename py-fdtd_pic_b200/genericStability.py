"""Numerical-dispersion correction of the plasma frequency and resolution guards.

Host-side mirror of ``genericStability.spatialStab`` (genericStability.py:12-62).  It is parity
relevant: the integrators assign its 4th return value to ``V.plasmaFreqE`` at the start of every
pass (Solver_Engine.py:231,286), so the Lorentz coefficients of pass 1 see a twice-corrected value.
The same numpy scalar functions are called in the same order as the reference does, so on a given
machine the result is bit-identical to the reference's.  ``sys.exit()`` guards become ValueError.
"""
from __future__ import annotations

import numpy as np
import scipy.constants as sci


def numerical_wavenumber(dz, freq, dt, wp, w0, gam):
    """|k| of the discretised Lorentz medium at ``freq`` (genericStability.py:27-36)."""
    rad = 2 * np.pi * freq
    c0 = sci.speed_of_light
    sw = np.sinc(np.pi * freq * dt)
    # Drude limit (resonantFreq = 0; the reference's expression divides by w0^2): es*w0^2 -> wp^2 - w0^2 = wp^2.
    # For w0 != 0 the reference's own expression is kept, bit for bit.
    es_w02 = wp ** 2 if w0 == 0 else ((wp ** 2) / (w0 ** 2) - 1) * w0 ** 2
    sqN = (rad ** 2 * sw ** 2 - es_w02 * np.cos(rad * dt) + 1.0j * gam * rad * sw)
    sqD = (rad ** 2 * sw ** 2 - w0 ** 2 * np.cos(rad * dt) + 1.0j * gam * rad * sw)
    arg = (rad / c0) * (dz / 2) * sw * np.sqrt(sqN / sqD)
    return abs((2 / dz) * np.arcsin(arg))


def spatialStab(fullTime, fullSpace, spaceStep, frequency, timeStep, plasmaFreq, resonantFreq, gamma,
                lim_Of_Stability=np.pi / 2):
    frq, tim, sp = frequency, timeStep, spaceStep
    rad = 2 * np.pi * frq
    twoPi = rad / frq
    c0 = sci.speed_of_light
    if rad * tim > np.pi / 2 and frq <= 5e9:
        raise ValueError(f"unstable timestep {rad * tim} {tim}")
    kNum = numerical_wavenumber(sp, frq, tim, plasmaFreq, resonantFreq, gamma)
    epsilon = 1 + ((plasmaFreq ** 2) / (resonantFreq ** 2 - (rad ** 2) - 1j * gamma * rad))
    refr = np.sqrt(abs(np.real(epsilon)))
    fix = (c0 * tim * np.sin((kNum * refr * sp) / 2)) / (refr * sp * np.sin((kNum * c0 * tim) / 2))
    pf = np.sqrt(abs(fix)) * plasmaFreq
    lamCont = c0 / frq
    lamDisc = twoPi / kNum
    diff = abs(lamCont - lamDisc)
    if kNum * sp > np.pi / 2:
        raise ValueError(f"unstable, wave is not resolved: {kNum * sp} {kNum}")
    if kNum * sp * fullSpace < 5 * lamDisc:
        raise ValueError(f"unstable because domain too small, {(lamDisc * 5) / sp}")
    return lamCont, lamDisc, diff, pf, fix
