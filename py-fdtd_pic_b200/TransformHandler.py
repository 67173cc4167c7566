"""Reflection extraction from probe traces.

Mirror of ``TransformHandler.RefTester`` (TransformHandler.py:34-78), the only function of that
module on the path ("next" row 1): FFT the trace, scale by 2/timeSteps, return
the value of the largest non-DC bin.  Plotting side effects of the reference are dropped.
"""
from __future__ import annotations

import numpy as np
from scipy import fftpack


def RefTester(V, P, y, m, mul=False):
    Y = fftpack.fft(y)
    Ypow = (2 * np.abs(Y)) / P.timeSteps
    freqs = fftpack.fftfreq(len(y), d=P.delT)
    indMax = int(np.argmax(Ypow))
    if indMax == 0:
        raise ValueError("Could not find non-DC freq")
    val = Ypow[indMax]
    half = int(len(Y) / 2)
    tail = (np.abs(Y[half + 10:len(Y) - 500]) * m * 2) / len(Y)
    return tail, freqs[half + 10:len(freqs) - 500], val


def reflection_spectrum(x1ColBe, x1ColAf, delT, fmin=None, fmax=None):
    """|FFT(reflected)| / |FFT(incident)| per bin (BASELINE config 2: reflection-vs-frequency from a
    single broadband pulse).  Generalises RefTester's single-peak ratio."""
    Yb = np.abs(fftpack.fft(np.asarray(x1ColBe)))
    Ya = np.abs(fftpack.fft(np.asarray(x1ColAf)))
    f = fftpack.fftfreq(len(x1ColBe), d=delT)
    sel = f > 0
    if fmin is not None:
        sel &= f >= fmin
    if fmax is not None:
        sel &= f <= fmax
    return f[sel], Ya[sel] / Yb[sel]
