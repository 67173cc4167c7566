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


def ref_tester_batch(traces, time_steps, keep_from=None, keep_to=None, check=True):
    """RefTester's scalar for a whole batch, on whatever device ``traces`` lives on (SURVEY section 8f row 1:
    the sweep's probe traces never leave the GPU).

    traces: torch tensor [members, >= time_steps] float64, one probe trace per row (only the first
    ``time_steps`` samples are used -- no zero padding, the bins must be the reference's).  Samples with
    index < keep_from[m] or > keep_to[m] are zeroed first (the reference's x1ColAf / x1ColBe read windows,
    Solver_Engine.py:360-368).  Returns (val, idx): val[m] = 2|FFT(y_m)|/time_steps at the largest bin,
    idx[m] that bin.  Raises ValueError like RefTester if the largest bin of any member is DC.
    """
    import torch
    T = int(time_steps)
    y = traces[:, :T]
    if keep_from is not None or keep_to is not None:
        n = torch.arange(T, device=y.device)[None, :]
        keep = torch.ones(y.shape, dtype=torch.bool, device=y.device)
        if keep_from is not None:
            keep &= n >= torch.as_tensor(keep_from, device=y.device).reshape(-1, 1)
        if keep_to is not None:
            keep &= n <= torch.as_tensor(keep_to, device=y.device).reshape(-1, 1)
        y = torch.where(keep, y, torch.zeros((), dtype=y.dtype, device=y.device))
    mag = torch.fft.rfft(y, dim=1).abs()          # |Y[k]| = |Y[T-k]|: the first maximum is in the half spectrum
    idx = torch.argmax(mag, dim=1)
    if check and bool((idx == 0).any()):      # check=False: the caller tests idx itself (keeps the stream asynchronous)
        raise ValueError("Could not find non-DC freq")
    val = 2.0 * mag.gather(1, idx[:, None])[:, 0] / T
    return val, idx
