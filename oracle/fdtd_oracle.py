"""CPU ORACLE for the Py-FDTD_PIC time-stepping hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement of the reference algorithm: host-side setup in NumPy / ``math`` (this file) and the
per-step loops in plain C (``fdtd_oracle.c``, built by ``oracle/Makefile`` into
``oracle/_build/liboracle.so``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import this module; the product package never does.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the *unmodified* reference run in the
build container through ``oracle/ref_shim.py`` (generator: ``oracle/make_golden.py``); the oracle is
checked against every one of them in ``tests/test_oracle_vs_golden.py``.  (The reference's own tests
hold no usable vectors -- SURVEY.md section 0, F6.)

Rounding notes that matter for fp64 parity:
  * jitclass members typed float32 are quantised on assignment: Params.courantNo / Amplitude /
    Periods / epsRe / muRe (MasterController.py:217-284), Variables.alpha3 (:129),
    CPML_Params.kappaMax / r_scale / r_a_scale / alphaMax (:341-347).
  * numba-compiled setup functions call libm (``exp``/``pow``), NOT numpy's SIMD kernels; they are
    restated here with ``math.exp`` / ``math.pow`` element by element.  Functions the reference
    runs in plain Python with numpy (Gaussian, SmoothTurnOn, spatialStab, Sig_Mod) are restated
    with the same numpy calls.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import scipy.constants as sci

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")

C0 = 299792458.0          # Params.c0, MasterController.py:294 ; Environment_Setup.py:34
CHAR_IMP = 376.730313668  # Params.CharImp, MasterController.py:293
EPS0 = sci.epsilon_0      # Params.permit_0, MasterController.py:291
MU0 = sci.mu_0            # Params.permea_0, MasterController.py:292


def f32(x) -> float:
    return float(np.float32(x))


# ----------------------------------------------------------------------------- setup: geometry
def env_setup(freq_in, domainSize, minim=400, maxim=600, *, eps_for_nlam=None, nonLinMed=False):
    """Environment_Setup.py:19-166.  ``eps_for_nlam`` = complex permittivity when VExists."""
    lamMin = C0 / freq_in
    if eps_for_nlam is not None:                                   # :38-42
        Nlam = int(60 * (np.real(eps_for_nlam)) ** 1.05)
        if nonLinMed:
            Nlam = int(200 * (np.real(eps_for_nlam)) ** 1.05)
    else:                                                          # :43-46
        Nlam = 350 if nonLinMed else 400
    dz = lamMin / Nlam
    delT = (dz / C0) * 0.95                                        # :49
    period = 1 / freq_in
    courantNo = (C0 * delT) / dz
    if courantNo > 3 or courantNo < 0:
        raise ValueError("courantNo is unstable")
    pmlWidth = 6 * int(lamMin / dz)                                # :62
    if pmlWidth >= 12000:
        raise ValueError("pmlWidth too big")
    Nz = int(domainSize / dz) + 2 * pmlWidth                       # :73
    if minim == maxim:
        minim -= 1
    N = None
    for N in range(minim, maxim):                                  # :84-95
        check = (freq_in * N) / (1 / delT)
        if int(check) - check == 0:
            break
    timeSteps = N if (N is not None and N >= minim) else minim     # :113-116
    timeSteps += int(timeSteps * (Nlam / 200))                     # :118
    if timeSteps >= 2 ** 15:
        raise ValueError("timeSteps too large")
    nzsrcFromPml = int(0.05 / dz)
    if nzsrcFromPml >= Nz * 0.65:
        raise ValueError("src is too far into domain")
    nzsrc = nzsrcFromPml + pmlWidth
    if nzsrc - 10 <= pmlWidth:
        raise ValueError("The probe for fft is in the PML region")
    matDist = int(0.1 / dz)
    mf = matDist + pmlWidth
    mr = Nz - 1
    if mr - mf < 10:
        raise ValueError("width is too small or negative")
    if matDist >= domainSize / dz:
        raise ValueError("Material starts in CPML region")
    if mf <= nzsrc:
        raise ValueError("Source is inside material")
    return dict(Nz=Nz, timeSteps=timeSteps, x1Loc=mf - 20, x2Loc=nzsrc - 100, mf=mf, mr=mr,
                pmlWidth=pmlWidth, nzsrc=nzsrc, lamMin=lamMin, dz=dz, delT=delT,
                courantNo=f32(courantNo), period=period, Nlam=Nlam)


def lorentz_eps(wp, w0, gam, freq):
    """BaseFDTD11.py:804-807 / Environment_Setup.py:24-27."""
    epsNum = wp * wp
    epsDom = (w0 * w0 - (2 * np.pi * freq * 2 * np.pi * freq) + 1j * gam * 2 * np.pi * freq)
    return 1 + epsNum / epsDom


def default_medium():
    """Variables.__init__, MasterController.py:177-179,196,202."""
    return dict(wp=float(np.sqrt(1.5 * (2 * np.pi * 20e9) ** 2)), gam=2 * np.pi * 20e9 * 0.1,
                w0=2 * np.pi * 20e9, alpha3=f32(0.7), chi3=1e-3)


def spatial_stab(Nz, dz, freq, dt, wp, w0, gam):
    """genericStability.py:12-62; returns the adjusted plasma frequency (and fix)."""
    rad = 2 * np.pi * freq
    twoPi = rad / freq
    c0 = sci.speed_of_light
    if rad * dt > np.pi / 2 and freq <= 5e9:
        raise ValueError("unstable timestep")
    sw = np.sinc(np.pi * freq * dt)
    # Drude limit w0 = 0 (not reachable in the reference, which divides by w0^2 here): es*w0^2 -> wp^2
    es_w02 = wp ** 2 if w0 == 0 else ((wp ** 2) / (w0 ** 2) - 1) * w0 ** 2
    sqN = (rad ** 2 * sw ** 2 - es_w02 * np.cos(rad * dt) + 1.0j * gam * rad * sw)
    sqD = (rad ** 2 * sw ** 2 - w0 ** 2 * np.cos(rad * dt) + 1.0j * gam * rad * sw)
    arg = (rad / c0) * (dz / 2) * sw * np.sqrt(sqN / sqD)
    ans = (2 / dz) * np.arcsin(arg)
    kNum = abs(ans)
    epsilon = 1 + ((wp ** 2) / (w0 ** 2 - (rad ** 2) - 1j * gam * rad))
    refr = np.sqrt(abs(np.real(epsilon)))
    fix = (c0 * dt * np.sin((kNum * refr * dz) / 2)) / (refr * dz * np.sin((kNum * c0 * dt) / 2))
    pf = np.sqrt(abs(fix)) * wp
    lamDisc = twoPi / kNum
    if kNum * dz > np.pi / 2:
        raise ValueError("unstable, wave is not resolved")
    if kNum * dz * Nz < 5 * lamDisc:
        raise ValueError("unstable because domain too small")
    return float(pf), float(fix)


# ----------------------------------------------------------------------------- setup: coefficients
def cpml_coefficients(L, pw, dz, dt, UpExMat, UpExHcompsCo, *, kappaMax=1.0, r_scale=4.0,
                      r_a_scale=1.0, alphaMax=0.05, cpml_m=True, cpml_p=True):
    """BaseFDTD11.py:195-357 via Solver_Engine.boundCondManager (:70-87)."""
    kappaMax, r_scale, r_a_scale, alphaMax = f32(kappaMax), f32(r_scale), f32(r_a_scale), f32(alphaMax)
    sigmaOpt = 0.5 * (0.8 * 1 / (dz * math.pow(MU0 / EPS0, 0.5)))          # MasterController.py:356-358
    kap = np.ones(L); sig = np.zeros(L); alp = np.zeros(L)
    kn = np.empty(pw); sn = np.empty(pw); an = np.empty(pw)
    for n in range(pw):                                                     # :237-240
        r = (pw - n) / pw
        kn[n] = 1 + (kappaMax - 1) * math.pow(r, r_scale)
        sn[n] = sigmaOpt * math.pow(r, r_scale)
        an[n] = alphaMax * math.pow((n + 1) / pw, r_a_scale)
    kap[:pw] = kn; kap[L - pw:] = kn[::-1]                                  # :257-264
    sig[:pw] = sn; sig[L - pw:] = sn[::-1]
    alp[:pw] = an; alp[L - pw:] = an[::-1]
    be = np.zeros(L); ce = np.zeros(L); cm = np.zeros(L)
    for nz in list(range(pw)) + list(range(L - pw, L)):                     # :278-294
        b = math.exp(-((sig[nz] * dt / (kap[nz] * EPS0)) + ((alp[nz] * dt) / EPS0)))
        be[nz] = b
        den = sig[nz] * kap[nz] + alp[nz] * kap[nz] * kap[nz]
        ce[nz] = (b - 1) * sig[nz] / den
        cm[nz] = (b - 1) * sig[nz] / (den * dz)
    bm = be.copy()
    Cb = np.zeros(L); C2 = np.zeros(L)
    for nz in list(range(pw)) + list(range(L - 1, L - pw, -1)):             # :306-330
        Cb[nz] = UpExHcompsCo[nz] * UpExMat[nz]
        C2[nz] = dt / MU0
    denE = np.ones(L); denH = np.ones(L)                                    # :332-357
    jj = pw
    for j in range(L):
        if j <= pw and cpml_m:
            denH[j] = 1 / kap[j]
        elif j >= L - pw and cpml_p:
            denH[j] = 1 / kap[jj]; jj -= 1
    jj = pw - 1
    for j in range(L):
        if j <= pw and cpml_m:
            denE[j] = 1 / kap[j]
        elif j >= L - pw and cpml_p:
            denE[j] = 1 / kap[jj]; jj -= 1
    return dict(beX=be, ceX=ce, bmY=bm, cmY=cm, Cb=Cb, C2=C2, denE=denE, denH=denH,
                sigma=sig, kappa=kap, alpha=alp, sigmaOpt=sigmaOpt)


def smooth_turn_on(T, dt, dz, courantNo, period, Periods, frq):
    """BaseFDTD11.py:104-120."""
    ppw = C0 / (frq * dz)
    Exs = np.zeros(T); Hys = np.zeros(T)
    for t in range(T):
        if t * dt < period * Periods:
            Exs[t] = float(np.sin(2.0 * np.pi / ppw * (courantNo * t)))
            Hys[t] = float(np.sin(2.0 * np.pi / ppw * (courantNo * (t + 1))))
    return Exs, Hys


def gaussian_src(T, freq, amplitude, courantNo):
    """BaseFDTD11.py:86-96."""
    t = np.arange(T)
    fc = 200
    tau = fc * 2.2
    arg = ((t - tau) * (t - tau)) / (fc * fc) * np.cos(2 * np.pi * freq * (t - tau))
    return amplitude * (np.exp(-arg) * 2) / courantNo


def sig_mod(sig, dt, AmpMod, tau):
    """Solver_Engine.py:126-140."""
    a = (2 * np.pi * (1 / tau)) * (np.arange(len(sig)) * dt)
    a = np.where(a > 50, 50, a)
    return (AmpMod + sig) * (1 / np.cosh(a))


def sources(c):
    """Solver_Engine.SourceManager :89-124 (+ Sig_Mod for the FreeSpace integrator :157-160)."""
    T = c.T
    if c.source == "sine":
        Exs, Hys = smooth_turn_on(T, c.dt, c.dz, c.courantNo, c.period, c.Periods, c.freq)
        Exp = np.zeros(T); Hyp = np.zeros(T)
        if c.mode == "nl":                                           # P.nonLinMed pump, :94-101
            Exp, Hyp = smooth_turn_on(T, c.dt, c.dz, c.courantNo, c.period, c.Periods, c.freq * 0.8)
            Exp = Exp * c.courantNo; Hyp = Hyp * c.courantNo
            Exp = Exp * 0.1; Hyp = Hyp * 0.01
        Exs = Exs * c.courantNo + Exp
        Hys = Hys * c.courantNo + Hyp
        if c.tfsf:
            Hys = Hys * (1 / CHAR_IMP)
    elif c.source == "gauss":
        Exs = gaussian_src(T, c.freq, c.Amplitude, c.courantNo)
        Hys = gaussian_src(T, c.freq, c.Amplitude, c.courantNo) if c.tfsf else np.zeros(T)
    else:
        raise ValueError(c.source)
    if c.mode == "free":
        tauIn = 1 / (c.freq / 5)
        Exs = sig_mod(Exs, c.dt, 1, tauIn)
        Hys = sig_mod(Hys, c.dt, 1 / CHAR_IMP, tauIn)
    return Exs, Hys


# ----------------------------------------------------------------------------- case container
@dataclass
class Case:
    mode: str                 # "free" | "lorentz" | "nl" | "lorentz_nl" (builder-defined composition)
    freq: float
    Nz: int
    T: int
    pw: int
    mf: int
    mr: int
    nzsrc: int
    x1Loc: int
    x2Loc: int
    dz: float
    dt: float
    courantNo: float          # float32-rounded
    period: float
    source: str = "sine"
    tfsf: bool = True
    Periods: float = 1.0
    Amplitude: float = 1.0
    epsRe: float = 1.0
    muRe: float = 1.0
    vidInterval: int = 50
    medium: dict = field(default_factory=default_medium)
    cpml: dict = field(default_factory=dict)   # optional overrides kappaMax / alphaMax / ...

    @property
    def L(self):
        return self.Nz + 1


def make_case(mode, freq, domainSize, minim, maxim, *, source="sine", tfsf=True, periods=1000.0,
              epsRe=1.0, amplitude=1.0, eps_for_nlam=None, **kw) -> Case:
    e = env_setup(freq, domainSize, minim, maxim, eps_for_nlam=eps_for_nlam, nonLinMed=(mode == "nl"))
    return Case(mode=mode, freq=freq, Nz=e["Nz"], T=e["timeSteps"], pw=e["pmlWidth"], mf=e["mf"],
                mr=e["mr"], nzsrc=e["nzsrc"], x1Loc=e["x1Loc"], x2Loc=e["x2Loc"], dz=e["dz"],
                dt=e["delT"], courantNo=e["courantNo"], period=e["period"], source=source,
                tfsf=tfsf, Periods=f32(periods), Amplitude=f32(amplitude), epsRe=f32(epsRe), **kw)


# ----------------------------------------------------------------------------- C stepping library
_d = ctypes.POINTER(ctypes.c_double)


class OrcGrid(ctypes.Structure):
    _fields_ = [
        ("L", ctypes.c_int), ("pw", ctypes.c_int), ("mf", ctypes.c_int), ("mr", ctypes.c_int),
        ("nzsrc", ctypes.c_int), ("tfsf", ctypes.c_int), ("cpml_m", ctypes.c_int), ("cpml_p", ctypes.c_int),
        ("dt_over_dz", ctypes.c_double), ("eps0", ctypes.c_double),
        ("polA", ctypes.c_double), ("polB", ctypes.c_double), ("polC", ctypes.c_double),
        ("cub_a", ctypes.c_double), ("cub_b", ctypes.c_double), ("cub_c", ctypes.c_double),
        ("nl_den0", ctypes.c_double), ("nl_den1", ctypes.c_double),
    ] + [(n, _d) for n in ("Ex", "Hy", "Dx", "P", "Pprev", "psiE", "psiH", "Acubic",
                           "Jx", "UpExMat", "denE", "UpHySelf", "UpHyMat", "denH",
                           "beX", "ceX", "Cb", "bmY", "cmY", "C2", "srcE", "srcH")] + [
        ("n_probes", ctypes.c_int), ("probe_idx", ctypes.POINTER(ctypes.c_int)), ("probe_out", _d),
        ("snap_interval", ctypes.c_int), ("snap_rows", ctypes.c_int), ("snap_out", _d),
    ]


_lib = None


def build_lib(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "fdtd_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "_build/liboracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build_lib()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.orc_run.argtypes = [ctypes.POINTER(OrcGrid)] + [ctypes.c_int] * 5
        _lib.orc_run.restype = ctypes.c_int
        _lib.orc_run_batch.argtypes = [ctypes.POINTER(OrcGrid), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        _lib.orc_cubic_root0.argtypes = [ctypes.c_double] * 4
        _lib.orc_cubic_root0.restype = ctypes.c_double
        _lib.orc_sizeof_grid.restype = ctypes.c_size_t
        assert _lib.orc_sizeof_grid() == ctypes.sizeof(OrcGrid)
    return _lib


MODE_ID = {"free": 0, "lorentz": 1, "nl": 2, "lorentz_nl": 3}
KERR_EPS_INF = 1.0   # Kerr-Lorentz composition (mode "lorentz_nl", see fdtd_oracle.c): Dx - P = eps0 (eps_inf + chi3 |E|^2) E


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def lorentz_abc(dt, wp, w0, gam):
    """BaseFDTD11.py:620-626."""
    D = (1 / dt ** 2) + (gam / (2 * dt))
    A = ((2 / dt ** 2) - w0 ** 2) / D
    B = ((gam / (2 * dt)) - 1 / dt ** 2) / D
    C = (EPS0 * wp ** 2) / D
    return A, B, C


def cubic_abc(freq, wp, w0, gam, alpha3, chi3):
    """BaseFDTD11.py:804-810."""
    eps = lorentz_eps(wp, w0, gam, freq)
    cub = (alpha3 * chi3) ** 2
    qua = 2 * np.real(alpha3 * eps * chi3)
    one = np.abs(eps) ** 2
    return float(cub), float(qua), float(one)


class PassArrays:
    """All arrays of one pass (fresh state), kept alive for the C struct."""

    def __init__(self, c: Case, wp: float, Exs, Hys, probes, snapshots: bool, Jx=None):
        L = c.L
        z = lambda: np.zeros(L)
        self.Ex, self.Hy, self.Dx, self.P, self.Pprev = z(), z(), z(), z(), z()
        self.psiE, self.psiH, self.Acubic = z(), z(), z()
        self.Jx = z() if Jx is None else np.ascontiguousarray(Jx, dtype=np.float64)
        # EmptySpaceCalc, BaseFDTD11.py:124-132
        self.UpHyMat = np.ones(L) * ((1 / CHAR_IMP) * c.courantNo)
        self.UpExMat = np.ones(L) * (CHAR_IMP * c.courantNo)
        self.UpHySelf = np.ones(L)
        UpExHcompsCo = np.ones(L)
        if c.mode == "free":            # Material + UpdateCoef, BaseFDTD11.py:136-182
            for k in range(c.mf, c.mr):
                self.UpExMat[k] = self.UpExMat[k] / c.epsRe
                self.UpHyMat[k] = self.UpHyMat[k] / c.muRe
        k = cpml_coefficients(L, c.pw, c.dz, c.dt, self.UpExMat, UpExHcompsCo, **c.cpml)
        self.coef = k
        self.srcE = np.ascontiguousarray(Exs / c.courantNo)
        self.srcH = np.ascontiguousarray(Hys / c.courantNo)
        self.probe_idx = np.asarray(probes, dtype=np.int32)
        self.probe_out = np.zeros((len(probes), c.T))
        rows = int(c.T / c.vidInterval)
        self.snap = np.zeros((rows, L)) if snapshots else None
        g = OrcGrid()
        g.L, g.pw, g.mf, g.mr, g.nzsrc = L, c.pw, c.mf, c.mr, c.nzsrc
        g.tfsf, g.cpml_m, g.cpml_p = int(c.tfsf), 1, 1
        g.dt_over_dz = c.dt / c.dz
        g.eps0 = EPS0
        m = c.medium
        g.polA, g.polB, g.polC = lorentz_abc(c.dt, wp, m["w0"], m["gam"])
        g.cub_a, g.cub_b, g.cub_c = cubic_abc(c.freq, wp, m["w0"], m["gam"], m["alpha3"], m["chi3"])
        g.nl_den0 = EPS0 * float(np.sqrt(1.2))
        g.nl_den1 = EPS0 * m["chi3"]
        if c.mode == "lorentz_nl":
            chi3 = float(m["chi3"])
            g.cub_a, g.cub_b, g.cub_c = chi3 ** 2, 2 * KERR_EPS_INF * chi3, KERR_EPS_INF ** 2
            g.nl_den0 = EPS0 * KERR_EPS_INF
        for n in ("Ex", "Hy", "Dx", "P", "Pprev", "psiE", "psiH", "Acubic", "Jx", "UpExMat",
                  "UpHySelf", "UpHyMat", "srcE", "srcH"):
            setattr(g, n, _dp(getattr(self, n)))
        for n in ("denE", "denH", "beX", "ceX", "Cb", "bmY", "cmY", "C2"):
            setattr(g, n, _dp(k[n]))
        g.n_probes = len(probes)
        g.probe_idx = self.probe_idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
        g.probe_out = _dp(self.probe_out)
        g.snap_interval = c.vidInterval
        g.snap_rows = rows if snapshots else 0
        g.snap_out = _dp(self.snap) if snapshots else None
        self.g = g


def run_case(c: Case, snapshots=False, nsteps=None):
    """Controller (MasterController.py:451-469) -> Integrator*1D (Solver_Engine.py:142-371)."""
    T = c.T if nsteps is None else nsteps
    fin_be = int(c.T * 0.7)
    start_af = int(c.T * 0.05)
    wp = c.medium["wp"]
    out = {}
    x1ColBe = np.zeros(c.T); x1ColAf = np.zeros(c.T)
    if c.mode in ("free", "lorentz", "lorentz_nl"):
        for i in range(2):
            if c.mode != "free":         # Solver_Engine.py:286 -- cumulative over the two passes
                wp, _ = spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp, c.medium["w0"], c.medium["gam"])
            Exs, Hys = sources(c)
            pa = PassArrays(c, wp, Exs, Hys, [c.x1Loc if i == 0 else c.x2Loc], snapshots and i == 1)
            lib().orc_run(ctypes.byref(pa.g), MODE_ID[c.mode], int(i == 1), 0, T, c.T)
            n = np.arange(c.T)
            if i == 0:
                x1ColBe = np.where(n <= fin_be, pa.probe_out[0], 0.0)      # :360-363
            else:
                x1ColAf = np.where(n >= start_af, pa.probe_out[0], 0.0)    # :365-368
        out.update(P=pa.P, Pprev=pa.Pprev, Dx=pa.Dx, Acubic=pa.Acubic)
    else:
        wp, _ = spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp, c.medium["w0"], c.medium["gam"])  # :231
        Exs, Hys = sources(c)
        pa = PassArrays(c, wp, Exs, Hys, [c.mf, c.mr], snapshots)
        lib().orc_run(ctypes.byref(pa.g), MODE_ID["nl"], 0, 0, T, c.T)
        out.update(Port1=pa.probe_out[0], Port2=pa.probe_out[1], Acubic=pa.Acubic, Dx=pa.Dx)
    out.update(Ex=pa.Ex, Hy=pa.Hy, psi_Ex=pa.psiE, psi_Hy=pa.psiH, Exs=Exs, Hys=Hys,
               x1ColBe=x1ColBe, x1ColAf=x1ColAf, plasmaFreqE=wp, coef=pa.coef,
               UpExMat=pa.UpExMat, UpHyMat=pa.UpHyMat, Ex_History=pa.snap)
    return out


def ref_tester(y, T, dt):
    """TransformHandler.RefTester :34-78 -> peak of 2|FFT|/timeSteps (DC peak is an error)."""
    from scipy import fftpack
    Y = fftpack.fft(y)
    Ypow = (2 * np.abs(Y)) / T
    ind = int(np.argmax(Ypow))
    if ind == 0:
        raise ValueError("Could not find non-DC freq")
    return float(Ypow[ind]), float(fftpack.fftfreq(len(y), d=dt)[ind])


def reflection(out, c: Case):
    """MasterController.results(RefCo=True) :485-500."""
    return ref_tester(out["x1ColAf"], c.T, c.dt)[0] / ref_tester(out["x1ColBe"], c.T, c.dt)[0]


def analytical_reflection(freq, wp, w0, gam):
    """BaseFDTD11.AnalyticalReflectionE :882-923 (returned value only)."""
    refr2 = np.real(np.sqrt(lorentz_eps(wp, w0, gam, freq)))
    return float(abs((refr2 - 1) / (1 + refr2)))


# ----------------------------------------------------------------------------- dormant models
def _cd(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def varin_pbar(mf, mr, eps0, chi1, chi3, alpha3, Ex, Qx3, Pbar3):
    """BaseFDTD11.py:567-577 ADE_NonLin_Pol_Ex_Pbar (in place on Pbar3)."""
    f = lib().orc_varin_pbar
    f.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.POINTER(ctypes.c_double)] * 3
    f.restype = None
    f(int(mf), int(mr), eps0, chi1, chi3, alpha3, _cd(Ex), _cd(Qx3), _cd(Pbar3))


def varin_lin(mf, mr, gammaE, omega0, dt, Jx, P, Pbar3):
    """BaseFDTD11.py:580-594 ADE_Lin_Curr_And_Pol_Varin (in place on Jx, P)."""
    f = lib().orc_varin_lin
    f.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 3 + [ctypes.POINTER(ctypes.c_double)] * 3
    f.restype = None
    f(int(mf), int(mr), gammaE, omega0, dt, _cd(Jx), _cd(P), _cd(Pbar3))


def varin_qg(mf, mr, gamma3, omega3, dt, Ex, Gx3, Qx3):
    """BaseFDTD11.py:596-609 ADE_Nonlin_Q_and_G (in place on Gx3, Qx3)."""
    f = lib().orc_varin_qg
    f.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 3 + [ctypes.POINTER(ctypes.c_double)] * 3
    f.restype = None
    f(int(mf), int(mr), gamma3, omega3, dt, _cd(Ex), _cd(Gx3), _cd(Qx3))


def kerr_nonlin(alpha3, eps0, chi3, dt, Ex, Eold, JxKerr):
    """BaseFDTD11.py:762-766 KerrNonlin (in place on JxKerr)."""
    f = lib().orc_kerr_nonlin
    f.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.POINTER(ctypes.c_double)] * 3
    f.restype = None
    f(len(Ex), alpha3, eps0, chi3, dt, _cd(Ex), _cd(Eold), _cd(JxKerr))


def mur1d(Nz, c0, dt, dz, Ex, Eold):
    """BaseFDTD11.py:769-788 MUR1DEx (in place on Ex)."""
    f = lib().orc_mur1d
    f.argtypes = [ctypes.c_int] + [ctypes.c_double] * 3 + [ctypes.POINTER(ctypes.c_double)] * 2
    f.restype = None
    f(int(Nz), c0, dt, dz, _cd(Ex), _cd(Eold))


def drude_j(k):
    """TESTBOXDIPSERSE.py:79-94 for the constants dict of pyfdtd_b200.drude_sandbox.constants(); returns (Ex, Hy, Jx)."""
    n = int(k["domain"])
    Ex, Hy, Jx, tE, tEo = (np.zeros(n) for _ in range(5))
    Hys = np.ascontiguousarray(k["Hys"], dtype=np.float64)
    f = lib().orc_drude_j
    f.argtypes = [ctypes.c_int] * 5 + [ctypes.c_double] * 5 + [ctypes.POINTER(ctypes.c_double)] * 6
    f.restype = None
    f(n, int(k["tim"]), int(k["src"]), int(k["matFront"]), int(k["matRear"]), k["cour"], k["kapE"], k["betaE"], k["perm0"],
      k["dt"], _cd(Hys), _cd(Ex), _cd(Hy), _cd(Jx), _cd(tE), _cd(tEo))
    return Ex, Hy, Jx
