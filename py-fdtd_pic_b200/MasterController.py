"""State / parameter objects and the driver entry points of the hot path.

Mirror of the reference's ``MasterController`` module for everything the time-stepping path uses:
``Params`` (:287), ``Variables`` (:145), ``CPML_Params`` (:351), ``CPML_Variables`` (:402),
``Controller`` (:451-469), ``results`` (:472-505) and ``LoopedSim`` (:530-606).  Call signatures,
attribute names and return tuples are the reference's; the time loops behind ``Controller`` run as
sm_100a CUDA kernels through the C-ABI (see Solver_Engine.py in this package).

The reference declares these classes as numba jitclasses whose spec quantises several members to
float32 / int32 on every assignment (duplicate spec keys: the last one wins).  The
plain-Python classes here reproduce that with a typed ``__setattr__`` so that e.g.
``P.courantNo == 0.949999988079071`` exactly as in the reference.

Out of scope (reference-only cosmetics): Reporter, plotter output, jinja report, __Main__ at import.
"""
from __future__ import annotations

import numpy as np
import scipy.constants as _sc

from . import BaseFDTD11, Environment_Setup as envDef, Solver_Engine as SE, TransformHandler as transH

_I32, _F32, _F64, _BOOL, _ARR, _ARR2, _CARR = "i32", "f32", "f64", "bool", "arr", "arr2", "carr"


class _TypedStruct:
    """Attribute container with jitclass-like member typing."""
    _types: dict = {}
    _lazy: dict = {}

    def __setattr__(self, name, value):
        t = self._types.get(name)
        if t == _F32:
            value = float(np.float32(value))
        elif t == _F64:
            value = float(value)
        elif t == _I32:
            value = int(np.int32(int(value)))
        elif t == _BOOL:
            value = bool(value)
        elif t == _ARR:
            value = np.asarray(value, dtype=np.float64)
            if value.ndim != 1:
                raise TypeError(f"{type(self).__name__}.{name} must be a 1-D float64 array")
        elif t == _ARR2:
            value = np.asarray(value, dtype=np.float64)
        elif t == _CARR:
            value = np.asarray(value, dtype=np.complex128)
        object.__setattr__(self, name, value)

    def __getattr__(self, name):
        # only reached when the attribute is not set yet: large, rarely used arrays are created on
        # first touch (the reference allocates six [T/interval, Nz+1] histories up front)
        lazy = type(self)._lazy
        if name in lazy:
            value = lazy[name](self)
            object.__setattr__(self, name, value)
            return value
        raise AttributeError(name)


class Variables(_TypedStruct):
    """Everything that changes during a simulation (MasterController.py:82-210)."""
    _types = dict(
        Nz=_I32, timeSteps=_F32, plasmaFreqE=_F64, gammaE=_F64, omega_0E=_F64, test=_I32, tempTempTest=_I32,
        attenAmnt=_I32, alpha3=_F32, nonLin3gammaE=_F64, nonLin3Omega_0E=_F64, chi1Stat=_F64, chi3Stat=_F64,
        cubPoly=_CARR, roots=_CARR,
        **{k: _ARR for k in (
            "UpHyMat UpExMat UpExHcompsCo UpExSelf UpHySelf UpHyEcompsCo Ex Hy Dx Jx x1ColBe x1ColAf x1Jx x1Hy "
            "x1ExOld x1JxOld x1HyOld epsilon mu polarisationCurr tempVarPol tempTempVarPol tempVarE tempTempVarE "
            "tempVarHy tempTempVarHy tempVarJx tempTempVarJx tempVarDx tempTempVarDx tempTest Gx3 Qx3 Pbar3 "
            "JxKerr JxRaman Acubic Port1 Port2").split()},
        **{k: _ARR2 for k in "Ex_History Hy_History Jx_History polCurr_History Dx_History Psi_e_History x1Atten".split()},
    )

    def __init__(self, Nz, timeSteps, interval, attenAmnt):
        L = Nz + 1
        object.__setattr__(self, "_L", L)
        object.__setattr__(self, "_rows", int(timeSteps / interval))
        object.__setattr__(self, "_T", int(timeSteps))
        for k in ("UpHyMat", "UpExMat", "UpExHcompsCo", "UpExSelf", "UpHySelf", "UpHyEcompsCo", "epsilon", "mu"):
            setattr(self, k, np.ones(L))
        for k in ("Ex", "Hy", "Dx", "Jx", "polarisationCurr", "tempVarPol", "tempTempVarPol", "tempVarE",
                  "tempTempVarE", "tempVarHy", "tempTempVarHy", "tempVarJx", "tempTempVarJx", "tempVarDx",
                  "tempTempVarDx", "tempTest", "Gx3", "Qx3", "Pbar3", "JxKerr", "JxRaman", "Acubic"):
            setattr(self, k, np.zeros(L))
        for k in ("x1ColBe", "x1ColAf", "x1Jx", "x1Hy", "x1ExOld", "x1JxOld", "x1HyOld", "Port1", "Port2"):
            setattr(self, k, np.zeros(int(timeSteps)))
        # Lorentz medium and Kerr constants, MasterController.py:177-179, 196-202
        self.plasmaFreqE = np.sqrt(((1.5) * (2 * np.pi * 20e9) ** 2))
        self.gammaE = 2 * np.pi * 20e9 * 0.1
        self.omega_0E = 2 * np.pi * 20e9
        self.test = 5
        self.tempTempTest = 0
        self.attenAmnt = attenAmnt
        self.x1Atten = np.zeros((attenAmnt, int(timeSteps)))
        self.alpha3 = 0.7
        self.nonLin3gammaE = 0
        self.nonLin3Omega_0E = 6e9
        self.chi1Stat = np.sqrt(1.2) - 1
        self.chi3Stat = 1e-3
        self.roots = np.zeros(4, dtype=np.complex128)

    def _build_ex_history(self):
        """Ex_History on first touch: zeros, plus the snapshot rows the integrators left ON THE DEVICE (Solver_Engine
        run_time_loop parks them there: a default run's history is 50 MB that most callers never read)."""
        h = np.zeros((self._rows, self._L))
        for t, lo, hi in self.__dict__.pop("_pending_history", []):
            h[lo:hi] = t[lo:hi].cpu().numpy()
        return h

    _lazy = {
        "Ex_History": lambda self: self._build_ex_history(),
        **{k: (lambda self: np.zeros((self._rows, self._L)))
           for k in "Hy_History Jx_History polCurr_History Dx_History Psi_e_History".split()},
        "cubPoly": lambda self: np.zeros((self._L, 4), dtype=np.complex128),
    }

    def __str__(self):
        return "Contains data that will change during sim"


class Params(_TypedStruct):
    """Values that stay constant during a simulation (MasterController.py:217-336)."""
    _types = dict(
        Nz=_I32, timeSteps=_I32, eLoss=_F64, mLoss=_F64, eSelfCo=_F64, eHcompsCo=_F64, hSelfCo=_F64, hEcompsCo=_F64,
        x1Loc=_I32, x2Loc=_I32, materialFrontEdge=_I32, materialRearEdge=_I32, pmlWidth=_I32, nzsrc=_I32,
        lamMin=_F64, dz=_F64, delT=_F64, courantNo=_F32, period=_F64, domainSize=_I32, freq_in=_F64,
        permit_0=_F64, permea_0=_F64, CharImp=_F64, c0=_F64, Nlam=_F64, MORmode=_BOOL, delayMOR=_I32,
        CPMLXp=_BOOL, CPMLXm=_BOOL, TFSF=_BOOL, SineCont=_BOOL, Gaussian=_BOOL, Ricker=_BOOL, Amplitude=_F32,
        Periods=_F32, LorentzMed=_BOOL, nonLinMed=_BOOL, FreeSpace=_BOOL, epsRe=_F32, muRe=_F32, vidMake=_BOOL,
        vidInterval=_I32, atten=_BOOL, julia=_BOOL, testMode=_BOOL,
    )

    def __init__(self, Nz, timeSteps, eLoss, mLoss, eSelfCo, eHcompsCo, hSelfCo, hEcompsCo, x1Loc, x2Loc,
                 materialFrontEdge, materialRearEdge, pmlWidth, nzsrc, lamMin, dz, delT, courantNo, period, Nlam,
                 MORmode, domainSize, freq_in, delayMOR, LorentzMed=False, nonLinMed=False, SineCont=False,
                 Gaussian=False, TFSF=False):
        self.permit_0 = _sc.epsilon_0
        self.permea_0 = _sc.mu_0
        self.CharImp = 376.730313668
        self.c0 = 299792458.0
        loc = locals()
        for k in ("freq_in", "lamMin", "Nlam", "dz", "delT", "courantNo", "materialFrontEdge", "materialRearEdge",
                  "Nz", "timeSteps", "x1Loc", "x2Loc", "nzsrc", "period", "eLoss", "eSelfCo", "eHcompsCo", "mLoss",
                  "hSelfCo", "hEcompsCo", "pmlWidth", "domainSize", "MORmode", "delayMOR"):
            setattr(self, k, loc[k])
        self.CPMLXp = True
        self.CPMLXm = True
        self.TFSF = TFSF
        self.SineCont = SineCont
        self.Gaussian = Gaussian
        self.Ricker = False
        self.Amplitude = 1.0
        self.Periods = 1.0
        self.LorentzMed = LorentzMed
        self.nonLinMed = nonLinMed
        self.FreeSpace = True
        self.epsRe = 1.0
        self.muRe = 1.0
        self.vidMake = True
        self.vidInterval = 50
        self.atten = False
        self.julia = False
        self.testMode = False

    def __str__(self):
        return "Class containing all values that remain constant throughout a sim"


class CPML_Params(_TypedStruct):
    """CPML grading constants (MasterController.py:341-360)."""
    _types = dict(kappaMax=_F32, r_scale=_F32, r_a_scale=_F32, sigmaEMax=_F64, sigmaHMax=_F64, sigmaOpt=_F64,
                  alphaMax=_F32)

    def __init__(self, dz):
        import math
        self.kappaMax = 1
        self.r_scale = 4
        self.r_a_scale = 1
        # the reference evaluates x**0.5 through libm pow inside a jitclass constructor
        self.sigmaEMax = 0.5 * (0.8 * (1) / (dz * math.pow(_sc.mu_0 / _sc.epsilon_0, 0.5)))
        self.sigmaHMax = self.sigmaEMax
        self.sigmaOpt = self.sigmaEMax
        self.alphaMax = 0.05

    def __str__(self):
        return "Class containing all CPML values that remain constant throughout a sim"


class CPML_Variables(_TypedStruct):
    """CPML arrays (MasterController.py:369-434); all length Nz+1 except the two probes."""
    _NAMES = ("kappa_Ex kappa_Hy psi_Ex psi_Hy alpha_Ex alpha_Hy sigma_Ex sigma_Hy beX bmY ceX cmY Ca Cb Cc C1 C2 C3 "
              "eLoss_CPML mLoss_CPML den_Hydz den_Exdz tempTempVarPsiEx tempVarPsiEx tempTempVarPsiHy tempVarPsiHy").split()
    _types = dict(Nz=_I32, **{k: _ARR for k in _NAMES},
                  psi_Ex_Probe=_ARR, psi_Hy_Probe=_ARR, psi_Ex_Old=_ARR, psi_Hy_Old=_ARR)

    def __init__(self, Nz, timeSteps):
        L = Nz + 1
        for k in self._NAMES:
            setattr(self, k, np.zeros(L))
        for k in ("psi_Ex_Probe", "psi_Hy_Probe", "psi_Ex_Old", "psi_Hy_Old"):
            setattr(self, k, np.zeros(int(timeSteps)))

    def __str__(self):
        return "Class containing all CPML values that vary throughout a sim"


# ------------------------------------------------------------------------------------------------
def Controller(V, P, C_V, C_P):
    """MasterController.py:451-469 -- dispatch to the integrator for the chosen medium."""
    probeReadFinishBe = int(P.timeSteps * 0.7)
    probeReadStartAf = int(P.timeSteps * 0.05)
    V.x1ColBe = np.zeros(P.timeSteps)
    V.x1ColAf = np.zeros(P.timeSteps)
    if P.LorentzMed:
        integ = SE.IntegratorLinLor1D
    elif P.FreeSpace:
        integ = SE.IntegratorFreeSpace1D
    elif P.nonLinMed:
        integ = SE.IntegratorNL1D
    else:
        raise ValueError("Controller: none of LorentzMed / FreeSpace / nonLinMed is set")
    V.Ex, V.Hy, Exs, Hys, C_V.psi_Ex, C_V.psi_Hy, V.x1ColBe, V.x1ColAf = integ(
        V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf)
    return V, P, C_V, C_P, Exs, Hys


def results(V, P, C_V, C_P, time_Vec, RefCo=False, FFT=False, AnalRefCo=False, attenRead=False):
    """MasterController.py:472-505 -- reflection / attenuation figures from the probe traces."""
    if attenRead:
        _, _, val = transH.RefTester(V, P, V.x1ColBe, 1)
        vals = np.array([transH.RefTester(V, P, row, 1)[2] for row in V.x1Atten])
        return vals / val
    if RefCo:
        _, _, val = transH.RefTester(V, P, V.x1ColBe, 1)
        _, _, val2 = transH.RefTester(V, P, V.x1ColAf, 1)
        return val2 / val
    if AnalRefCo:
        return BaseFDTD11.AnalyticalReflectionE(V, P)
    return "results ran to end"


class Reporter:
    """Minimal stand-in for the reference's Reporter jitclass (MasterController.py:56-76)."""

    def __init__(self):
        self.dict1 = {"": ""}

    def printer(self, item="", name="", show=False):
        self.dict1[name] = item
        if show:
            print(item)


def plotter(xAxisData, yAxisData1, yAxisData2, **kwargs):
    """Plotting is out of scope; the data are returned so callers can plot them."""
    return np.asarray(xAxisData), np.asarray(yAxisData1), np.asarray(yAxisData2)


def LoopedSim(Rep, V, P, C_V, C_P, MORmode, domainSize, lowLimTim, highLimTim,
              stringparamSweep="Input frequency sweep", loop=False, Low=3e9, Interval=1e8, RefCoBool=True,
              points=20, batched=True):
    """MasterController.py:530-606.

    loop=False: one Controller run (+ reflection figures for a Lorentz medium).
    loop=True : the reference's 20-point frequency sweep.  The members are independent, so with
    ``batched=True`` (default) all of them are advanced together by the fused tile engine
    (``pf_run_batch``); ``batched=False`` runs them one after another like the reference.  Either way
    the per-member setup chain is the reference's, including the quirk that member i's grid
    resolution is derived from member i-1's dispersion-corrected plasma frequency (:547).
    The sweep result is left in ``LoopedSim.last_sweep`` = (freqs, measured, analytical).
    """
    if loop:
        from . import sweep
        freqs, measured, analytical, last = sweep.frequency_sweep(
            V, P, domainSize, lowLimTim, highLimTim, Low=Low, Interval=Interval, points=points, batched=batched)
        LoopedSim.last_sweep = (freqs, measured, analytical)
        V, P, C_V, C_P, Exs, Hys = last
        return V, P, C_V, C_P, Exs, Hys
    V, P, C_V, C_P, Exs, Hys = Controller(V, P, C_V, C_P)
    t = np.arange(0, len(V.x1ColBe)) * P.delT
    if P.LorentzMed and not P.FreeSpace:
        LoopedSim.last_sweep = (np.array([P.freq_in]),
                                np.array([results(V, P, C_V, C_P, t, RefCo=True)]),
                                np.array([results(V, P, C_V, C_P, t, AnalRefCo=True)]))
    return V, P, C_V, C_P, Exs, Hys


LoopedSim.last_sweep = None
