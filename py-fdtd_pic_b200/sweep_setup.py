"""Setup of a whole sweep at once: the per-member host chain, vectorised over members.

The reference builds every sweep member one after another in Python: ``envSetup`` (Environment_Setup.py:19-166, with
its time-step bin search :84-118), ``Params`` / ``Variables`` / ``CPML_*`` objects with ~80 arrays of Nz+1 doubles,
``spatialStab`` (genericStability.py:12-62), the CPML coefficient builders (BaseFDTD11.py:222-357) and
``SourceManager`` (Solver_Engine.py:89-124) -- per pass.  The per-member mirror of that chain (MasterController /
Solver_Engine.prepare_pass in this package) costs ~14 ms per member, ten times the GPU time of the member's whole run.

Here the same VALUES are produced without the objects:

* ``envSetup_many``   -- envSetup over an array of frequencies (NumPy; the bin search is one [members, N] table).
* ``MemberTable``     -- structure of arrays: one row per member of everything a ``PfGrid`` descriptor and the native
                         input builder need for one pass (geometry, scalars of the update rules, CPML grading, source).
* ``lorentz_sweep_tables`` / ``nonlinear_sweep_table`` -- the tables of the two-pass Lorentz integrator
                         (IntegratorLinLor1D) and of the cubic integrator (IntegratorNL1D) for (frequency, amplitude) lists.
* the CPML profiles and source tables themselves are written by ``pf_host_sweep_inputs`` (csrc/pf_setup.cu) straight into
  the batch's pinned upload buffer, on all host threads (``sweep.MemberBatch.from_table``).

Every number is bit-identical to what the per-member chain produces (tests/test_sweep_setup.py); the tile engine's
canonical scalars (cE0 ... c2_pml) are by construction what ``_device.canonical_form`` extracts from the arrays the
per-member chain builds.  No device is needed for anything in this module.
"""
from __future__ import annotations

import numpy as np
import scipy.constants as _sc

from . import _native as nat
from . import genericStability as gStab

C0 = 299792458.0            # Environment_Setup.py:34, MasterController.py:294
CHAR_IMP = 376.730313668    # MasterController.py:293
ENV_FIELDS = ("Nz", "timeSteps", "eLoss", "mLoss", "eSelfCo", "eHcompsCo", "hSelfCo", "hEcompsCo", "x1Loc", "x2Loc",
              "materialFrontEdge", "materialRearEdge", "pmlWidth", "nzsrc", "lamMin", "dz", "delT", "courantNo", "period",
              "Nlam")


def _f32(x):
    """jitclass float32 members: the value a float32 field hands back, as float64."""
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def _libm_pow(x, e):
    """x ** e the way CPython / numba evaluate it for scalars: C library pow, one element at a time."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    nat.check(nat.lib().pf_host_pow(x.ctypes.data, float(e), y.ctypes.data, x.size), "pf_host_pow")
    return y


def _trunc_int(x):
    return np.trunc(x).astype(np.int64)


def pick_time_steps_many(freq, delT, minim, maxim):
    """Environment_Setup.py:77-116 for arrays: first N in [minim, maxim) with freq*N*delT on an integer, else maxim-1."""
    if minim == maxim:
        minim -= 1
    N = np.arange(minim, maxim, dtype=np.float64)
    out = np.empty(len(freq), dtype=np.int64)
    # [members, N] in slabs of 256 members (a 1024 x 1000 table of doubles would also be fine; this bounds memory for any size)
    for lo in range(0, len(freq), 256):
        f, dt = freq[lo:lo + 256, None], delT[lo:lo + 256, None]
        bp = (f * N[None, :]) / (1 / dt)
        hit = np.trunc(bp) - bp == 0
        first = np.argmax(hit, axis=1)
        out[lo:lo + 256] = np.where(hit.any(axis=1), first, len(N) - 1) + minim
    return out


def envSetup_many(freqs, domainSize, minim=400, maxim=600, *, nonLinMed=False, LorMed=False, Nlam=None):
    """``Environment_Setup.envSetup`` (Environment_Setup.py:19-166) for an array of frequencies.

    Returns a dict of arrays keyed like the reference's 20-tuple.  ``Nlam``: points per wavelength per member; None = the
    reference's VExists=False rule (350 for a nonlinear medium, else 400).  The guards raise ValueError with the
    reference's messages, naming the first offending member."""
    f = np.ascontiguousarray(freqs, dtype=np.float64)
    n = len(f)
    if Nlam is None:
        Nlam = np.full(n, 350 if nonLinMed else 400, dtype=np.int64)
    Nlam = np.asarray(Nlam)
    lamMin = C0 / f
    dz = lamMin / Nlam
    delT = (dz / C0) * 0.95
    period = 1 / f
    courantNo = (C0 * delT) / dz

    def guard(bad, msg):
        if np.any(bad):
            i = int(np.flatnonzero(bad)[0])
            raise ValueError(msg(i))
    guard((courantNo > 3) | (courantNo < 0), lambda i: f"{courantNo[i]} courantNo is unstable")
    pw = 6 * _trunc_int(lamMin / dz)
    guard(pw >= 12000, lambda i: f"pmlWidth too big {pw[i]}")
    Nz = _trunc_int(domainSize / dz) + 2 * pw
    T = pick_time_steps_many(f, delT, int(minim), int(maxim))
    T = T + _trunc_int(T * (Nlam / 200))
    guard(T >= 2 ** 15, lambda i: "timeSteps too large")
    srcOffset = _trunc_int(0.05 / dz)
    guard(srcOffset >= Nz * 0.65, lambda i: f"{srcOffset[i]} src is too far into domain")
    nzsrc = srcOffset + pw
    guard(nzsrc - 10 <= pw, lambda i: "The probe for fft is in the PML region")
    slabOffset = _trunc_int(0.1 / dz)
    mf = slabOffset + pw
    mr = Nz - 1
    guard(mr - mf < 10, lambda i: f"{mr[i] - mf[i]} width is too small or negative")
    guard(slabOffset >= domainSize / dz, lambda i: "Material starts in CPML region")
    guard(mf <= nzsrc, lambda i: "Source is inside material")
    one, zero = np.ones(n), np.zeros(n)
    return dict(Nz=Nz, timeSteps=T, eLoss=zero.astype(np.int64), mLoss=zero.astype(np.int64), eSelfCo=one, eHcompsCo=one,
                hSelfCo=one, hEcompsCo=one, x1Loc=mf - 20, x2Loc=nzsrc - 100, materialFrontEdge=mf, materialRearEdge=mr,
                pmlWidth=pw, nzsrc=nzsrc, lamMin=lamMin, dz=dz, delT=delT, courantNo=courantNo, period=period, Nlam=Nlam)


# ------------------------------------------------------------------------------------------------ medium
def default_medium(n):
    """MasterController.Variables.__init__ :177-179, 196-202."""
    w0 = 2 * np.pi * 20e9
    return dict(wp=np.full(n, float(np.sqrt(((1.5) * (2 * np.pi * 20e9) ** 2)))), gam=np.full(n, 2 * np.pi * 20e9 * 0.1),
                w0=np.full(n, w0), alpha3=float(np.float32(0.7)), chi3Stat=1e-3)


def corrected_plasma_freq(env, freq, wp, w0, gam, times):
    """V.plasmaFreqE after ``times`` applications of spatialStab (Solver_Engine.py:231,286: once per pass, cumulative).
    spatialStab is evaluated per member through the reference-exact scalar mirror (complex sqrt / arcsin)."""
    out = np.array(wp, dtype=np.float64)
    T, Nz, dz, dt = env["timeSteps"], env["Nz"], env["dz"], env["delT"]
    for i in range(len(out)):
        v = float(out[i])
        for _ in range(times):
            v = float(gStab.spatialStab(int(T[i]), int(Nz[i]), float(dz[i]), float(freq[i]), float(dt[i]), v,
                                        float(w0[i]), float(gam[i]))[3])
        out[i] = v
    return out


def lorentz_abc_many(delT, wp, w0, gam, eps0):
    """BaseFDTD11._lorentz_abc (BaseFDTD11.py:620-626) over members (Python-float `**` = libm pow)."""
    dt2 = _libm_pow(delT, 2)
    D = (1 / dt2) + (gam / (2 * delT))
    A = ((2 / dt2) - _libm_pow(w0, 2)) / D
    B = ((gam / (2 * delT)) - 1 / dt2) / D
    C = (eps0 * _libm_pow(wp, 2)) / D
    return A, B, C


def cubic_abc_many(freq, wp, w0, gam, alpha3, chi3):
    """BaseFDTD11._cubic_abc (BaseFDTD11.py:804-810), per member with the reference's scalar complex arithmetic."""
    n = len(freq)
    cub, qua, one = np.empty(n), np.empty(n), np.empty(n)
    for i in range(n):
        w = 2 * np.pi * float(freq[i])
        p, o, g = float(wp[i]), float(w0[i]), float(gam[i])
        eps = 1 + (p * p) / (o * o - (w * w) + 1j * g * w)
        cub[i] = float((alpha3 * chi3) ** 2)
        qua[i] = float(2 * np.real(alpha3 * eps * chi3))
        one[i] = float(np.abs(eps) ** 2)
    return cub, qua, one


# ------------------------------------------------------------------------------------------------ table
class MemberTable:
    """One row per sweep member for ONE pass: everything the PfGrid descriptor and pf_host_sweep_inputs need."""

    INT_FIELDS = ("L", "T", "nsteps", "pw", "mf", "mr", "nzsrc", "flags", "share")
    SCALARS = ("dt_over_dz", "eps0", "polA", "polB", "polC", "cub_a", "cub_b", "cub_c", "nl_den0", "nl_den1", "cE0", "cE1",
               "cH0", "cH1", "c2_pml")
    SETUP = ("dz", "delT", "kappaMax", "r_scale", "r_a_scale", "sigmaOpt", "alphaMax", "c0", "freq", "courantNo", "period",
             "periods", "charImp", "amp")

    def __init__(self, n):
        self.n = int(n)
        for k in self.INT_FIELDS:
            setattr(self, k, np.zeros(n, dtype=np.int64))
        for k in self.SCALARS + self.SETUP:
            setattr(self, k, np.zeros(n))
        self.share = np.arange(n, dtype=np.int64)   # member whose CPML profiles this member uses (itself = owns them)
        self.probes = np.zeros((n, 1), dtype=np.int64)   # [n, n_probes] probe cells
        self.tfsf = np.ones(n, dtype=np.int64)
        self.pump = np.zeros(n, dtype=np.int64)
        self.src_kind = np.ones(n, dtype=np.int64)   # 1: SmoothTurnOn sine built natively; 0: caller-filled tables
        self.env = None

    def select(self, idx):
        """Rows idx (e.g. the members one rank owns); ``share`` is re-based, a member whose profile owner is not in the
        selection owns its profiles."""
        idx = np.asarray(idx, dtype=np.int64)
        t = MemberTable(len(idx))
        for k in self.INT_FIELDS + self.SCALARS + self.SETUP + ("tfsf", "pump", "src_kind"):
            setattr(t, k, getattr(self, k)[idx].copy())
        t.probes = self.probes[idx].copy()
        pos = {int(g): j for j, g in enumerate(idx)}
        t.share = np.array([pos.get(int(self.share[g]), j) for j, g in enumerate(idx)], dtype=np.int64)
        first = {}
        for j in range(len(idx)):          # an owner outside the selection: the first selected member of its group owns
            key = int(self.share[idx[j]])
            if key not in pos:
                t.share[j] = first.setdefault(key, j)
        if self.env is not None:
            t.env = {k: v[idx] for k, v in self.env.items()}
        return t


def _cpml_params(dz):
    """MasterController.CPML_Params (:341-360): kappaMax, r_scale, r_a_scale, alphaMax are float32 members;
    sigmaEMax = sigmaOpt = 0.5*(0.8/(dz*sqrt(mu0/eps0))) with libm pow(x, 0.5)."""
    root = _libm_pow(np.full(1, _sc.mu_0 / _sc.epsilon_0), 0.5)[0]
    sigma = 0.5 * (0.8 * (1) / (dz * root))
    return dict(kappaMax=float(np.float32(1)), r_scale=float(np.float32(4)), r_a_scale=float(np.float32(1)),
                sigmaOpt=sigma, alphaMax=float(np.float32(0.05)))


def _base_table(env, freq, amp, *, periods, tfsf, pump, probes, flags_extra=0, nsteps=None, amplitude=1.0):
    n = len(freq)
    t = MemberTable(n)
    t.env = env
    cN = _f32(env["courantNo"])
    t.L[:] = env["Nz"] + 1
    t.T[:] = env["timeSteps"]
    t.nsteps[:] = env["timeSteps"] if nsteps is None else nsteps
    t.pw[:], t.mf[:], t.mr[:], t.nzsrc[:] = env["pmlWidth"], env["materialFrontEdge"], env["materialRearEdge"], env["nzsrc"]
    t.probes = np.asarray(probes, dtype=np.int64).reshape(n, -1)
    t.flags[:] = ((nat.PF_F_TFSF if tfsf else 0) | nat.PF_F_CPML_M | nat.PF_F_CPML_P | nat.PF_F_CANONICAL | flags_extra)
    t.tfsf[:] = 1 if tfsf else 0
    t.pump[:] = 1 if pump else 0
    t.dz[:], t.delT[:] = env["dz"], env["delT"]
    cp = _cpml_params(env["dz"])
    for k in ("kappaMax", "r_scale", "r_a_scale", "alphaMax"):
        getattr(t, k)[:] = cp[k]
    t.sigmaOpt[:] = cp["sigmaOpt"]
    t.c0[:], t.freq[:], t.courantNo[:], t.period[:] = C0, freq, cN, env["period"]
    t.periods[:] = float(np.float32(periods))
    t.charImp[:] = CHAR_IMP
    t.amp[:] = amp
    # scalars every mode shares
    eps0 = _sc.epsilon_0
    t.dt_over_dz[:] = env["delT"] / env["dz"]
    t.eps0[:] = eps0
    t.cE0[:] = CHAR_IMP * cN              # EmptySpaceCalc (BaseFDTD11.py:124-132)
    t.cE1[:] = t.cE0
    t.cH0[:] = (1 / CHAR_IMP) * cN
    t.cH1[:] = t.cH0
    t.c2_pml[:] = env["delT"] / _sc.mu_0  # CPML_Hy_Update_Coef (BaseFDTD11.py:321-330)
    return t


def _arith_flags(fma, fp32, newton):
    return (nat.PF_F_FMA if fma else 0) | (nat.PF_F_FP32 if fp32 else 0) | (nat.PF_F_NEWTON if newton else 0)


def lorentz_sweep_tables(freqs, amps, domainSize, lowLimTim, highLimTim, *, periods=1000, tfsf=True, nsteps=None,
                         fma=False, fp32=False, kerr_lorentz=False, env=None):
    """Tables of the two passes of ``IntegratorLinLor1D`` (Solver_Engine.py:275-371) for members (freqs[i], amps[i]):
    pass 0 = incident run probed at x1Loc, pass 1 = run with the polarisation update probed at x2Loc.  The grids come
    from ``envSetup(f, domainSize, lowLimTim, highLimTim, LorMed=True)``; the medium is the reference default
    (MasterController.py:177-179) with the dispersion correction applied once per pass, cumulatively."""
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    amps = np.broadcast_to(np.asarray(amps, dtype=np.float64), freqs.shape).copy()
    if env is None:
        env = envSetup_many(freqs, domainSize, lowLimTim, highLimTim, LorMed=True)
    n = len(freqs)
    med = default_medium(n)
    # one correction chain per DISTINCT grid (members of one frequency share it)
    uniq, first, inv = np.unique(freqs, return_index=True, return_inverse=True)
    envu = {k: v[first] for k, v in env.items()}
    wp1u = corrected_plasma_freq(envu, uniq, med["wp"][first], med["w0"][first], med["gam"][first], 1)
    wp2u = corrected_plasma_freq(envu, uniq, wp1u, med["w0"][first], med["gam"][first], 1)
    tables = []
    for pass_idx, wpu in enumerate((wp1u, wp2u)):
        probes = env["x1Loc"] if pass_idx == 0 else env["x2Loc"]
        t = _base_table(env, freqs, amps, periods=periods, tfsf=tfsf, pump=False, probes=probes, nsteps=nsteps,
                        flags_extra=_arith_flags(fma, fp32, False))
        wp = wpu[inv]
        A, B, C = lorentz_abc_many(env["delT"], wp, med["w0"], med["gam"], _sc.epsilon_0)
        t.polA[:], t.polB[:], t.polC[:] = A, B, C
        if kerr_lorentz:
            from . import BaseFDTD11
            chi3, einf = float(med["chi3Stat"]), BaseFDTD11.KERR_EPS_INF
            t.cub_a[:], t.cub_b[:], t.cub_c[:] = chi3 ** 2, 2 * einf * chi3, einf ** 2
            t.nl_den0[:], t.nl_den1[:] = _sc.epsilon_0 * einf, _sc.epsilon_0 * chi3
        else:
            cu, qu, on = cubic_abc_many(uniq, wpu, med["w0"][first], med["gam"][first], med["alpha3"], med["chi3Stat"])
            t.cub_a[:], t.cub_b[:], t.cub_c[:] = cu[inv], qu[inv], on[inv]
            t.nl_den0[:] = _sc.epsilon_0 * float(np.sqrt(1.2))
            t.nl_den1[:] = _sc.epsilon_0 * med["chi3Stat"]
        t.share = first[inv].astype(np.int64)
        t.wp = wp
        tables.append(t)
    return tables


def nonlinear_sweep_table(freqs, amps, domainSize, lowLimTim, highLimTim, *, periods=1000, tfsf=True, nsteps=None,
                          fp32=False, newton=False, env=None):
    """Table of ``IntegratorNL1D`` (Solver_Engine.py:220-271; one pass, ports at the slab edges) for members
    (freqs[i], amps[i]); grids from ``envSetup(f, ..., nonLinMed=True)``, the source carries the 0.8 f pump."""
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    amps = np.broadcast_to(np.asarray(amps, dtype=np.float64), freqs.shape).copy()
    if env is None:
        env = envSetup_many(freqs, domainSize, lowLimTim, highLimTim, nonLinMed=True)
    n = len(freqs)
    med = default_medium(n)
    uniq, first, inv = np.unique(freqs, return_index=True, return_inverse=True)
    envu = {k: v[first] for k, v in env.items()}
    wp1u = corrected_plasma_freq(envu, uniq, med["wp"][first], med["w0"][first], med["gam"][first], 1)
    probes = np.stack([env["materialFrontEdge"], env["materialRearEdge"]], axis=1)
    t = _base_table(env, freqs, amps, periods=periods, tfsf=tfsf, pump=True, probes=probes, nsteps=nsteps,
                    flags_extra=_arith_flags(False, fp32, newton))
    wp = wp1u[inv]
    A, B, C = lorentz_abc_many(env["delT"], wp, med["w0"], med["gam"], _sc.epsilon_0)
    t.polA[:], t.polB[:], t.polC[:] = A, B, C
    cu, qu, on = cubic_abc_many(uniq, wp1u, med["w0"][first], med["gam"][first], med["alpha3"], med["chi3Stat"])
    t.cub_a[:], t.cub_b[:], t.cub_c[:] = cu[inv], qu[inv], on[inv]
    t.nl_den0[:] = _sc.epsilon_0 * float(np.sqrt(1.2))
    t.nl_den1[:] = _sc.epsilon_0 * med["chi3Stat"]
    t.share = first[inv].astype(np.int64)
    t.wp = wp
    return t


def build_inputs(table, out, off_beX, off_ceX, off_cmY, off_srcE, off_srcH, n_src, threads=0):
    """Fill ``out`` (float64 host array, e.g. a batch's pinned upload buffer) with every member's CPML profiles and source
    tables at the given offsets (in doubles; off_beX < 0 = profiles shared, not written) -- pf_host_sweep_inputs."""
    n = table.n
    spec = np.zeros(n, dtype=np.dtype(nat.PfSetupMember))
    spec["L"], spec["pw"], spec["n_src"] = table.L, table.pw, n_src
    spec["src_kind"], spec["tfsf"], spec["pump"] = table.src_kind, table.tfsf, table.pump
    spec["eps0"] = table.eps0
    for k in MemberTable.SETUP:
        spec[k] = getattr(table, k)
    spec["off_beX"], spec["off_ceX"], spec["off_cmY"] = off_beX, off_ceX, off_cmY
    spec["off_srcE"], spec["off_srcH"] = off_srcE, off_srcH
    nat.check(nat.lib().pf_host_sweep_inputs(spec.ctypes.data, n, out.ctypes.data, int(threads)), "pf_host_sweep_inputs")
