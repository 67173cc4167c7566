"""The three time-stepping integrators, their source / boundary managers and probes.

Mirror of the reference's ``Solver_Engine`` module: ``IntegratorFreeSpace1D`` (:142),
``IntegratorNL1D`` (:220), ``IntegratorLinLor1D`` (:275), ``boundCondManager`` (:70),
``SourceManager`` (:89), ``Sig_Mod`` (:133), ``probeSim`` (:16), ``vidMake`` (:57).  Signatures and the
returned 8-tuple ``(Ex, Hy, Exs, Hys, psi_Ex, psi_Hy, x1ColBe, x1ColAf)`` are the reference's.

The per-pass setup is host-side (BaseFDTD11 mirror).  The ``for counts in range(P.timeSteps)`` loop
(:167 / :236 / :294) -- the hot path -- is ONE call into libpyfdtd_b200 per pass (``pf_run_pass``),
which is also where the reference has its own (unimplemented) external-integrator hook
``JH.Jul_Integrator_Prep`` (:164-165).

Engine selection: ``ENGINE = "auto"`` uses the fused, temporally blocked tile engine whenever the
coefficient arrays have its canonical piecewise form (always true for arrays built by this module's
setup chain) and falls back to the general one-kernel-per-leaf-op engine otherwise (still CUDA).
"""
from __future__ import annotations

import numpy as np

from . import BaseFDTD11
from . import _device as dev
from . import _native as nat
from . import genericStability as gStab

ENGINE = "auto"      # "auto" | "tile" | "ops"
USE_FMA = False      # True: allow FMA contraction in the kernels (PF_F_FMA; not bit-identical)
CUBIC = "closed"     # "closed": the reference's closed-form cubic root (CubicEquationSolver.solve);
                     # "newton": PF_F_NEWTON, the same root by Newton iteration (~3x fewer instructions; differs by the
                     # closed form's cancellation error, <= 1e-10 absolute on Acubic)
KERR_LORENTZ = False  # True: IntegratorLinLor1D runs the Kerr-Lorentz composition (PF_LORENTZ_NL, BASELINE config 5's
                      # "dispersive and nonlinear" material; not a reference integrator -- see include/pyfdtd_b200.h)
USE_FP32 = False     # True: optional single-precision mode of the tile engine (PF_F_FP32; stated tolerance 1e-5
                     # of the trace peak -- not a parity mode; arrays stay fp64 at the boundary)
LAST_RUN_INFO = {}   # engine used, bytes moved, kernel launches of the last pass (for tests / bench)


# ------------------------------------------------------------------------------------------------ probes / history
def probeSim(V, P, C_V, C_P, counts, val="null", af=False, granAttn=25, whichField="Ex", attenRead=False,
             nonlinear=False):
    """Solver_Engine.py:16-54 -- record one probe sample (host-side API; the integrators record
    probes on the device every step and store whole traces)."""
    if nonlinear:
        if whichField == "Ex":
            V.Port1[counts] = V.Ex[P.materialFrontEdge]
            V.Port2[counts] = V.Ex[P.materialRearEdge]
            return V.Port1, V.Port2
        return V.x1ColBe
    if attenRead:
        for i, idx in enumerate(atten_probe_cells(V, P, granAttn)):
            if whichField == "Ex":
                V.x1Atten[i][counts] = V.Ex[idx]
        return V.x1Atten
    if af:
        V.x1ColAf[counts] = val
        return V.x1ColAf
    V.x1ColBe[counts] = val
    return V.x1ColBe


def atten_probe_cells(V, P, granAttn=25):
    """Attenuation probe cells: materialFrontEdge + granAttn*i (Solver_Engine.py:32)."""
    return np.arange(P.materialFrontEdge, granAttn * V.attenAmnt + P.materialFrontEdge, granAttn)[: len(V.x1Atten)]


def vidMake(V, P, C_V, C_P, counts, field, whichField="Ex"):
    """Solver_Engine.py:57-68."""
    row = int(counts / P.vidInterval)
    if whichField == "Ex":
        if row < len(V.Ex_History):
            V.Ex_History[row] = field
        return V.Ex_History
    return "vidMake has gone to the end of the returns "


# ------------------------------------------------------------------------------------------------ managers
def boundCondManager(V, P, C_V, C_P):
    """Solver_Engine.py:70-87 -- build every CPML coefficient array."""
    if P.CPMLXp or P.CPMLXm:
        (C_V.sigma_Ex, C_V.sigma_Hy, C_V.alpha_Ex, C_V.alpha_Hy, C_V.kappa_Ex,
         C_V.kappa_Hy) = BaseFDTD11.CPML_ScalingCalc(V, P, C_V, C_P)
        C_V.beX, C_V.ceX = BaseFDTD11.CPML_Ex_RC_Define(V, P, C_V, C_P)
        C_V.bmY, C_V.cmY = BaseFDTD11.CPML_HY_RC_Define(V, P, C_V, C_P)
        C_V.eLoss_CPML, C_V.Ca, C_V.Cb, C_V.Cc = BaseFDTD11.CPML_Ex_Update_Coef(V, P, C_V, C_P)
        C_V.mLoss_CPML, C_V.C1, C_V.C2, C_V.C3 = BaseFDTD11.CPML_Hy_Update_Coef(V, P, C_V, C_P)
        C_V.den_Exdz, C_V.den_Hydz = BaseFDTD11.denominators(V, P, C_V, C_P)
    return C_V


def SourceManager(V, P, C_V, C_P):
    """Solver_Engine.py:89-124 -- per-step source tables Exs, Hys."""
    if P.SineCont == True:  # noqa: E712  (kept: the reference tests identity with True)
        Exs, Hys = BaseFDTD11.SmoothTurnOn(V, P)
        Exp, Hyp = np.zeros(P.timeSteps), np.zeros(P.timeSteps)
        if P.nonLinMed:
            Exp, Hyp = BaseFDTD11.SmoothTurnOn(V, P, tempfreq=P.freq_in * 0.8)   # pump at 0.8 f
            Exp = np.asarray(Exp) * P.courantNo
            Hyp = np.asarray(Hyp) * P.courantNo
            Exp = Exp * 0.1
            Hyp = Hyp * 0.01
        Exs = np.asarray(Exs) * P.courantNo + Exp
        Hys = np.asarray(Hys) * P.courantNo + Hyp
        if P.TFSF == True:  # noqa: E712
            Hys = Hys * (1 / P.CharImp)
        return Exs, Hys
    if P.Gaussian == True:  # noqa: E712
        Exs = BaseFDTD11.Gaussian(V, P)
        Hys = BaseFDTD11.Gaussian(V, P) if P.TFSF == True else np.zeros(len(Exs))  # noqa: E712
        return Exs, Hys
    return [], []


def SechArr(a):
    """Solver_Engine.py:126-131."""
    a = np.where(a > 50, 50, a)
    return 1 / np.cosh(a)


def Sig_Mod(V, P, sig, AmpCarr=1, AmpMod=1, tau=14.6e-10):
    """Solver_Engine.py:133-140 -- sech envelope."""
    t = np.arange(len(sig)) * (P.delT)
    return (AmpMod + sig) * SechArr((2 * np.pi * (1 / tau)) * t)


# ------------------------------------------------------------------------------------------------ the hot loop
def _pick_engine(P, arrs, Jx, probe_idx):
    if ENGINE == "ops":
        return nat.PF_ENGINE_OPS, None
    canon = dev.canonical_form(P, arrs)
    if Jx is not None and (USE_FMA or USE_FP32):
        canon = None            # the tile engine carries a current slot in exact arithmetic only
    slab_src_clash = P.TFSF and (P.materialFrontEdge - 1 <= P.nzsrc - 1 < P.materialRearEdge)
    ok = canon is not None and dev.probes_ok_for_tiles(probe_idx) and not slab_src_clash
    if ENGINE == "tile" and not ok:
        raise ValueError("ENGINE='tile' requested but the grid is not in the tile engine's canonical form")
    return (nat.PF_ENGINE_TILE, canon) if ok else (nat.PF_ENGINE_OPS, None)


class PassRun:
    """One pass of the time loop in flight on the device: ``start`` uploads and enqueues every launch of the pass on a
    CUDA stream and returns at once; ``finish`` waits, downloads and writes the results back into V / C_V.  The two passes
    of IntegratorFreeSpace1D / IntegratorLinLor1D do not read each other's fields (each begins from zeroed fields), so
    _two_pass starts the second while the first is still running: the host-side setup of pass 1 and both passes' kernels
    -- a default grid is 7-15 tiles, far less than the machine -- overlap."""

    def __init__(self, V, P, C_V, C_P, mode, do_pol, Exs, Hys, probe_idx, snapshots=False, n0=0, nsteps=None, stream=None,
                 stage_tag="stage", defer=False):
        torch = nat.require_cuda()
        lib = nat.lib()
        self.torch, self.V, self.C_V, self.P, self.mode = torch, V, C_V, P, mode
        T = int(P.timeSteps)
        nsteps = T - n0 if nsteps is None else int(nsteps)
        self.n0, self.nsteps = n0, nsteps
        L = len(V.Ex)
        arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
        Jx = V.Jx if np.any(V.Jx != 0.0) else None
        engine, canon = _pick_engine(P, arrs, Jx, probe_idx)
        scal = BaseFDTD11.grid_scalars(V, P, kerr_lorentz=(mode == "lorentz_nl"))
        flags = BaseFDTD11.grid_flags(P, USE_FMA, USE_FP32, CUBIC == "newton")
        if USE_FP32 and engine != nat.PF_ENGINE_TILE:
            raise ValueError("USE_FP32 is a mode of the tile engine; this grid needs the general per-op engine")
        if canon is not None:
            scal.update(cE0=canon[0], cE1=canon[1], cH0=canon[2], cH1=canon[3], c2_pml=canon[4])
            flags |= nat.PF_F_CANONICAL
        self.engine = engine
        self.stream = torch.cuda.current_stream() if stream is None else stream
        with torch.cuda.stream(self.stream):
            g = dev.DeviceGrid(L=L, T=T, arrays=arrs, scalars=scal, srcE=np.asarray(Exs) / P.courantNo,
                               srcH=np.asarray(Hys) / P.courantNo, probe_idx=list(probe_idx), flags=flags, Jx=Jx,
                               stage_tag=stage_tag)
            self.g = g
            stream_ptr = nat.current_stream_ptr()
            scratch = None
            sbytes = lib.pf_run_scratch_bytes(g.ref(), 1, engine)
            if sbytes:
                scratch = torch.empty(sbytes, dtype=torch.uint8, device=g.device)
            self.scratch = scratch
            snap_t = None
            rows = int(P.timeSteps / P.vidInterval)
            if snapshots and rows > 0:
                snap_t = torch.zeros((rows, L), dtype=torch.float64, device=g.device)
            self.snap_t, self.rows = snap_t, rows
            mode_id = dev.MODE_ID[mode]
            self.pprev2 = None

            def call(first, count):
                nat.check(lib.pf_run_pass(g.ref(), mode_id, int(do_pol), first, count, engine,
                                          snap_t.data_ptr() if snap_t is not None else None,
                                          int(P.vidInterval) if snap_t is not None else 0, rows if snap_t is not None else 0,
                                          scratch.data_ptr() if scratch is not None else None, sbytes, stream_ptr), "pf_run_pass")

            def enqueue():
                if mode in ("lorentz", "lorentz_nl") and do_pol and nsteps >= 1:
                    # keep P^{N-2} as well so V.tempTempVarPol / V.tempVarPol end up as the reference leaves them
                    if nsteps > 1:
                        call(n0, nsteps - 1)
                    self.pprev2 = g.tensor_view("Pprev").clone()
                    call(n0 + nsteps - 1, 1)
                elif nsteps > 0:
                    call(n0, nsteps)
            self._enqueue = enqueue
        self.launches = 0
        if not defer:
            self.launch()

    def launch(self):
        """Enqueue the pass's kernels on its stream (done by the constructor unless ``defer``: the state is captured and
        uploaded at construction, so the host arrays may change before this is called)."""
        if self._enqueue is None:
            return
        lib = nat.lib()
        n_before = lib.pf_launch_count()
        with self.torch.cuda.stream(self.stream):
            self._enqueue()
        self._enqueue = None
        self.launches = lib.pf_launch_count() - n_before

    def finish(self, write_state=True):
        """Wait for the pass, return its probe traces [len(probe_idx), timeSteps]; with write_state the final fields go back
        into V / C_V as the reference's loop leaves them (a pass whose fields the next prepare_pass zeroes can skip that)."""
        torch, V, C_V, P, g, mode = self.torch, self.V, self.C_V, self.P, self.g, self.mode
        with torch.cuda.stream(self.stream):
            if not write_state:
                traces = g.fetch_probes()
                LAST_RUN_INFO.update(engine="tile" if self.engine == nat.PF_ENGINE_TILE else "ops", h2d_bytes=g.h2d_bytes,
                                     d2h_bytes=g.d2h_bytes, launches=self.launches, cells=g.L, steps=self.nsteps)
                return traces
            out = g.fetch(["Ex", "Hy", "Dx", "P", "Pprev", "psiE", "psiH", "Acubic"])
            V.Ex, V.Hy, V.Dx = out["Ex"], out["Hy"], out["Dx"]
            C_V.psi_Ex, C_V.psi_Hy = out["psiE"], out["psiH"]
            if mode in ("lorentz", "lorentz_nl"):
                V.polarisationCurr = out["P"]
                V.tempVarPol = out["Pprev"]
                if self.pprev2 is not None:
                    V.tempTempVarPol = self.pprev2.cpu().numpy()
            if mode in ("nl", "lorentz_nl"):
                V.Acubic = out["Acubic"]
            snap_t, rows, n0, nsteps, L = self.snap_t, self.rows, self.n0, self.nsteps, g.L
            if snap_t is not None:
                n_abs = np.arange(rows) * int(P.vidInterval)
                done = np.flatnonzero((n_abs > 0) & (n_abs >= n0) & (n_abs < n0 + nsteps))
                if len(done):                               # a contiguous block of rows
                    lo, hi = int(done[0]), int(done[-1]) + 1
                    if "Ex_History" in V.__dict__ or not hasattr(V, "_build_ex_history") or getattr(V, "_rows", rows) != rows:
                        stage = dev.pinned_buffer(rows * L, tag="history").view(rows, L)     # already materialised: update it now
                        stage.copy_(snap_t, non_blocking=True)
                        torch.cuda.current_stream().synchronize()
                        V.Ex_History[lo:hi] = stage.numpy()[lo:hi]
                    else:
                        # leave the rows on the device; V.Ex_History downloads them when it is first read (vidMake / VideoMaker)
                        V.__dict__.setdefault("_pending_history", []).append((snap_t, lo, hi))
        LAST_RUN_INFO.update(engine="tile" if self.engine == nat.PF_ENGINE_TILE else "ops", h2d_bytes=g.h2d_bytes,
                             d2h_bytes=g.d2h_bytes, launches=self.launches, cells=g.L, steps=nsteps)
        return out["probe_out"]


def run_time_loop(V, P, C_V, C_P, mode, do_pol, Exs, Hys, probe_idx, snapshots=False, n0=0, nsteps=None):
    """The reference's per-pass time loop, executed by libpyfdtd_b200 on the current CUDA device.

    Uploads V/C_V state + coefficients (one H2D), runs ``nsteps`` steps of integrator ``mode`` starting at
    absolute step ``n0``, downloads state + probe traces (one D2H) and writes them back into V / C_V.
    Returns the probe traces, shape [len(probe_idx), timeSteps].
    """
    return PassRun(V, P, C_V, C_P, mode, do_pol, Exs, Hys, probe_idx, snapshots=snapshots, n0=n0, nsteps=nsteps).finish()


def prepare_pass(V, P, C_V, C_P, lorentz, nonlinear=False):
    """Host-side setup the reference repeats at the start of every pass (Solver_Engine.py:144-160,
    223-232, 278-287): zero the fields, rebuild update / CPML coefficients, re-apply the dispersion
    correction of the plasma frequency (cumulative!), build the source tables.  Returns (C_V, Exs, Hys)."""
    (V.tempVarPol, V.tempTempVarE, V.tempVarE, V.tempTempVarPol, V.polarisationCurr, V.Ex, V.Dx,
     V.Hy) = BaseFDTD11.FieldInit(V, P)
    V.UpHyMat, V.UpExMat = BaseFDTD11.EmptySpaceCalc(V, P)
    free = not lorentz and not nonlinear
    if free:
        (V.epsilon, V.mu, V.UpExHcompsCo, V.UpExSelf, V.UpHyEcompsCo,
         V.UpHySelf) = BaseFDTD11.Material(V, P)
        V.UpHyMat, V.UpExMat = BaseFDTD11.UpdateCoef(V, P)
    C_V = BaseFDTD11.CPML_FieldInit(V, P, C_V, C_P)
    C_V = boundCondManager(V, P, C_V, C_P)
    if not free:
        _, _, _, V.plasmaFreqE, _ = gStab.spatialStab(P.timeSteps, P.Nz, P.dz, P.freq_in, P.delT,
                                                       V.plasmaFreqE, V.omega_0E, V.gammaE)
    Exs, Hys = SourceManager(V, P, C_V, C_P)
    if free:
        tauIn = 1 / (P.freq_in / 5)
        Exs = Sig_Mod(V, P, Exs, tau=tauIn)
        Hys = Sig_Mod(V, P, Hys, AmpMod=1 / P.CharImp, tau=tauIn)
    return C_V, Exs, Hys


_SIDE = {}


def _side_stream(torch):
    d = torch.cuda.current_device()
    if d not in _SIDE:
        _SIDE[d] = torch.cuda.Stream(device=d)
    return _SIDE[d]


def _two_pass(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf, lorentz):
    """Shared body of IntegratorFreeSpace1D / IntegratorLinLor1D: pass 0 = incident run (probe x1Loc),
    pass 1 = run with the medium's polarisation (probe x2Loc, history, attenuation probes)."""
    torch = nat.require_cuda()
    n = np.arange(P.timeSteps)
    mode = ("lorentz_nl" if KERR_LORENTZ else "lorentz") if lorentz else "free"
    # Pass 0 is enqueued on a side stream and left running: prepare_pass of pass 1 starts from zeroed fields and reads
    # nothing pass 0 writes (its only output is the x1Loc trace), so the reference's order of host-side effects is kept
    # while pass 1's setup and kernels overlap pass 0's.
    side = _side_stream(torch)
    side.wait_stream(torch.cuda.current_stream())
    C_V, Exs, Hys = prepare_pass(V, P, C_V, C_P, lorentz)
    V.test = 0
    run0 = PassRun(V, P, C_V, C_P, mode, False, Exs, Hys, [P.x1Loc], stream=side, stage_tag="stage0", defer=True)
    C_V, Exs, Hys = prepare_pass(V, P, C_V, C_P, lorentz)
    V.test = 0
    atten = list(atten_probe_cells(V, P)) if P.atten else []
    run1 = PassRun(V, P, C_V, C_P, mode, lorentz, Exs, Hys, [P.x2Loc] + atten, snapshots=True)
    run0.launch()          # pass 1 (snapshots, more probes) is the longer one: its kernels are enqueued first
    traces = run0.finish(write_state=False)
    V.x1ColBe = np.where(n <= probeReadFinishBe, traces[0], V.x1ColBe)
    traces = run1.finish()
    LAST_RUN_INFO["launches"] = run0.launches + run1.launches
    window = n >= probeReadStartAf
    V.x1ColAf = np.where(window, traces[0], V.x1ColAf)
    for k in range(len(atten)):
        V.x1Atten[k] = np.where(window, traces[1 + k], V.x1Atten[k])
    return V.Ex, V.Hy, Exs, Hys, C_V.psi_Ex, C_V.psi_Hy, V.x1ColBe, V.x1ColAf


def IntegratorFreeSpace1D(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf):
    """Solver_Engine.py:142-215 -- vacuum / dielectric slab + CPML, two passes."""
    return _two_pass(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf, lorentz=False)


def IntegratorLinLor1D(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf):
    """Solver_Engine.py:275-371 -- Lorentz ADE medium + CPML, two passes."""
    return _two_pass(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf, lorentz=True)


def IntegratorNL1D(V, P, C_V, C_P, probeReadFinishBe, probeReadStartAf):
    """Solver_Engine.py:220-271 -- cubic nonlinear medium, per-cell cubic solve every step, one pass."""
    C_V, Exs, Hys = prepare_pass(V, P, C_V, C_P, lorentz=False, nonlinear=True)
    traces = run_time_loop(V, P, C_V, C_P, "nl", False, Exs, Hys, [P.materialFrontEdge, P.materialRearEdge],
                           snapshots=True)
    V.Port1, V.Port2 = traces[0], traces[1]
    return V.Ex, V.Hy, Exs, Hys, C_V.psi_Ex, C_V.psi_Hy, V.x1ColBe, V.x1ColAf
