"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_shim.py).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py [--only NAME] [--list]

Each file stores the inputs that define the case (so tests can rebuild it without the reference),
the reference's outputs, and the library versions / physical constants used.  Library versions
matter because the reference has no pinned requirements.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> kwargs for ref_shim.build_objects (+ freq/domain/window)
SINGLE_CASES = {
    # config 1 family: vacuum/dielectric Yee + CPML (IntegratorFreeSpace1D)
    "free_sine_eps4": dict(freq=9e9, dom=0.2, win=(400, 600), mode="free", source="sine", periods=1000, epsRe=4.0),
    "free_gauss_eps4": dict(freq=9e9, dom=0.2, win=(400, 600), mode="free", source="gauss", periods=1000, epsRe=4.0),
    "free_gauss_notfsf": dict(freq=7e9, dom=0.15, win=(300, 350), mode="free", source="gauss", periods=1.0, epsRe=2.25, tfsf=False),
    # config 2 family: Lorentz ADE + CPML (IntegratorLinLor1D)
    "lorentz_sine": dict(freq=9e9, dom=0.2, win=(400, 600), mode="lorentz", source="sine", periods=1000),
    "lorentz_gauss": dict(freq=9e9, dom=0.2, win=(400, 600), mode="lorentz", source="gauss", periods=1000),
    "lorentz_sine_6g": dict(freq=6e9, dom=0.25, win=(700, 800), mode="lorentz", source="sine", periods=3.0),
    # config 3 family: cubic nonlinear (IntegratorNL1D, needs the CubicSolver shim)
    "nl_sine": dict(freq=9e9, dom=0.12, win=(500, 520), mode="nl", source="sine", periods=1000),
    "nl_sine_amp": dict(freq=8e9, dom=0.12, win=(480, 500), mode="nl", source="sine", periods=1000, amplitude=1.0),
    # default __Main__ geometry (MasterController.py:620-663) with LorMed=True: the SURVEY 8c scalars
    "lorentz_default_full": dict(freq=9e9, dom=0.7, win=(7000, 8000), mode="lorentz", source="sine", periods=1000),
    "free_default_full": dict(freq=9e9, dom=0.7, win=(7000, 8000), mode="free", source="sine", periods=1000),
}


def versions():
    import numba
    import scipy
    import scipy.constants as sc
    return dict(numpy=np.__version__, scipy=scipy.__version__, numba=numba.__version__,
                python=sys.version.split()[0], epsilon_0=sc.epsilon_0, mu_0=sc.mu_0, c=sc.speed_of_light)


def run_single(name, spec):
    ref = ref_shim.load_reference()
    kw = dict(spec)
    freq, dom, win, mode = kw.pop("freq"), kw.pop("dom"), kw.pop("win"), kw.pop("mode")
    with contextlib.redirect_stdout(io.StringIO()):
        V, P, C_V, C_P = ref_shim.build_objects(ref, freq, dom, *win, mode=mode, **kw)
        t0 = time.time()
        V, P, C_V, C_P, Exs, Hys = ref.MC.Controller(V, P, C_V, C_P)
        wall = time.time() - t0
        refl = anal = np.nan
        if mode == "lorentz":
            tvec = np.arange(0, len(V.x1ColBe)) * P.delT
            try:   # RefTester sys.exit()s when the spectral peak is DC (TransformHandler.py:63-65)
                refl = ref.MC.results(V, P, C_V, C_P, tvec, RefCo=True)
            except SystemExit:
                refl = np.nan
            anal = ref.MC.results(V, P, C_V, C_P, tvec, AnalRefCo=True)
    full = name.endswith("_full")
    out = dict(
        spec=json.dumps(dict(spec, freq=freq, dom=dom, win=list(win), mode=mode)),
        versions=json.dumps(versions()),
        ref_wall_seconds=wall,
        Nz=P.Nz, timeSteps=P.timeSteps, pmlWidth=P.pmlWidth, nzsrc=P.nzsrc, mf=P.materialFrontEdge,
        mr=P.materialRearEdge, x1Loc=P.x1Loc, x2Loc=P.x2Loc, dz=P.dz, delT=P.delT,
        courantNo=P.courantNo, period=P.period, Nlam=P.Nlam, plasmaFreqE=V.plasmaFreqE,
        Ex=V.Ex, Hy=V.Hy, psi_Ex=C_V.psi_Ex, psi_Hy=C_V.psi_Hy, x1ColBe=V.x1ColBe, x1ColAf=V.x1ColAf,
        Exs=np.asarray(Exs), Hys=np.asarray(Hys), reflection=refl, analytical_reflection=anal,
        sumEx=float(np.sum(V.Ex)), maxAbsEx=float(np.max(np.abs(V.Ex))), maxAbsHy=float(np.max(np.abs(V.Hy))),
    )
    if not full:
        out.update(beX=C_V.beX, ceX=C_V.ceX, bmY=C_V.bmY, cmY=C_V.cmY, Cb=C_V.Cb, C2=C_V.C2,
                   den_Exdz=C_V.den_Exdz, den_Hydz=C_V.den_Hydz, UpExMat=V.UpExMat, UpHyMat=V.UpHyMat,
                   Ex_History_rows=V.Ex_History[:: max(1, len(V.Ex_History) // 4)])
    if mode == "lorentz":
        out.update(polarisationCurr=V.polarisationCurr, Dx=V.Dx, tempTempVarPol=V.tempTempVarPol)
    if mode == "nl":
        out.update(Acubic=V.Acubic, Port1=V.Port1, Port2=V.Port2, Dx=V.Dx)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: Nz={P.Nz} T={P.timeSteps} ref wall {wall:.1f}s sumEx={out['sumEx']!r}", flush=True)


def run_sweep(name="lorentz_sweep", freq0=6e9, interval=5e8, dom=0.3, win=(2000, 2200)):
    """MasterController.LoopedSim(loop=True) :533-569 -- 20 points, captured through plotter()."""
    ref = ref_shim.load_reference()
    captured = {}

    def plotter(x, yAxisData1=None, yAxisData2=None, **k):
        captured.update(x=np.asarray(x), measured=np.asarray(yAxisData1), analytical=np.asarray(yAxisData2))

    ref.MC.plotter = plotter
    members = []
    orig_controller = ref.MC.Controller

    def controller(V, P, C_V, C_P):
        r = orig_controller(V, P, C_V, C_P)
        members.append(dict(freq=P.freq_in, Nz=P.Nz, T=P.timeSteps, Nlam=P.Nlam, dz=P.dz, delT=P.delT,
                            pw=P.pmlWidth, wp_after=V.plasmaFreqE, maxBe=float(np.max(V.x1ColBe)),
                            maxAf=float(np.max(V.x1ColAf))))
        return r

    ref.MC.Controller = controller
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            V, P, C_V, C_P = ref_shim.build_objects(ref, freq0, dom, *win, mode="lorentz", source="sine", periods=1.0)
            P.Periods = 1.0
            Rep = ref.MC.Reporter()
            t0 = time.time()
            ref.MC.LoopedSim(Rep, V, P, C_V, C_P, False, dom, win[0], win[1], loop=True, Low=freq0, Interval=interval)
            wall = time.time() - t0
    finally:
        ref.MC.Controller = orig_controller
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        spec=json.dumps(dict(freq0=freq0, interval=interval, dom=dom, win=list(win))),
        versions=json.dumps(versions()), ref_wall_seconds=wall, members=json.dumps(members),
        freqs=captured["x"], measured=captured["measured"], analytical=captured["analytical"])
    print(f"{name}: 20 points, ref wall {wall:.1f}s measured={captured['measured'][:4]}", flush=True)


def run_cubic(name="cubic_roots"):
    """Known-answer vectors for CubicEquationSolver.solve (root[0]) over the three real branches."""
    ref = ref_shim.load_reference()
    solve = getattr(ref.CES, "_orig_solve", ref.CES.solve)
    rng = np.random.default_rng(20261017)
    rows = []
    # (i) the NL-path polynomial family: a,b,c from the default medium at 6-10.5 GHz, d = -q^2
    import fdtd_oracle as fo
    med = fo.default_medium()
    for f in np.linspace(6e9, 10.5e9, 8):
        a, b, c = fo.cubic_abc(f, med["wp"], med["w0"], med["gam"], med["alpha3"], med["chi3"])
        for q2 in np.concatenate([10.0 ** rng.uniform(-8, 3, 40), [1e-8 * 1.0001, 1.0, 123.456]]):
            rows.append((a, b, c, -q2))
    # (ii) generic cubics hitting h<=0 (three real roots), h>0, and the degenerate branches
    for _ in range(300):
        r = rng.uniform(-5, 5, 3)
        a = rng.uniform(0.1, 3)
        poly = a * np.poly(r)
        rows.append(tuple(poly))
    for _ in range(200):
        rows.append(tuple(rng.uniform(-4, 4, 4)))
    rows += [(1.0, -6.0, 12.0, -8.0), (2.0, 0.0, 0.0, -16.0), (0.0, 2.0, -3.0, 1.0), (0.0, 0.0, 4.0, -2.0),
             (0.0, 1.0, 2.0, 5.0), (1.0, 0.0, 0.0, 8.0)]
    coeffs = np.array(rows, dtype=np.float64)
    root0 = np.zeros(len(rows), dtype=np.complex128)
    roots = np.full((len(rows), 3), np.nan + 0j, dtype=np.complex128)     # all roots, NaN-padded
    nroots = np.zeros(len(rows), dtype=np.int32)
    for i, (a, b, c, d) in enumerate(coeffs):
        r = np.asarray(solve(float(a), float(b), float(c), float(d)), dtype=np.complex128)
        root0[i] = r[0]
        roots[i, : len(r)] = r
        nroots[i] = len(r)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), coeffs=coeffs, root0=root0, roots=roots, nroots=nroots,
                        versions=json.dumps(versions()))
    print(f"{name}: {len(rows)} polynomials", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--skip-existing", action="store_true")
    a = ap.parse_args()
    names = list(SINGLE_CASES) + ["lorentz_sweep", "cubic_roots"]
    if a.list:
        print("\n".join(names))
        return
    os.makedirs(OUT, exist_ok=True)
    for n in names:
        if a.only and a.only != n:
            continue
        if a.skip_existing and os.path.exists(os.path.join(OUT, n + ".npz")):
            continue
        if n in SINGLE_CASES:
            run_single(n, SINGLE_CASES[n])
        elif n == "lorentz_sweep":
            run_sweep()
        else:
            run_cubic()


if __name__ == "__main__":
    main()
