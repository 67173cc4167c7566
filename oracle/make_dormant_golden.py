"""Golden vectors of the reference's DORMANT leaf functions and of its Drude scratch script -- run in the build container
only (needs /root/reference).  Writes tests/golden/dormant_leaf_ops.npz and tests/golden/drude_sandbox.npz.

    PYTHONBREAKPOINT=0 python oracle/make_dormant_golden.py [--full-drude]

* Leaf ops (BaseFDTD11.py: ADE_NonLin_Pol_Ex_Pbar :567, ADE_Lin_Curr_And_Pol_Varin :580, ADE_Nonlin_Q_and_G :596,
  KerrNonlin :762, MUR1DEx :769): the UNMODIFIED reference functions, called through oracle/ref_shim.py on seeded random
  state, three rounds in a row (state carried over), every array recorded after every call.  ADE_NonLin_Pol_Ex_Pbar holds a
  debugging ``breakpoint()`` that fires whenever Pbar3 is non-zero on entry: PYTHONBREAKPOINT=0 makes it a no-op.
* Drude script (TESTBOXDIPSERSE.py): the unmodified script TEXT is exec'd with matplotlib stubbed; only the four size
  literals (domain, tim, src, matFront) are substituted in the text.  NOTE: the script's scheme is unstable as written (its E
  coefficient 2 dt/(2 eps0 + beta dt) makes the effective Courant number ~sqrt(2)): |Ex| grows ~5.7x per step and overflows
  to inf/nan after ~400 steps whatever the size, so the vector stops at 150 steps (|Ex| ~ 1e110, every bit still
  significant) with the slab front 50 cells behind the source so that the J update is exercised.  --full-drude also runs
  the script exactly as written (domain 14000, tim 5000: minutes of CPython loops, all-nan result).
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

os.environ.setdefault("PYTHONBREAKPOINT", "0")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import versions  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def leaf_ops():
    ref = ref_shim.load_reference()
    spec = dict(freq=9e9, dom=0.15, win=(300, 320))
    V, P, C_V, C_P = ref_shim.build_objects(ref, spec["freq"], spec["dom"], *spec["win"], mode="lorentz")
    rng = np.random.default_rng(2024)
    n = len(V.Qx3)                      # the reference allocates these members with Nz entries
    L = len(V.Ex)
    V.Ex = rng.uniform(-2, 2, L)
    V.tempTempVarE = rng.uniform(-2, 2, L)
    V.Qx3 = rng.uniform(0, 3, n)
    V.Gx3 = rng.uniform(-1e10, 1e10, n)
    V.Jx = np.zeros(L)
    V.Jx[:] = rng.uniform(-1e-2, 1e-2, L)
    V.polarisationCurr = rng.uniform(-1e-11, 1e-11, L)
    V.nonLin3gammaE = 2 * np.pi * 1e9 * 0.3
    V.Pbar3 = np.zeros(n)
    state0 = {k: np.array(getattr(V, k)) for k in ("Ex", "tempTempVarE", "Qx3", "Gx3", "Jx", "polarisationCurr", "Pbar3")}
    rec = {}
    B = ref.BaseFDTD11
    for r in range(3):
        rec[f"r{r}_Pbar3"] = np.array(B.ADE_NonLin_Pol_Ex_Pbar(V, P))
        Jx, Pol = B.ADE_Lin_Curr_And_Pol_Varin(V, P)
        rec[f"r{r}_Jx"], rec[f"r{r}_P"] = np.array(Jx), np.array(Pol)
        G, Q, _ = B.ADE_Nonlin_Q_and_G(V, P)
        rec[f"r{r}_Gx3"], rec[f"r{r}_Qx3"] = np.array(G), np.array(Q)
        rec[f"r{r}_JxKerr"] = np.array(B.KerrNonlin(V, P, r))
        rec[f"r{r}_Ex"] = np.array(B.MUR1DEx(V, P, C_V, C_P))
        V.tempTempVarE = V.tempTempVarE * 0.5 + 0.25 * V.Ex          # something new for the next round
        rec[f"r{r}_Eold_next"] = np.array(V.tempTempVarE)
    scal = dict(mf=int(P.materialFrontEdge), mr=int(P.materialRearEdge), Nz=int(P.Nz), dz=float(P.dz), delT=float(P.delT),
                c0=float(P.c0), permit_0=float(P.permit_0), chi1Stat=float(V.chi1Stat), chi3Stat=float(V.chi3Stat),
                alpha3=float(V.alpha3), gammaE=float(V.gammaE), omega_0E=float(V.omega_0E),
                nonLin3gammaE=float(V.nonLin3gammaE), nonLin3Omega_0E=float(V.nonLin3Omega_0E))
    np.savez_compressed(os.path.join(OUT, "dormant_leaf_ops.npz"), spec=np.array(repr(spec)), versions=np.array(repr(versions())),
                        scalars=np.array(repr(scal)), **{f"in_{k}": v for k, v in state0.items()}, **rec)
    print("dormant_leaf_ops.npz:", len(rec), "arrays, mf/mr", scal["mf"], scal["mr"], "L", L, "n", n)


def drude(domain=None, tim=None, src=None, matFront=None):
    import types
    import contextlib
    import io
    for name in ("matplotlib", "matplotlib.pylab"):
        if name not in sys.modules:
            ref_shim._stub_module(name)
    sys.modules["matplotlib"].__path__ = []
    text = open(os.path.join(ref_shim.REF_DIR, "TESTBOXDIPSERSE.py")).read()
    for key, val in (("domain", domain), ("tim", tim), ("src", src), ("matFront", matFront)):
        if val is not None:
            text, nsub = re.subn(rf"^{key}\s*=\s*\d+", f"{key} = {val}", text, count=1, flags=re.M)
            assert nsub == 1, key
    g = {"__name__": "TESTBOXDIPSERSE"}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(text, "TESTBOXDIPSERSE.py", "exec"), g)
    return {k: np.asarray(g[k], dtype=np.float64) for k in ("Ex", "Hy", "Jx", "Hys")} | {
        k: g[k] for k in ("domain", "tim", "src", "matFront", "matRear", "dz", "dt", "cour", "betaE", "kapE", "perm0", "freq", "nl")}


def main():
    leaf_ops()
    small = drude(domain=1400, tim=150, src=210, matFront=260)
    np.savez_compressed(os.path.join(OUT, "drude_sandbox.npz"), versions=np.array(repr(versions())),
                        **{k: np.asarray(v) for k, v in small.items()})
    print("drude_sandbox.npz: max|Ex|", float(np.max(np.abs(small["Ex"]))), "max|Jx|", float(np.max(np.abs(small["Jx"]))))
    if "--full-drude" in sys.argv:
        full = drude()
        np.savez_compressed(os.path.join(OUT, "drude_sandbox_full.npz"), versions=np.array(repr(versions())),
                            **{k: np.asarray(v) for k, v in full.items()})
        print("drude_sandbox_full.npz: max|Ex|", float(np.max(np.abs(full["Ex"]))))


if __name__ == "__main__":
    main()
