"""Shim loader for the UNMODIFIED reference (fraserac/Py-FDTD_PIC) -- TEST INFRASTRUCTURE ONLY.

The reference cannot be imported as shipped: it imports a
non-existent ``Tests`` package, matplotlib/pyttsx3 (absent here), runs a full simulation at
import of MasterController and defines ``__repr__`` on two jitclasses (rejected by numba 0.65).
This module makes it runnable *without copying or editing any reference file*:

  * stub modules for matplotlib / pyttsx3 / natsort / moviepy are placed in ``sys.modules``;
  * a synthetic ``Tests`` package maps ``Tests.genericStability`` / ``Tests.BulkTest`` onto the
    reference's flat files and provides no-op ``Validation_Physics`` / ``Integration_Tester``;
  * ``MasterController.py`` is exec'd from its source text, truncated before ``def __Main__``
    (MasterController.py:620) with the two ``__repr__`` methods (:362-363, :436-437) removed;
  * for the nonlinear path only, ``CubicEquationSolver.CubicSolver`` (missing in the reference,
    BaseFDTD11.py:836) is supplied as a 4-tuple packer and ``solve`` accepts that tuple and pads
    its result to the complex128[4] that ``Variables.roots`` is typed as (MasterController.py:140).

It only works where ``/root/reference`` exists (the build container); nothing on the GPU box
imports it.  ``oracle/make_golden.py`` uses it to generate ``tests/golden/*.npz``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF_DIR = os.environ.get("PYFDTD_REFERENCE_DIR", "/root/reference")


class _Anything:
    """Object whose every attribute/call/index is a no-op returning another _Anything."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __getitem__(self, k):
        return _Anything()

    def __setattr__(self, k, v):
        pass


def _stub_module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr: _Anything()  # type: ignore[attr-defined]
    sys.modules[name] = mod
    return mod


def _load_flat(modname: str, filename: str) -> types.ModuleType:
    import importlib.util

    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_DIR, filename))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_loaded = None


def load_reference(quiet: bool = True):
    """Return a namespace with the reference's classes and functions (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_DIR):
        raise RuntimeError(f"reference tree {REF_DIR} not present (only exists in the build container)")

    # --- absent third-party modules -> inert stubs
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "pyttsx3", "natsort",
                 "moviepy", "moviepy.video", "moviepy.video.io", "moviepy.video.io.ImageSequenceClip"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub_module(name)
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "__path__"):
        sys.modules["matplotlib"].__path__ = []  # make "import matplotlib.pylab" resolvable

    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)

    sink = io.StringIO()
    ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
    with ctx:
        # --- synthetic Tests package over the flat files
        tests_pkg = types.ModuleType("Tests")
        tests_pkg.__path__ = []
        sys.modules["Tests"] = tests_pkg
        tests_pkg.genericStability = _load_flat("Tests.genericStability", "genericStability.py")
        tests_pkg.BulkTest = _load_flat("Tests.BulkTest", "BulkTest.py")
        _stub_module("Tests.Validation_Physics", VideoMaker=lambda *a, **k: None)
        _stub_module("Tests.Integration_Tester", testerFuncVector=lambda *a, **k: {})

        import CubicEquationSolver  # reference, flat import
        import numpy as np

        if not hasattr(CubicEquationSolver, "CubicSolver"):
            _orig_solve = CubicEquationSolver.solve

            def CubicSolver(a, b, c, d):
                return tuple(float(np.real(x)) for x in (a, b, c, d))

            def solve(*args):
                if len(args) == 1:
                    args = args[0]
                r = np.asarray(_orig_solve(*args), dtype=np.complex128)
                out = np.zeros(4, dtype=np.complex128)
                out[: len(r)] = r
                return out

            CubicEquationSolver.CubicSolver = CubicSolver
            CubicEquationSolver._orig_solve = _orig_solve
            CubicEquationSolver.solve = solve

        import BaseFDTD11  # noqa: F401  (reference)
        import Solver_Engine  # noqa: F401
        import Environment_Setup  # noqa: F401
        import TransformHandler  # noqa: F401

        # --- MasterController: source text, truncated, __repr__ removed, never written to disk
        with open(os.path.join(REF_DIR, "MasterController.py"), "r") as fh:
            src = fh.read()
        src = src[: src.index("def __Main__")]
        lines = src.split("\n")
        keep = []
        skip = 0
        for ln in lines:
            if skip:
                skip -= 1
                continue
            if ln.strip().startswith("def __repr__"):
                skip = 1  # the single-line body
                continue
            keep.append(ln)
        mc = types.ModuleType("MasterController")
        mc.__file__ = os.path.join(REF_DIR, "MasterController.py")
        sys.modules["MasterController"] = mc
        cwd = os.getcwd()
        exec(compile("\n".join(keep), mc.__file__, "exec"), mc.__dict__)
        os.chdir(cwd)

    ns = types.SimpleNamespace(
        MC=mc,
        BaseFDTD11=sys.modules["BaseFDTD11"],
        SE=sys.modules["Solver_Engine"],
        envDef=sys.modules["Environment_Setup"],
        transH=sys.modules["TransformHandler"],
        gStab=sys.modules["Tests.genericStability"],
        CES=CubicEquationSolver,
    )
    _loaded = ns
    return ns


def build_objects(ref, freq_in, domainSize, minim, maxim, *, mode, source="sine", tfsf=True,
                  periods=1000.0, epsRe=1.0, atten_amount=10, vid_interval=50, amplitude=1.0):
    """Replicate MasterController.__Main__ (:620-663) object construction for a chosen mode.

    mode in {"free", "lorentz", "nl"}; source in {"sine", "gauss"}.
    """
    nonLin = mode == "nl"
    lor = mode == "lorentz"
    tup = ref.envDef.envSetup(freq_in, domainSize, minim, maxim, nonLinMed=nonLin, LorMed=lor)
    P = ref.MC.Params(*tup, False, domainSize, freq_in, 20)
    P.vidInterval = vid_interval
    V = ref.MC.Variables(P.Nz, P.timeSteps, P.vidInterval, atten_amount)
    C_P = ref.MC.CPML_Params(P.dz)
    C_V = ref.MC.CPML_Variables(P.Nz, P.timeSteps)
    P.atten = False
    P.epsRe = epsRe
    P.CPMLXp = True
    P.CPMLXm = True
    P.TFSF = tfsf
    P.Gaussian = source == "gauss"
    P.SineCont = source == "sine"
    P.Periods = periods
    P.Amplitude = amplitude
    P.LorentzMed = lor
    P.nonLinMed = nonLin
    P.FreeSpace = mode == "free"
    P.julia = False
    P.testMode = True
    return V, P, C_V, C_P
