"""GPU parity of the reference's dormant models through the host mirrors / the C-ABI: Varin Kerr + Raman
ADE, Kerr current, Mur ABC (BaseFDTD11.py:567-609, 762-788) and the Drude J-form scratch script (TESTBOXDIPSERSE.py:79-94),
against vectors produced by the UNMODIFIED reference functions / script (oracle/make_dormant_golden.py).  Bit-exact."""
import ast
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import BaseFDTD11, drude_sandbox
    from test_host_layer import build_objects
    return BaseFDTD11, drude_sandbox, build_objects


def test_dormant_leaf_ops_match_the_reference(mods):
    B, _, build_objects = mods
    g = np.load(os.path.join(ROOT, "tests", "golden", "dormant_leaf_ops.npz"), allow_pickle=False)
    k = ast.literal_eval(str(g["scalars"]))
    spec = ast.literal_eval(str(g["spec"]))
    V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=spec["freq"], dom=spec["dom"], win=list(spec["win"])))
    assert (P.Nz, P.materialFrontEdge, P.materialRearEdge, P.delT, P.dz) == (k["Nz"], k["mf"], k["mr"], k["delT"], k["dz"])
    assert (V.chi1Stat, V.chi3Stat, V.alpha3, V.gammaE, V.omega_0E) == (k["chi1Stat"], k["chi3Stat"], k["alpha3"], k["gammaE"], k["omega_0E"])
    V.nonLin3gammaE = k["nonLin3gammaE"]
    V.Ex, V.tempTempVarE = g["in_Ex"].copy(), g["in_tempTempVarE"].copy()
    V.Qx3, V.Gx3, V.Jx = g["in_Qx3"].copy(), g["in_Gx3"].copy(), g["in_Jx"].copy()
    V.polarisationCurr, V.Pbar3 = g["in_polarisationCurr"].copy(), g["in_Pbar3"].copy()
    for r in range(3):
        assert np.array_equal(B.ADE_NonLin_Pol_Ex_Pbar(V, P), g[f"r{r}_Pbar3"]), r
        Jx, Pol = B.ADE_Lin_Curr_And_Pol_Varin(V, P)
        assert np.array_equal(Jx, g[f"r{r}_Jx"]) and np.array_equal(Pol, g[f"r{r}_P"]), r
        G, Q, _ = B.ADE_Nonlin_Q_and_G(V, P)
        assert np.array_equal(G, g[f"r{r}_Gx3"]) and np.array_equal(Q, g[f"r{r}_Qx3"]), r
        assert np.array_equal(B.KerrNonlin(V, P, r), g[f"r{r}_JxKerr"]), r
        assert np.array_equal(B.MUR1DEx(V, P, C_V, C_P), g[f"r{r}_Ex"]), r
        V.tempTempVarE = V.tempTempVarE * 0.5 + 0.25 * V.Ex
        assert np.array_equal(V.tempTempVarE, g[f"r{r}_Eold_next"])


def test_drude_scratch_script_matches_the_reference(mods):
    _, drude_sandbox, _ = mods
    g = np.load(os.path.join(ROOT, "tests", "golden", "drude_sandbox.npz"), allow_pickle=False)
    Ex, Hy, Jx = drude_sandbox.run(int(g["domain"]), int(g["tim"]), float(g["freq"]), int(g["nl"]), int(g["src"]), int(g["matFront"]))
    assert np.array_equal(Ex, g["Ex"]) and np.array_equal(Hy, g["Hy"]) and np.array_equal(Jx, g["Jx"])
    assert np.max(np.abs(Jx)) > 0


def test_dormant_ops_reject_bad_descriptors(mods):
    import ctypes
    from pyfdtd_b200 import _native as nat
    d = nat.PfDormant()
    assert nat.lib().pf_varin_pbar(ctypes.byref(d), None) == -1
    dj = nat.PfDrudeJ()
    assert nat.lib().pf_drude_j_run(ctypes.byref(dj), 0, 1, None) == -1
