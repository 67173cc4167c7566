"""CPU tests of the product's host layer: parameter objects, envSetup, coefficient builders and
source tables against the reference's outputs stored in tests/golden (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pyfdtd_b200  # noqa: F401
from pyfdtd_b200 import BaseFDTD11, Environment_Setup as envDef, MasterController as MC, Solver_Engine as SE
from pyfdtd_b200 import _device as dev, _native as nat, genericStability as gStab
from conftest import ROOT, load_golden


def build_objects(spec):
    mode = spec["mode"]
    tup = envDef.envSetup(spec["freq"], spec["dom"], *spec["win"], nonLinMed=(mode == "nl"), LorMed=(mode == "lorentz"))
    P = MC.Params(*tup, False, spec["dom"], spec["freq"], 20)
    P.vidInterval = 50
    V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 10)
    C_P = MC.CPML_Params(P.dz)
    C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
    P.epsRe = spec.get("epsRe", 1.0)
    P.TFSF = spec.get("tfsf", True)
    P.Gaussian = spec.get("source", "sine") == "gauss"
    P.SineCont = spec.get("source", "sine") == "sine"
    P.Periods = spec.get("periods", 1000.0)
    P.Amplitude = spec.get("amplitude", 1.0)
    P.LorentzMed = mode == "lorentz"
    P.nonLinMed = mode == "nl"
    P.FreeSpace = mode == "free"
    return V, P, C_V, C_P


def test_float32_quantisation_of_members():
    """jitclass float32 / int32 members."""
    tup = envDef.envSetup(9e9, 0.7, 7000, 8000)
    P = MC.Params(*tup, False, 0.7, 9e9, 20)
    assert P.courantNo == 0.949999988079071
    assert P.domainSize == 0
    assert (P.Nz, P.timeSteps, P.pmlWidth, P.nzsrc, P.materialFrontEdge, P.materialRearEdge, P.x1Loc, P.x2Loc) == (
        13193, 23997, 2394, 2994, 3594, 13192, 3574, 2894)
    C_P = MC.CPML_Params(P.dz)
    assert C_P.alphaMax == 0.05000000074505806
    V = MC.Variables(P.Nz, P.timeSteps, 50, 10)
    assert V.alpha3 == 0.699999988079071
    P.epsRe = 2.2
    assert P.epsRe == float(np.float32(2.2))
    tupn = envDef.envSetup(9e9, 0.7, 7000, 8000, nonLinMed=True)
    Pn = MC.Params(*tupn, False, 0.7, 9e9, 20)
    assert (Pn.Nz, Pn.pmlWidth, Pn.timeSteps) == (11555, 2100, 19250)


def test_envsetup_guards_raise():
    with pytest.raises(ValueError):
        envDef.envSetup(9e9, 0.7, 30000, 30100)        # timeSteps too large
    with pytest.raises(ValueError):
        envDef.envSetup(9e9, 0.05, 400, 600)           # slab starts beyond the domain
    V, P, C_V, C_P = build_objects(dict(mode="free", freq=9e9, dom=0.2, win=[400, 600]))
    P.Nz = 30000
    with pytest.raises(ValueError):
        BaseFDTD11.FieldInit(V, P)


@pytest.mark.parametrize("name", ["free_sine_eps4", "free_gauss_eps4", "free_gauss_notfsf", "lorentz_sine",
                                  "lorentz_gauss", "lorentz_sine_6g", "nl_sine"])
def test_setup_chain_matches_reference(name):
    g = load_golden(name)
    V, P, C_V, C_P = build_objects(g["spec"])
    for k_mine, k_g in (("Nz", "Nz"), ("timeSteps", "timeSteps"), ("pmlWidth", "pmlWidth"), ("nzsrc", "nzsrc"),
                        ("materialFrontEdge", "mf"), ("materialRearEdge", "mr"), ("x1Loc", "x1Loc"),
                        ("x2Loc", "x2Loc"), ("dz", "dz"), ("delT", "delT"), ("courantNo", "courantNo")):
        assert getattr(P, k_mine) == g[k_g], k_mine
    mode = g["spec"]["mode"]
    BaseFDTD11.FieldInit(V, P)
    V.UpHyMat, V.UpExMat = BaseFDTD11.EmptySpaceCalc(V, P)
    if mode == "free":
        BaseFDTD11.Material(V, P)
        V.UpHyMat, V.UpExMat = BaseFDTD11.UpdateCoef(V, P)
    C_V = BaseFDTD11.CPML_FieldInit(V, P, C_V, C_P)
    C_V = SE.boundCondManager(V, P, C_V, C_P)
    for mine, theirs in ((C_V.beX, "beX"), (C_V.ceX, "ceX"), (C_V.bmY, "bmY"), (C_V.cmY, "cmY"), (C_V.Cb, "Cb"),
                         (C_V.C2, "C2"), (C_V.den_Exdz, "den_Exdz"), (C_V.den_Hydz, "den_Hydz"),
                         (V.UpExMat, "UpExMat"), (V.UpHyMat, "UpHyMat")):
        assert np.array_equal(mine, g[theirs]), theirs
    Exs, Hys = SE.SourceManager(V, P, C_V, C_P)
    if mode == "free":
        tauIn = 1 / (P.freq_in / 5)
        Exs = SE.Sig_Mod(V, P, Exs, tau=tauIn)
        Hys = SE.Sig_Mod(V, P, Hys, AmpMod=1 / P.CharImp, tau=tauIn)
    np.testing.assert_allclose(Exs, g["Exs"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(Hys, g["Hys"], rtol=1e-13, atol=0)
    if mode != "free":
        wp = V.plasmaFreqE
        for _ in range(2 if mode == "lorentz" else 1):
            wp = gStab.spatialStab(P.timeSteps, P.Nz, P.dz, P.freq_in, P.delT, wp, V.omega_0E, V.gammaE)[3]
        assert wp == pytest.approx(float(g["plasmaFreqE"]), rel=1e-14)
    # arrays produced by this chain are always in the tile engine's canonical form
    arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
    canon = dev.canonical_form(P, arrs, None)
    assert canon is not None
    assert canon[0] == V.UpExMat[0] and canon[1] == V.UpExMat[P.materialFrontEdge]


def test_canonical_form_rejects_general_arrays():
    V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=9e9, dom=0.2, win=[400, 600]))
    BaseFDTD11.FieldInit(V, P)
    BaseFDTD11.EmptySpaceCalc(V, P)
    C_V = BaseFDTD11.CPML_FieldInit(V, P, C_V, C_P)
    C_V = SE.boundCondManager(V, P, C_V, C_P)
    arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
    assert dev.canonical_form(P, arrs, None) is not None
    V.UpExMat[100] *= 1.0000001
    assert dev.canonical_form(P, BaseFDTD11._host_arrays(V, C_V, V.tempVarPol), None) is None
    V.UpExMat[100] = V.UpExMat[0]
    C_V.den_Hydz[5] = 0.5
    assert dev.canonical_form(P, BaseFDTD11._host_arrays(V, C_V, V.tempVarPol), None) is None
    C_V.den_Hydz[5] = 1.0
    assert dev.canonical_form(P, arrs, np.ones(len(V.Ex))) is not None     # a current slot does not change the coefficient form
    assert dev.probes_ok_for_tiles([10, 12, 40]) and not dev.probes_ok_for_tiles([10, 12, 14])


def test_kappa_not_one_denominators():
    """denominators() keeps the reference's mirrored indexing when kappaMax != 1."""
    import fdtd_oracle as fo
    V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=9e9, dom=0.2, win=[400, 600]))
    C_P.kappaMax = 3.0
    BaseFDTD11.FieldInit(V, P)
    BaseFDTD11.EmptySpaceCalc(V, P)
    C_V = BaseFDTD11.CPML_FieldInit(V, P, C_V, C_P)
    C_V = SE.boundCondManager(V, P, C_V, C_P)
    k = fo.cpml_coefficients(P.Nz + 1, P.pmlWidth, P.dz, P.delT, V.UpExMat, V.UpExHcompsCo, kappaMax=3.0)
    assert np.array_equal(C_V.den_Exdz, k["denE"]) and np.array_equal(C_V.den_Hydz, k["denH"])
    assert np.array_equal(C_V.beX, k["beX"]) and np.array_equal(C_V.cmY, k["cmY"])
    assert dev.canonical_form(P, BaseFDTD11._host_arrays(V, C_V, V.tempVarPol), None) is None


def test_library_exports_every_header_symbol():
    """The C-ABI library loads and exports exactly what include/pyfdtd_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "pyfdtd_b200.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    lib = nat.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pf_abi_version() == 2


def test_ctypes_structs_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pyfdtd_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(PfGrid), offsetof(PfGrid, Ex), '
                   'offsetof(PfGrid, probe_out), sizeof(PfPic), offsetof(PfPic, z), offsetof(PfPic, Jx), '
                   'sizeof(PfSetupMember), offsetof(PfSetupMember, off_srcH), sizeof(PfDormant), offsetof(PfDormant, JxKerr), '
                   'sizeof(PfDrudeJ), offsetof(PfDrudeJ, Hys));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    want = [ctypes.sizeof(nat.PfGrid), nat.PfGrid.Ex.offset, nat.PfGrid.probe_out.offset,
            ctypes.sizeof(nat.PfPic), nat.PfPic.z.offset, nat.PfPic.Jx.offset,
            ctypes.sizeof(nat.PfSetupMember), nat.PfSetupMember.off_srcH.offset, ctypes.sizeof(nat.PfDormant),
            nat.PfDormant.JxKerr.offset, ctypes.sizeof(nat.PfDrudeJ), nat.PfDrudeJ.Hys.offset]
    assert got == want


def test_hot_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    V, P, C_V, C_P = build_objects(dict(mode="free", freq=9e9, dom=0.2, win=[400, 600]))
    with pytest.raises(nat.NativeError):
        MC.Controller(V, P, C_V, C_P)
    BaseFDTD11.FieldInit(V, P)
    with pytest.raises(nat.NativeError):
        BaseFDTD11.ADE_ExUpdate(V, P, C_V, C_P, 0)


def test_native_libm_maps_equal_cpython_math():
    """pf_host_exp / pf_host_pow (setup-chain helpers in the product library) are the same glibc calls as
    math.exp / math.pow: bit-identical, including the memoised second call."""
    import math
    from pyfdtd_b200 import BaseFDTD11 as B
    rng = np.random.default_rng(5)
    x = -np.abs(rng.normal(size=3000)) * 30
    B._LIBM_CACHE.clear()
    for _ in range(2):
        got = B._libm_map("exp", x)
        assert np.array_equal(got, np.array([math.exp(v) for v in x]))
    d = rng.uniform(0, 1, 3000)
    for e in (4.0, 1.0, 2.5):
        assert np.array_equal(B._libm_map("pow", d, e), np.array([math.pow(v, e) for v in d]))


def test_batched_ref_tester_matches_reference_ref_tester():
    """TransformHandler.ref_tester_batch (the device-side post-processing of sweeps, run here on CPU tensors)
    against the RefTester mirror on traces of the unmodified reference; read windows; DC-peak error."""
    import torch
    from types import SimpleNamespace
    from pyfdtd_b200 import TransformHandler as TH
    from conftest import load_golden
    g = load_golden("lorentz_sine")
    be, af = np.asarray(g["x1ColBe"]), np.asarray(g["x1ColAf"])
    T = len(be)
    P = SimpleNamespace(timeSteps=T, delT=1e-13)
    want = [TH.RefTester(None, P, y, 1)[2] for y in (be, af)]
    pad = np.zeros((2, T + 7))                       # rows are padded in the device pool: only [:T] may be used
    pad[0, :T], pad[1, :T] = be, af
    pad[:, T:] = 123.0
    val, idx = TH.ref_tester_batch(torch.from_numpy(pad), T)
    np.testing.assert_allclose(val.numpy(), want, rtol=1e-12)
    assert (idx.numpy() > 0).all()
    # windows: zero outside [keep_from, keep_to] exactly like the integrators do (Solver_Engine.py:360-368)
    rng = np.random.default_rng(3)
    raw = rng.normal(size=(3, T)) + np.sin(0.05 * np.arange(T))[None, :]
    kf, kt = np.array([0, 10, 40]), np.array([T - 1, T - 30, int(T * 0.7)])
    val, _ = TH.ref_tester_batch(torch.from_numpy(raw), T, keep_from=kf, keep_to=kt)
    n = np.arange(T)
    for m in range(3):
        y = np.where((n >= kf[m]) & (n <= kt[m]), raw[m], 0.0)
        assert val[m].item() == pytest.approx(TH.RefTester(None, P, y, 1)[2], rel=1e-12)
    with pytest.raises(ValueError):
        TH.ref_tester_batch(torch.ones((1, T), dtype=torch.float64), T)


def test_pic_oracle_fused_summation_tree_agrees_with_the_plain_one():
    """oracle/pic_oracle.py: deposit_fused (the tree of pf_pic_step_sorted) and deposit (pf_pic_deposit) sum the same
    shares in different orders -- equal to rounding, and exactly charge-conserving in the same sense."""
    import pic_oracle as po
    L, dz, dt = 257, 8.3e-5, 2.6e-13
    n = 40_000
    z, ux, uz, w, cell = po.make_beam(n, L, dz, seed=3, thermal=0.3)
    rng = np.random.default_rng(0)
    Ex, Hy = 2e5 * rng.standard_normal(L), 5e2 * rng.standard_normal(L)
    zo, uxo, uzo, wo, co = po.sort_by_cell(z, ux, uz, w, cell)
    zp, uxp, uzp, cp = po.push(zo, uxo, uzo, Ex, Hy, dz=dz, dt=dt, q_over_m=-1.75882001076e11, c=299792458.0,
                               mu0=1.25663706127e-06)
    assert np.max(np.abs(cp.astype(np.int64) - co)) <= 1
    for S in (1, 3, 8):
        Jf = po.deposit_fused(zp, uxp, uzp, wo, co, cp, L, S, dz=dz, c=299792458.0, jx_scale=-1.6e-19)
        Jp = po.deposit(*po.sort_by_cell(zp, uxp, uzp, wo, cp), L, dz=dz, c=299792458.0, jx_scale=-1.6e-19)
        assert np.max(np.abs(Jf - Jp)) <= 1e-13 * np.max(np.abs(Jp))
    assert po.sub_warps(20_000_000, 13194) == 6 and po.sub_warps(1000, 4097) == 1 and po.sub_warps(100_000_000, 13194) == 8


def test_python_constants_match_header_enums(tmp_path):
    """Modes, engines and flags of _native.py are the header's enum values (compiled from the header itself)."""
    names = ["PF_FREE", "PF_LORENTZ", "PF_NL", "PF_LORENTZ_NL", "PF_ENGINE_OPS", "PF_ENGINE_TILE", "PF_F_TFSF",
             "PF_F_CPML_M", "PF_F_CPML_P", "PF_F_CANONICAL", "PF_F_FMA", "PF_F_FP32", "PF_F_NEWTON",
             "PF_BLOCK_F_TABLES_VALID", "PF_BLOCK_F_SWAPPED", "PF_BLOCK_F_EDGE_TILES", "PF_BLOCK_F_INNER_TILES"]
    src = tmp_path / "en.c"
    src.write_text('#include <stdio.h>\n#include "pyfdtd_b200.h"\nint main(){printf("' + " ".join(["%d"] * len(names)) +
                   '\\n", ' + ", ".join(names) + ");return 0;}\n")
    exe = tmp_path / "en"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    assert got == [getattr(nat, n) for n in names]
    from pyfdtd_b200 import _device as dev
    assert dev.MODE_ID == {"free": nat.PF_FREE, "lorentz": nat.PF_LORENTZ, "nl": nat.PF_NL, "lorentz_nl": nat.PF_LORENTZ_NL}


def test_kerr_lorentz_scalars_and_flags():
    """Host-side description of the PF_LORENTZ_NL composition and of the optional arithmetic flags."""
    V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=9e9, dom=0.15, win=[300, 320]))
    lin = BaseFDTD11.grid_scalars(V, P)
    ker = BaseFDTD11.grid_scalars(V, P, kerr_lorentz=True)
    chi3 = float(V.chi3Stat)
    assert (ker["cub_a"], ker["cub_b"], ker["cub_c"]) == (chi3 ** 2, 2 * chi3, 1.0)
    assert ker["nl_den0"] == P.permit_0 and ker["nl_den1"] == P.permit_0 * chi3
    assert all(ker[k] == lin[k] for k in ("polA", "polB", "polC", "dt_over_dz", "eps0", "mf", "mr"))
    base = BaseFDTD11.grid_flags(P)
    assert BaseFDTD11.grid_flags(P, fp32=True) == base | nat.PF_F_FP32
    assert BaseFDTD11.grid_flags(P, newton=True) == base | nat.PF_F_NEWTON
    assert BaseFDTD11.grid_flags(P, fma=True, fp32=True, newton=True) == base | nat.PF_F_FMA | nat.PF_F_FP32 | nat.PF_F_NEWTON
