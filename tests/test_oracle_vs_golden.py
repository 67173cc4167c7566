"""The CPU oracle against the outputs of the UNMODIFIED reference (tests/golden, made by
oracle/make_golden.py).  This is what pins the oracle (the reference's own tests hold
no vectors).  Bitwise equality is expected when the host libm / numpy SIMD dispatch match the machine
that generated the goldens; the hard bound is BASELINE's 1e-10 relative."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import fdtd_oracle as fo
from conftest import load_golden

SINGLE = ["free_sine_eps4", "free_gauss_eps4", "free_gauss_notfsf", "lorentz_sine", "lorentz_gauss",
          "lorentz_sine_6g", "nl_sine", "nl_sine_amp", "lorentz_default_full", "free_default_full"]
RTOL = 1e-10


def rel_err(got, want):
    scale = np.max(np.abs(want))
    if scale == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want))) / scale)


def case_from_golden(g):
    s = g["spec"]
    c = fo.make_case(s["mode"], s["freq"], s["dom"], *s["win"], source=s.get("source", "sine"),
                     tfsf=s.get("tfsf", True), periods=s.get("periods", 1000.0), epsRe=s.get("epsRe", 1.0),
                     amplitude=s.get("amplitude", 1.0))
    return c


@pytest.mark.parametrize("name", SINGLE)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    v = g["versions"]
    if (v["epsilon_0"], v["mu_0"]) != (fo.EPS0, fo.MU0):
        pytest.skip("scipy.constants differ from the ones the goldens were generated with")
    c = case_from_golden(g)
    assert (c.Nz, c.T, c.pw, c.nzsrc, c.mf, c.mr, c.x1Loc, c.x2Loc) == tuple(
        int(g[k]) for k in ("Nz", "timeSteps", "pmlWidth", "nzsrc", "mf", "mr", "x1Loc", "x2Loc"))
    assert c.dz == float(g["dz"]) and c.dt == float(g["delT"]) and c.courantNo == float(g["courantNo"])
    out = fo.run_case(c, snapshots=True)
    pairs = [("Ex", "Ex"), ("Hy", "Hy"), ("psi_Ex", "psi_Ex"), ("psi_Hy", "psi_Hy"), ("x1ColBe", "x1ColBe"),
             ("x1ColAf", "x1ColAf"), ("Exs", "Exs"), ("Hys", "Hys")]
    if c.mode == "lorentz":
        pairs += [("P", "polarisationCurr"), ("Dx", "Dx")]
    if c.mode == "nl":
        pairs += [("Port1", "Port1"), ("Port2", "Port2"), ("Dx", "Dx"), ("Acubic", "Acubic")]
    exact = True
    for mine, theirs in pairs:
        assert rel_err(out[mine], g[theirs]) <= RTOL, (name, mine)
        exact &= np.array_equal(out[mine], g[theirs])
    assert out["plasmaFreqE"] == pytest.approx(float(g["plasmaFreqE"]), rel=1e-14)
    if "beX" in g:
        for mine, theirs in (("beX", "beX"), ("ceX", "ceX"), ("cmY", "cmY"), ("Cb", "Cb"), ("C2", "C2"),
                             ("denE", "den_Exdz"), ("denH", "den_Hydz")):
            assert rel_err(out["coef"][mine], g[theirs]) <= 1e-14, (name, mine)
        step = max(1, len(out["Ex_History"]) // 4)
        assert rel_err(out["Ex_History"][::step], g["Ex_History_rows"]) <= RTOL
    if c.mode == "lorentz" and np.isfinite(g["reflection"]):
        assert fo.reflection(out, c) == pytest.approx(float(g["reflection"]), rel=1e-9)
        m = c.medium
        assert fo.analytical_reflection(c.freq, out["plasmaFreqE"], m["w0"], m["gam"]) == pytest.approx(
            float(g["analytical_reflection"]), rel=1e-12)
    print(f"{name}: oracle {'bitwise ==' if exact else 'within 1e-10 of'} reference")


def test_survey_golden_scalars():
    """SURVEY.md 8c survey-time scalars of the default geometry, re-derived from the golden files."""
    lor = load_golden("lorentz_default_full")
    free = load_golden("free_default_full")
    assert float(lor["sumEx"]) == pytest.approx(-94.55240287165128, rel=1e-12)
    assert float(lor["maxAbsEx"]) == pytest.approx(1.3231301386834344, rel=1e-12)
    assert float(free["sumEx"]) == pytest.approx(738.6155800460517, rel=1e-12)
    assert int(lor["Nz"]) == 13193 and int(lor["timeSteps"]) == 23997 and int(lor["pmlWidth"]) == 2394


def test_cubic_root_known_answers():
    g = load_golden("cubic_roots")
    lib = fo.lib()
    co, want = g["coeffs"], g["root0"]
    got = np.array([lib.orc_cubic_root0(*map(float, row)) for row in co])
    scale = np.maximum(np.abs(want.real), 1e-300)
    # root[0] is real in every branch except the complex quadratic fallback (real part compared)
    assert np.max(np.abs(got - want.real) / scale) <= 1e-12
    nl_family = co[:, 3] < 0
    assert np.array_equal(got[: 8 * 43], want.real[: 8 * 43]), "NL-path polynomials must match bit for bit"
    assert nl_family.any()


def test_sweep_reflection_matches_reference():
    """LoopedSim(loop=True) golden: 20 members, each derived from the previous member's adjusted wp."""
    g = load_golden("lorentz_sweep")
    s = g["spec"]
    med = fo.default_medium()
    wp_prev = med["wp"]
    freq = s["freq0"]
    for i, mem in enumerate(g["members"]):
        eps = fo.lorentz_eps(wp_prev, med["w0"], med["gam"], freq)
        c = fo.make_case("lorentz", freq, s["dom"], *s["win"], source="sine", periods=1.0, eps_for_nlam=eps)
        assert (c.Nz, c.T) == (mem["Nz"], mem["T"]), i
        if i in (0, 7, 19):   # three members are stepped in full; the rest check the setup chain only
            out = fo.run_case(c)
            assert fo.reflection(out, c) == pytest.approx(float(g["measured"][i]), rel=1e-9)
            wp_prev = out["plasmaFreqE"]
        else:
            wp1, _ = fo.spatial_stab(c.Nz, c.dz, c.freq, c.dt, med["wp"], med["w0"], med["gam"])
            wp_prev, _ = fo.spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp1, med["w0"], med["gam"])
        assert wp_prev == pytest.approx(mem["wp_after"], rel=1e-14)
        assert fo.analytical_reflection(freq, wp_prev, med["w0"], med["gam"]) == pytest.approx(
            float(g["analytical"][i]), rel=1e-12)
        freq = freq + s["interval"]


def test_kerr_lorentz_composition_reduces_to_the_reference_lorentz_path():
    """Mode "lorentz_nl" is builder-defined (the reference's nonlinear loop has no dispersion ADE).  Its
    linear limit chi3 = 0 must be the reference's Lorentz integrator bit for bit (golden), and a non-zero chi3
    must change the result at the expected order (chi3 |E|^2)."""
    import fdtd_oracle as fo
    from conftest import load_golden
    g = load_golden("lorentz_gauss")
    s = g["spec"]
    kw = dict(source=s.get("source", "sine"), tfsf=s.get("tfsf", True), periods=s.get("periods", 1000.0),
              epsRe=s.get("epsRe", 1.0), amplitude=s.get("amplitude", 1.0))
    c = fo.make_case("lorentz_nl", s["freq"], s["dom"], *s["win"], **kw)
    c.medium = dict(c.medium, chi3=0.0)
    out = fo.run_case(c)
    for nm, gold in (("Ex", "Ex"), ("Hy", "Hy"), ("P", "polarisationCurr"), ("x1ColAf", "x1ColAf")):
        assert np.array_equal(out[nm], g[gold]), nm
    c = fo.make_case("lorentz_nl", s["freq"], s["dom"], *s["win"], **kw)
    out = fo.run_case(c)
    peak = np.max(np.abs(g["x1ColAf"]))
    diff = np.max(np.abs(out["x1ColAf"] - g["x1ColAf"])) / peak
    a2 = float(np.max(out["Acubic"]))
    assert a2 > 0 and 1e-3 * c.medium["chi3"] * a2 < diff < 10 * c.medium["chi3"] * a2


def test_drude_limit_of_the_lorentz_ade_behaves_like_a_plasma():
    """BASELINE names Lorentz/Drude terms; the reference steps a Drude medium only in a scratch script
    (TESTBOXDIPSERSE.py:79-94) and its setup chain divides by omega_0^2 (genericStability.py:30).  The Drude
    medium is the omega_0 = 0 limit of the same ADE (P'' + gamma P' = eps0 wp^2 E): overdense (f < fp) it reflects
    almost everything, underdense little, and less at higher frequency -- steady-state amplitude ratios against the
    complex Fresnel coefficient of eps = 1 - wp^2/(w^2 + i gamma w) (the scheme itself is only ~25 % accurate on
    reflection: the reference's own Lorentz sweep measures 0.18 against an analytical 0.26)."""
    import fdtd_oracle as fo
    got = {}
    for f0, fp in ((9e9, 5e9), (7e9, 5e9), (9e9, 12e9)):
        c = fo.make_case("lorentz", f0, 0.2, 2000, 2200, source="sine", periods=1000.0)
        c.medium = dict(c.medium, w0=0.0, wp=2 * np.pi * fp, gam=2 * np.pi * 0.2e9)
        out = fo.run_case(c)
        assert np.isfinite(out["Ex"]).all() and np.max(np.abs(out["P"])) > 0
        n = c.T
        a_inc = np.max(np.abs(out["x1ColBe"][int(0.5 * n):int(0.7 * n)]))
        a_ref = np.max(np.abs(out["x1ColAf"][int(0.75 * n):]))
        w = 2 * np.pi * f0
        nn = np.sqrt(1 - out["plasmaFreqE"] ** 2 / (w * w + 1j * c.medium["gam"] * w))
        got[(f0, fp)] = (a_ref / a_inc, abs((1 - nn) / (1 + nn)))
    for (f0, fp), (meas, fresnel) in got.items():
        assert abs(meas - fresnel) <= 0.05, (f0, fp, meas, fresnel)
    assert got[(9e9, 12e9)][0] > 0.95                       # overdense: mirror
    assert got[(9e9, 5e9)][0] < got[(7e9, 5e9)][0] < 0.25     # underdense: weak, falling with frequency


def test_oracle_newton_root_and_the_ill_conditioning_of_the_closed_form_on_kerr_coefficients():
    """oracle/fdtd_oracle.c: orc_cubic_root_newton (the root PF_LORENTZ_NL is specified with) is the positive root of
    the polynomial to rounding level, while the reference's closed form, on the same Kerr coefficients
    (chi3^2, 2 chi3, 1), runs in its three-real-root branch and misses the root by orders of magnitude more --
    which is why the composition is not specified through it."""
    import ctypes
    import fdtd_oracle as fo
    lib = fo.lib()
    lib.orc_cubic_root_newton.argtypes = [ctypes.c_double] * 4
    lib.orc_cubic_root_newton.restype = ctypes.c_double
    chi3 = 1e-3
    a, b, c = chi3 ** 2, 2 * chi3, 1.0
    rng = np.random.default_rng(5)
    worst_newton, worst_closed = 0.0, 0.0
    for q2 in 10.0 ** rng.uniform(-6, 4, 400):
        xn = lib.orc_cubic_root_newton(a, b, c, q2)
        xc = lib.orc_cubic_root0(a, b, c, -q2)
        exact = q2 / (1 + chi3 * xn) ** 2            # fixed point of A (1 + chi3 A)^2 = q2, evaluated at the Newton root
        worst_newton = max(worst_newton, abs(xn - exact) / xn)
        resid_c = abs(((a * xc + b) * xc + c) * xc - q2) / q2
        worst_closed = max(worst_closed, resid_c)
        assert abs(((a * xn + b) * xn + c) * xn - q2) / q2 <= 4e-16 * (1 + 3 * chi3 * xn) + 1e-15
    assert worst_newton <= 1e-14
    assert worst_closed > 1e-12                         # the closed form is measurably off on these coefficients
    # admissible random polynomials: Newton == largest real root of numpy.roots
    for _ in range(100):
        a, b, c = 10.0 ** rng.uniform(-8, -2), 10.0 ** rng.uniform(-5, -1), 10.0 ** rng.uniform(-1, 1)
        q2 = 10.0 ** rng.uniform(-6, 3)
        r = np.roots([a, b, c, -q2])
        want = float(np.max(r[np.abs(r.imag) < 1e-9 * np.abs(r.real).max()].real))
        assert lib.orc_cubic_root_newton(a, b, c, q2) == pytest.approx(want, rel=1e-10)


def test_pic_oracle_is_pinned_to_known_physics():
    """The PIC model has no reference implementation, so its CPU definition (oracle/pic_oracle.py) is pinned
    to known answers instead: free drift, the exact momentum gain in a uniform E field, conservation of |u| and the
    relativistic gyro-frequency in a uniform B field, specular walls, and charge-current conservation of the CIC deposit
    (both summation trees)."""
    import pic_oracle as po
    C0, MU0, QM, Q = 299792458.0, 1.25663706127e-06, -1.75882001076e11, -1.602176634e-19
    L, dz, dt = 2049, 1e-4, 2e-13
    kw = dict(dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
    rng = np.random.default_rng(1)
    n = 1000
    z = rng.uniform(0.2, 0.8, n) * (L - 1) * dz
    ux, uz = 1e7 * rng.standard_normal(n), 2e8 * (1 + 0.01 * rng.standard_normal(n))
    zero = np.zeros(L)
    # (a) no fields: z += vz dt, momenta untouched
    z1, ux1, uz1, c1 = po.push(z, ux, uz, zero, zero, **kw)
    g = np.sqrt(1 + (ux ** 2 + uz ** 2) / C0 ** 2)
    assert np.array_equal(ux1, ux) and np.array_equal(uz1, uz)
    np.testing.assert_allclose(z1, z + uz / g * dt, rtol=1e-15)
    assert np.array_equal(c1, np.floor(z1 / dz).astype(np.int32))
    # (b) uniform Ex: two half kicks add up to q/m E dt exactly (to rounding), uz unchanged
    E0 = 3e5
    _, ux2, uz2, _ = po.push(z, ux, uz, np.full(L, E0), zero, **kw)
    np.testing.assert_allclose(ux2 - ux, QM * E0 * dt, rtol=1e-9)
    assert np.array_equal(uz2, uz)
    # (c) uniform By: |u| conserved, rotation angle per step = 2 atan(q B dt / (2 gamma m))
    B0 = 0.5
    H0 = np.full(L, B0 / MU0)
    _, ux3, uz3, _ = po.push(z, ux, uz, zero, H0, **kw)
    np.testing.assert_allclose(np.hypot(ux3, uz3), np.hypot(ux, uz), rtol=1e-14)
    ang = np.arctan2(ux3, uz3) - np.arctan2(ux, uz)
    want = 2 * np.arctan(QM * dt * 0.5 * B0 / g)
    np.testing.assert_allclose(np.abs(ang), np.abs(want), rtol=1e-9)
    # (d) specular walls keep particles inside and flip uz
    zmax = (L - 1) * dz
    zw, _, uzw, _ = po.push(np.array([1e-9, zmax - 1e-9]), np.zeros(2), np.array([-2e8, 2e8]), zero, zero, **kw)
    assert np.all((zw >= 0) & (zw <= zmax)) and uzw[0] > 0 and uzw[1] < 0
    # (e) deposit: total current = sum q w vx for both summation trees, each particle's share goes to its two nodes
    w = rng.uniform(0.5, 2.0, n) * 1e10
    cell = np.clip(np.floor(z / dz).astype(np.int64), 0, L - 2).astype(np.int32)
    zs, uxs, uzs, ws, cs = po.sort_by_cell(z, ux, uz, w, cell)
    J = po.deposit(zs, uxs, uzs, ws, cs, L, dz=dz, c=C0, jx_scale=Q)
    gs = np.sqrt(1 + (uxs ** 2 + uzs ** 2) / C0 ** 2)
    assert np.sum(J) == pytest.approx(Q * np.sum(ws * uxs / gs), rel=1e-12)
    Jf = po.deposit_fused(zs, uxs, uzs, ws, cs, cs, L, 2, dz=dz, c=C0, jx_scale=Q)
    assert np.sum(Jf) == pytest.approx(np.sum(J), rel=1e-12)
    one = po.deposit(np.array([10.25 * dz]), np.array([1e7]), np.array([0.0]), np.array([1.0]), np.array([10], dtype=np.int32),
                     L, dz=dz, c=C0, jx_scale=1.0)
    vx = 1e7 / np.sqrt(1 + (1e7 / C0) ** 2)
    assert one[10] == pytest.approx(0.75 * vx, rel=1e-12) and one[11] == pytest.approx(0.25 * vx, rel=1e-12)
    assert np.count_nonzero(one) == 2


def test_kerr_lorentz_oracle_state_satisfies_its_constitutive_relation():
    """Mode "lorentz_nl" end state: inside the slab Acubic = |Ex|^2 and Dx - P = eps0 (eps_inf + chi3 |Ex|^2) Ex to rounding
    level -- the composition's definition holds for the fields the oracle returns (strongly nonlinear drive)."""
    import fdtd_oracle as fo
    c = fo.make_case("lorentz_nl", 9e9, 0.15, 300, 320, source="gauss", amplitude=4.0)
    out = fo.run_case(c)
    sl = slice(c.mf, c.mr)
    Ex, Dn, A = out["Ex"][sl], (out["Dx"] - out["P"])[sl], out["Acubic"][sl]
    chi3 = c.medium["chi3"]
    live = A > 0
    assert live.sum() > 100 and np.max(A) > 10.0                      # chi3 |E|^2 up to a few per cent and beyond
    np.testing.assert_allclose(A[live], Ex[live] ** 2, rtol=1e-12)
    np.testing.assert_allclose(Dn[live], fo.EPS0 * (fo.KERR_EPS_INF + chi3 * Ex[live] ** 2) * Ex[live], rtol=1e-12)
    dead = ~live                                                       # below the reference's 1e-8 threshold: linear law
    np.testing.assert_allclose(Dn[dead], fo.EPS0 * Ex[dead], rtol=1e-12, atol=1e-30)


def test_builder_defined_models_have_not_drifted():
    """tests/golden/builder_*.npz (oracle/make_builder_golden.py): regression pins of the models the reference does not
    contain -- Kerr-Lorentz composition, Drude limit, PIC step.  Not reference parity: they hold the DEFINITIONS still."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_builder_golden", os.path.join(ROOT, "oracle", "make_builder_golden.py"))
    mbg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mbg)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        mbg.OUT = tmp
        mbg.kerr_lorentz(); mbg.drude(); mbg.pic()
        for name in ("builder_kerr_lorentz", "builder_drude", "builder_pic"):
            new = np.load(os.path.join(tmp, name + ".npz"))
            old = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
            assert sorted(new.files) == sorted(old.files)
            for k in old.files:
                a, b = np.asarray(new[k], dtype=np.float64), np.asarray(old[k], dtype=np.float64)
                scale = max(float(np.max(np.abs(b))), 1e-300)
                assert float(np.max(np.abs(a - b))) <= 1e-11 * scale, (name, k)


# ------------------------------------------------------------------------------------------------ PIC: external physics
def _cold_plasma_run(wp_over_w, steps, record_from):
    """The PIC model of oracle/pic_oracle.py coupled to the oracle's vacuum integrator through the Jx slot, driven by the
    TF/SF sine source: a half-space of cold electrons (at rest, 4 per cell, quiet start) from the slab front into the right
    CPML.  Returns (Ex samples [steps - record_from, L], case, first plasma cell)."""
    import ctypes
    import pic_oracle as po
    C0, MU0, QM, Q = 299792458.0, 1.25663706127e-06, -1.75882001076e11, -1.602176634e-19
    c = fo.make_case("free", 9e9, 0.15, 1500, 1600, source="sine", periods=1000, epsRe=1.0)
    assert c.T >= steps
    L, dz, dt = c.L, c.dz, c.dt
    # continuous sine through the TF/SF point (SmoothTurnOn's tables without the free-space run's sech envelope),
    # switched on over three periods
    nn = np.arange(c.T)
    ppw = C0 / (c.freq * dz)
    ramp = np.minimum(1.0, nn * dt * c.freq / 3.0)
    Exs = ramp * np.sin(2.0 * np.pi / ppw * (c.courantNo * nn)) * c.courantNo
    Hys = ramp * np.sin(2.0 * np.pi / ppw * (c.courantNo * (nn + 1))) * c.courantNo * (1 / fo.CHAR_IMP)
    pa = fo.PassArrays(c, 0.0, Exs, Hys, [], False, Jx=np.zeros(L))
    a, b = c.mf, L - c.pw // 2
    ppc = 4
    cells = np.arange(a, b)
    z = ((cells[:, None] + (np.arange(ppc)[None, :] + 0.5) / ppc) * dz).ravel()
    w_src = 2 * np.pi * c.freq
    n_e = (wp_over_w * w_src) ** 2 * fo.EPS0 * (Q / QM) / Q ** 2            # wp^2 = n q^2 / (eps0 m)
    wgt = np.full(len(z), n_e * dz / ppc)                                   # slot = J dz  (pic.py: weights per unit area)
    ux, uz = np.zeros(len(z)), np.zeros(len(z))
    cell = np.clip(np.floor(z / dz).astype(np.int64), 0, L - 2).astype(np.int32)

    def deposit(z, ux, uz, cell):           # pic_oracle.deposit's CIC definition, summed by bincount (order-free)
        g = np.sqrt(1.0 + (ux * ux + uz * uz) / (C0 * C0))
        wv = wgt * (ux / g)
        f = z * (1.0 / dz) - cell
        return Q * (np.bincount(cell, wv * (1.0 - f), L) + np.bincount(cell + 1, wv * f, L))
    uxt = 1e3 * np.sin(np.arange(len(z)))   # the bincount deposit IS the oracle's deposit up to summation order
    ref = po.deposit(z, uxt, uz, wgt, cell, L, dz=dz, c=C0, jx_scale=Q)      # (z is generated cell by cell: already sorted)
    assert np.allclose(deposit(z, uxt, uz, cell), ref, rtol=0, atol=1e-14 * np.max(np.abs(ref)))
    rec = np.zeros((steps - record_from, L))
    for n in range(steps):
        pa.Jx[:] = deposit(z, ux, uz, cell)
        fo.lib().orc_run(ctypes.byref(pa.g), fo.MODE_ID["free"], 0, n, 1, c.T)
        z, ux, uz, cell = po.push(z, ux, uz, pa.Ex, pa.Hy, dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
        if n >= record_from:
            rec[n - record_from] = pa.Ex
    return rec, c, a


def test_pic_model_reproduces_cold_plasma_dispersion():
    """External cross-check of the builder-defined PIC model (push + CIC gather + CIC deposit + the Jx coupling): an
    electromagnetic wave entering a cold electron plasma must obey the textbook dispersion relation
    w^2 = wp^2 + c^2 k^2 -- not a property any single piece of the model has by construction.
    Underdense (wp = 0.6 w): the wave number inside the plasma is k = (w/c) sqrt(1 - wp^2/w^2) = 0.8 w/c.
    Overdense (wp = 1.5 w): the field is evanescent with decay constant kappa = (w/c) sqrt(wp^2/w^2 - 1)."""
    steps, rec0 = 4200, 3000
    # --- underdense: phase advance between two cells inside the plasma (forward wave only: half-space, absorbed in the CPML)
    rec, c, a = _cold_plasma_run(0.6, steps, rec0)
    w = 2 * np.pi * c.freq
    t = (np.arange(rec0, steps) + 1) * c.dt
    demod = np.exp(-1j * w * t) @ rec                    # complex amplitude of exp(+i w t) in every cell
    phase = np.unwrap(np.angle(demod[a + 100: a + 600])) # forward wave exp(i(wt - kz)): the phase falls linearly with z
    k_meas = -np.polyfit(np.arange(500) * c.dz, phase, 1)[0]
    k_theory = (w / 299792458.0) * np.sqrt(1 - 0.36)
    assert abs(k_meas / k_theory - 1) < 0.01, (k_meas, k_theory)
    assert abs(k_meas / (w / 299792458.0) - 1) > 0.15    # ... and it is NOT the vacuum wave number
    # --- overdense: exponential decay inside the plasma
    rec, c, a = _cold_plasma_run(1.5, steps, rec0)
    demod = np.abs(np.exp(-1j * w * t) @ rec)
    q1, q2 = a + 40, a + 140
    kappa_meas = np.log(demod[q1] / demod[q2]) / ((q2 - q1) * c.dz)
    kappa_theory = (w / 299792458.0) * np.sqrt(2.25 - 1)
    assert abs(kappa_meas / kappa_theory - 1) < 0.01, (kappa_meas, kappa_theory)


def test_kerr_lorentz_composition_has_the_textbook_nonlinear_index():
    """External cross-check of the builder-defined PF_LORENTZ_NL composition (Lorentz ADE + instantaneous Kerr law on
    Dx - P): for a monochromatic field E0 cos(wt) the Kerr polarisation eps0 chi3 E^3 has the component
    (3/4) eps0 chi3 E0^2 at w, i.e. the medium's effective permittivity is eps_L(w) + (3/4) chi3 |E|^2 -- the standard n2
    result, which nothing in the cell-wise law states.  Measured: the extra phase lag a strong wave accumulates along the
    slab relative to a weak one, against (w/c)^2 (3/4) chi3 / (2 k) * integral |E(z)|^2 dz with the measured amplitude."""
    import ctypes
    c = fo.make_case("lorentz_nl", 9e9, 0.3, 2500, 2600, source="sine", periods=1000)
    wp = c.medium["wp"]
    for _ in range(2):
        wp, _ = fo.spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp, c.medium["w0"], c.medium["gam"])
    Exs, Hys = fo.sources(c)
    probes = list(range(c.mf + 50, c.mf + 1900, 50))
    assert probes[-1] < c.L - c.pw - 100
    w = 2 * np.pi * c.freq
    n0, N = c.T - 2600, 2527                                  # six periods (421.05 steps each), after the transient
    win = np.hanning(N)
    tt = (np.arange(n0, n0 + N) + 1) * c.dt
    amps = {}
    for amp in (0.01, 8.0):
        pa = fo.PassArrays(c, wp, Exs * amp, Hys * amp, probes, False)
        fo.lib().orc_run(ctypes.byref(pa.g), fo.MODE_ID["lorentz_nl"], 1, 0, c.T, c.T)
        amps[amp] = (pa.probe_out[:, n0:n0 + N] * win) @ np.exp(-1j * w * tt) * (2.0 / win.sum())
    lo, hi = amps[0.01] / 0.01, amps[8.0]
    z = np.asarray(probes) * c.dz
    k_lo = -np.polyfit(z, np.unwrap(np.angle(lo)), 1)[0]
    assert 1.5 < k_lo / (w / 299792458.0) < 1.9                # Re n of the Lorentz medium at 9 GHz: ~1.70
    E2 = np.abs(hi) ** 2
    assert 25.0 < E2[0] < 60.0 and E2[-1] < 0.6 * E2[0]        # chi3 |E|^2 ~ 4 %, decaying with the linear absorption
    dphi = np.unwrap(np.angle(hi / (lo * 8.0)))
    lag = -(dphi - dphi[0])                                    # extra phase lag of the strong wave, from the first probe on
    chi3 = c.medium["chi3"]
    dk = (w / 299792458.0) ** 2 * 0.75 * chi3 * E2 / (2.0 * k_lo)
    theory = np.concatenate([[0.0], np.cumsum(0.5 * (dk[1:] + dk[:-1]) * np.diff(z))])
    assert lag[-1] > 0.05
    assert abs(lag[-1] / theory[-1] - 1) < 0.02, (lag[-1], theory[-1])          # measured here: 0.9989
    assert np.max(np.abs(lag - theory)) < 0.05 * theory[-1]


# ------------------------------------------------------------------------------------------------ dormant models
def _dormant_golden():
    import ast
    g = np.load(os.path.join(ROOT, "tests", "golden", "dormant_leaf_ops.npz"), allow_pickle=False)
    return g, ast.literal_eval(str(g["scalars"]))


def test_oracle_reproduces_the_reference_dormant_leaf_functions():
    """Varin Kerr + Raman ADE, Kerr current and Mur ABC (BaseFDTD11.py:567-609, 762-788): the C restatement against the
    outputs of the unmodified reference functions (oracle/make_dormant_golden.py), three chained rounds, bit for bit."""
    g, k = _dormant_golden()
    Ex, Eold = g["in_Ex"].copy(), g["in_tempTempVarE"].copy()
    Q, G, J, Pol, Pbar = (g[f"in_{n}"].copy() for n in ("Qx3", "Gx3", "Jx", "polarisationCurr", "Pbar3"))
    for r in range(3):
        fo.varin_pbar(k["mf"], k["mr"], k["permit_0"], k["chi1Stat"], k["chi3Stat"], k["alpha3"], Ex, Q, Pbar)
        assert np.array_equal(Pbar, g[f"r{r}_Pbar3"]), r
        fo.varin_lin(k["mf"], k["mr"], k["gammaE"], k["omega_0E"], k["delT"], J, Pol, Pbar)
        assert np.array_equal(J, g[f"r{r}_Jx"]) and np.array_equal(Pol, g[f"r{r}_P"]), r
        fo.varin_qg(k["mf"], k["mr"], k["nonLin3gammaE"], k["nonLin3Omega_0E"], k["delT"], Ex, G, Q)
        assert np.array_equal(G, g[f"r{r}_Gx3"]) and np.array_equal(Q, g[f"r{r}_Qx3"]), r
        JK = np.zeros(len(Ex))
        fo.kerr_nonlin(k["alpha3"], k["permit_0"], k["chi3Stat"], k["delT"], Ex, Eold, JK)
        assert np.array_equal(JK, g[f"r{r}_JxKerr"]), r
        fo.mur1d(k["Nz"], k["c0"], k["delT"], k["dz"], Ex, Eold)
        assert np.array_equal(Ex, g[f"r{r}_Ex"]), r
        Eold = Eold * 0.5 + 0.25 * Ex
        assert np.array_equal(Eold, g[f"r{r}_Eold_next"])
    assert np.max(np.abs(g["r2_Pbar3"])) > 0 and np.max(np.abs(g["r2_JxKerr"])) > 0


def test_oracle_reproduces_the_reference_drude_scratch_script():
    """TESTBOXDIPSERSE.py:79-94 (Drude medium in J form, hard source, no PML) stepped as written: the C restatement against
    the arrays the unmodified script text leaves behind (sizes substituted; the scheme is unstable, see make_dormant_golden)."""
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import drude_sandbox
    g = np.load(os.path.join(ROOT, "tests", "golden", "drude_sandbox.npz"), allow_pickle=False)
    k = drude_sandbox.constants(int(g["domain"]), int(g["tim"]), float(g["freq"]), int(g["nl"]), int(g["src"]), int(g["matFront"]))
    for name in ("dz", "dt", "cour", "betaE", "kapE", "perm0"):
        assert k[name] == float(g[name]), name                      # the host mirror of the script's constants
    assert np.array_equal(k["Hys"], g["Hys"])
    Ex, Hy, Jx = fo.drude_j(k)
    assert np.array_equal(Ex, g["Ex"]) and np.array_equal(Hy, g["Hy"]) and np.array_equal(Jx, g["Jx"])
    assert np.all(np.isfinite(Ex)) and np.max(np.abs(Jx)) > 0
