"""GPU parity of the PIC push / sort / deposit against oracle/pic_oracle.py (the model definition;
the reference has no particle code).  Tolerance 1e-12 relative (BASELINE north_star); bit-exact expected."""
import numpy as np
import pytest

import pic_oracle as po

pytestmark = pytest.mark.gpu

C0, MU0, QM, Q = 299792458.0, 1.25663706127e-06, -1.75882001076e11, -1.602176634e-19


@pytest.fixture(scope="module")
def pic():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import pic as picmod
    return picmod


def fields(L, seed=3):
    rng = np.random.default_rng(seed)
    x = np.arange(L)
    Ex = 2e5 * np.sin(2 * np.pi * x / 173.0) + 1e3 * rng.standard_normal(L)
    Hy = (3e5 / 376.73) * np.cos(2 * np.pi * x / 173.0) + rng.standard_normal(L)
    return Ex, Hy


def rel(got, want):
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300))


@pytest.mark.parametrize("n", [1, 1000, 200_003])
def test_push_sort_deposit_match_oracle(pic, n):
    import torch
    L, dz, dt = 4097, 8.3e-5, 2.6e-13
    z, ux, uz, w, cell = po.make_beam(n, L, dz, seed=7)
    Ex, Hy = fields(L)
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    tEx, tHy = torch.as_tensor(Ex, device="cuda"), torch.as_tensor(Hy, device="cuda")
    state = (z, ux, uz)
    for step in range(3):
        ps.push(tEx, tHy)
        zo, uxo, uzo, co = po.push(*state, Ex, Hy, dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
        h = ps.host()
        assert rel(h["z"], zo) <= 1e-12 and rel(h["ux"], uxo) <= 1e-12 and rel(h["uz"], uzo) <= 1e-12
        assert np.array_equal(h["cell"], co)
        state = (zo, uxo, uzo)
        ps2 = None
    # sort: stable by cell
    zo, uxo, uzo, wo, co = po.sort_by_cell(state[0], state[1], state[2], w, co)
    ps.sort()
    h = ps.host()
    assert np.array_equal(h["cell"], co)
    assert np.array_equal(h["z"], zo) or rel(h["z"], zo) <= 1e-12
    # deposit
    J = ps.deposit().cpu().numpy()
    Jo = po.deposit(h["z"], h["ux"], h["uz"], h["w"], h["cell"], L, dz=dz, c=C0, jx_scale=Q)
    assert rel(J, Jo) <= 1e-12
    assert np.array_equal(J, Jo), "deterministic deposition is expected to be bit-identical"
    # conservation: total deposited current = sum of q w vx
    g = np.sqrt(1 + (h["ux"] ** 2 + h["uz"] ** 2) / C0 ** 2)
    assert np.sum(J) == pytest.approx(Q * np.sum(h["w"] * h["ux"] / g), rel=1e-9)


def test_deposit_is_deterministic_and_order_independent_of_launch(pic):
    L, dz, dt = 2049, 8.3e-5, 2.6e-13
    z, ux, uz, w, cell = po.make_beam(50_000, L, dz, seed=11)
    a = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    b = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    Ja = a.deposit().cpu().numpy()
    for _ in range(3):
        Jb = b.deposit().cpu().numpy()
        assert np.array_equal(Ja, Jb)


def test_walls_reflect_and_empty_set(pic):
    import torch
    L, dz, dt = 513, 1e-4, 3e-13
    zmax = (L - 1) * dz
    z = np.array([1e-9, zmax - 1e-9, 0.5 * zmax])
    uz = np.array([-2e8, 2e8, 0.0])
    ux = np.zeros(3)
    w = np.ones(3)
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    zero = torch.zeros(L, dtype=torch.float64, device="cuda")
    ps.push(zero, zero)
    h = ps.host()
    zo, uxo, uzo, co = po.push(z, ux, uz, np.zeros(L), np.zeros(L), dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
    assert np.array_equal(h["z"], zo) and np.array_equal(h["uz"], uzo)
    assert h["uz"][0] > 0 and h["uz"][1] < 0 and np.all((h["z"] >= 0) & (h["z"] <= zmax))
    empty = pic.ParticleSet(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), L, dz, dt)
    empty.push(zero, zero)
    assert not np.any(empty.deposit().cpu().numpy())


def test_coupled_step_feeds_the_jx_slot(pic):
    """One coupled step: the deposited current is what ADE_ExUpdate subtracts (BaseFDTD11.py:667)."""
    import ctypes
    import fdtd_oracle as fo
    from pyfdtd_b200 import BaseFDTD11, Solver_Engine as SE, _device as dev
    from test_host_layer import build_objects
    spec = dict(mode="free", freq=9e9, dom=0.15, win=[300, 320], source="sine", periods=1000, epsRe=1.0)
    V, P, C_V, C_P = build_objects(spec)
    C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False)
    L = len(V.Ex)
    z, ux, uz, w, cell = po.make_beam(20_000, L, P.dz, seed=5)
    ux = ux + 5e7
    ps = pic.ParticleSet(z, ux, uz, w * 1e3, L, P.dz, P.delT)
    arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
    g = dev.DeviceGrid(L=L, T=P.timeSteps, arrays=arrs, scalars=BaseFDTD11.grid_scalars(V, P),
                       srcE=np.asarray(Exs) / P.courantNo, srcH=np.asarray(Hys) / P.courantNo, probe_idx=[],
                       flags=BaseFDTD11.grid_flags(P))
    sim = pic.CoupledPIC(g, ps, mode="free")
    sim.step()
    # oracle: deposit -> one reference step with that Jx -> push
    ps0 = po.sort_by_cell(z, ux, uz, w * 1e3, cell)
    Jo = po.deposit(*ps0, L, dz=P.dz, c=C0, jx_scale=Q)
    c = fo.make_case("free", 9e9, 0.15, 300, 320, source="sine", periods=1000, epsRe=1.0)
    pa = fo.PassArrays(c, 0.0, Exs, Hys, [], False, Jx=Jo)
    fo.lib().orc_run(ctypes.byref(pa.g), 0, 0, 0, 1, c.T)
    out = g.fetch(["Ex", "Hy"], probes=False)
    assert np.max(np.abs(Jo)) > 0
    assert np.array_equal(out["Ex"], pa.Ex) and np.array_equal(out["Hy"], pa.Hy)
    zo, uxo, uzo, co = po.push(ps0[0], ps0[1], ps0[2], pa.Ex, pa.Hy, dz=P.dz, dt=P.delT, q_over_m=QM, c=C0, mu0=MU0)
    zo, uxo, uzo, wo, co = po.sort_by_cell(zo, uxo, uzo, ps0[3], co)      # the coupled step keeps the set cell-sorted
    h = ps.host()
    assert np.array_equal(h["cell"], co)
    assert rel(h["z"], zo) <= 1e-12 and rel(h["ux"], uxo) <= 1e-12


@pytest.mark.parametrize("n,L", [(5000, 4097), (300_001, 4097), (200_003, 130), (77_777, 33)])
def test_fused_push_resort_equals_push_then_stable_sort(pic, n, L):
    """pf_pic_push_sorted (count / scan / move, no radix sort) == oracle push followed by a stable sort,
    bit for bit, over several steps; the deposit of the result matches too.  The crowded cases (1.5-2.4 k
    particles per cell) run several warps per cell."""
    import torch
    dz, dt = 8.3e-5, 2.6e-13
    z, ux, uz, w, cell = po.make_beam(n, L, dz, seed=21, thermal=0.3)
    Ex, Hy = fields(L, seed=9)
    tEx, tHy = torch.as_tensor(Ex, device="cuda"), torch.as_tensor(Hy, device="cuda")
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    zo, uxo, uzo, wo, co = po.sort_by_cell(z, ux, uz, w, cell)
    for step in range(4):
        ps.push_sorted(tEx, tHy)
        zo, uxo, uzo, co = po.push(zo, uxo, uzo, Ex, Hy, dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
        zo, uxo, uzo, wo, co = po.sort_by_cell(zo, uxo, uzo, wo, co)
        h = ps.host()
        assert np.array_equal(h["cell"], co), step
        for k, want in (("z", zo), ("ux", uxo), ("uz", uzo), ("w", wo)):
            assert np.array_equal(h[k], want), (step, k)
    assert not ps.cfl_violated()
    J = ps.deposit().cpu().numpy()
    Jo = po.deposit(zo, uxo, uzo, wo, co, L, dz=dz, c=C0, jx_scale=Q)
    assert np.array_equal(J, Jo)


@pytest.mark.parametrize("n,L", [(5000, 4097), (300_001, 4097), (200_003, 130), (77_777, 33), (0, 65)])
def test_fused_step_push_resort_deposit(pic, n, L):
    """pf_pic_step_sorted: the particles equal pf_pic_push_sorted's bit for bit; Jx equals the oracle's statement of
    the fused summation tree bit for bit, and the plain deposit of the same particles to rounding."""
    import torch
    dz, dt = 8.3e-5, 2.6e-13
    z, ux, uz, w, cell = po.make_beam(n, L, dz, seed=23, thermal=0.3)
    Ex, Hy = fields(L, seed=9)
    tEx, tHy = torch.as_tensor(Ex, device="cuda"), torch.as_tensor(Hy, device="cuda")
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    S = ps.sub_warps()
    assert S == po.sub_warps(n, L)
    zo, uxo, uzo, wo, co = po.sort_by_cell(z, ux, uz, w, cell)
    for step in range(3):
        J = ps.step_sorted(tEx, tHy).cpu().numpy()
        zp, uxp, uzp, cp = po.push(zo, uxo, uzo, Ex, Hy, dz=dz, dt=dt, q_over_m=QM, c=C0, mu0=MU0)
        Jf = po.deposit_fused(zp, uxp, uzp, wo, co, cp, L, S, dz=dz, c=C0, jx_scale=Q)   # pushed, old order
        zo, uxo, uzo, wo, co = po.sort_by_cell(zp, uxp, uzp, wo, cp)
        h = ps.host()
        for k, want in (("z", zo), ("ux", uxo), ("uz", uzo), ("w", wo), ("cell", co)):
            assert np.array_equal(h[k], want), (step, k)
        assert np.array_equal(J, Jf), step
        if n:
            Jp = po.deposit(zo, uxo, uzo, wo, co, L, dz=dz, c=C0, jx_scale=Q)
            assert rel(J, Jp) <= 1e-12
    J2 = ps.deposit().cpu().numpy()            # the plain deposit still works on the result
    if n:
        assert rel(J2, J) <= 1e-12
    else:
        assert not np.any(J)


def test_coupled_fused_step_matches_unfused(pic):
    """CoupledPIC(fused=True) (field step, then push + re-sort + deposit in one pass) against the unfused sequence."""
    from pyfdtd_b200 import BaseFDTD11, Solver_Engine as SE, _device as dev
    from test_host_layer import build_objects
    spec = dict(mode="free", freq=9e9, dom=0.15, win=[300, 320], source="sine", periods=1000, epsRe=1.0)
    outs = []
    for fused in (False, True):
        V, P, C_V, C_P = build_objects(spec)
        C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False)
        L = len(V.Ex)
        z, ux, uz, w, cell = po.make_beam(20_000, L, P.dz, seed=5)
        ps = pic.ParticleSet(z, ux + 5e7, uz, w * 1e3, L, P.dz, P.delT)
        arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
        g = dev.DeviceGrid(L=L, T=P.timeSteps, arrays=arrs, scalars=BaseFDTD11.grid_scalars(V, P),
                           srcE=np.asarray(Exs) / P.courantNo, srcH=np.asarray(Hys) / P.courantNo, probe_idx=[],
                           flags=BaseFDTD11.grid_flags(P))
        sim = pic.CoupledPIC(g, ps, mode="free", fused=fused)
        for _ in range(20):
            sim.step()
        f = g.fetch(["Ex", "Hy"], probes=False)
        outs.append((f["Ex"], f["Hy"], ps.host()))
    (ExA, HyA, hA), (ExB, HyB, hB) = outs
    assert np.max(np.abs(ExA)) > 0
    assert rel(ExB, ExA) <= 1e-12 and rel(HyB, HyA) <= 1e-12
    assert rel(hB["z"], hA["z"]) <= 1e-12 and rel(hB["ux"], hA["ux"]) <= 1e-12 and rel(hB["uz"], hA["uz"]) <= 1e-12


@pytest.mark.parametrize("engine", ["ops", "tile"])
@pytest.mark.parametrize("mode", ["free", "lorentz", "nl"])
def test_beam_inside_the_medium_drives_the_fields(pic, mode, engine):
    """CoupledPIC with the beam overlapping the slab, every material mode, both engines, against the oracle chain
    deposit -> one field step with that Jx (ADE_ExUpdate outside, ADE_DxUpdate inside the slab) -> push -> stable sort.
    Bit-identical fields for the linear modes; the cubic mode is held to its 1e-10."""
    import ctypes
    import fdtd_oracle as fo
    from pyfdtd_b200 import Solver_Engine as SE
    from test_host_layer import build_objects
    spec = dict(mode=mode, freq=9e9, dom=0.15, win=[300, 320], source="sine", periods=1000, epsRe=1.0)
    V, P, C_V, C_P = build_objects(spec)
    C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=(mode == "lorentz"), nonlinear=(mode == "nl"))
    L = len(V.Ex)
    z, ux, uz, w, cell = po.make_beam(30_000, L, P.dz, seed=7)              # uniform over 5 % .. 95 % of the grid: most of it in the slab
    ux = ux + 5e7
    w = w * 1e-2            # beam-driven |Ex| ~ 10 V/m in the slab: the amplitude range of the nonlinear sweeps (at 1e6 V/m the
                            # closed-form cubic is ill-conditioned at 1e-9 between any two libm implementations)
    assert np.sum((z / P.dz > P.materialFrontEdge) & (z / P.dz < P.materialRearEdge)) > 10_000
    ps = pic.ParticleSet(z, ux, uz, w, L, P.dz, P.delT)
    g = pic.coupled_grid(V, P, C_V, C_P, Exs, Hys, mode=mode)
    sim = pic.CoupledPIC(g, ps, mode=mode, engine=engine)
    do_pol = mode == "lorentz"
    nsteps = 6
    for _ in range(nsteps):
        sim.step(do_pol=do_pol)
    # oracle
    c = fo.make_case(mode, 9e9, 0.15, 300, 320, source="sine", periods=1000, epsRe=1.0)
    pa = fo.PassArrays(c, V.plasmaFreqE, np.asarray(Exs), np.asarray(Hys), [], False, Jx=np.zeros(L))
    zo, uxo, uzo, wo, co = po.sort_by_cell(z, ux, uz, w, cell)
    for n in range(nsteps):
        pa.Jx[:] = po.deposit(zo, uxo, uzo, wo, co, L, dz=P.dz, c=C0, jx_scale=Q)
        fo.lib().orc_run(ctypes.byref(pa.g), fo.MODE_ID[mode], int(do_pol), n, 1, c.T)
        zo, uxo, uzo, co = po.push(zo, uxo, uzo, pa.Ex, pa.Hy, dz=P.dz, dt=P.delT, q_over_m=QM, c=C0, mu0=MU0)
        zo, uxo, uzo, wo, co = po.sort_by_cell(zo, uxo, uzo, wo, co)
    out = g.fetch(["Ex", "Hy", "Dx", "P"], probes=False)
    slab = slice(P.materialFrontEdge, P.materialRearEdge)
    assert np.max(np.abs(pa.Ex[slab])) > 0 and (mode == "free" or np.max(np.abs(pa.Dx[slab])) > 0)   # the beam did drive the medium
    if mode == "nl":
        assert rel(out["Ex"], pa.Ex) <= 1e-10 and rel(out["Hy"], pa.Hy) <= 1e-10 and rel(out["Dx"], pa.Dx) <= 1e-10
    else:
        for name in ("Ex", "Hy", "Dx") + (("P",) if mode == "lorentz" else ()):
            if mode == "free" and name == "Dx":
                continue
            assert np.array_equal(out[name], getattr(pa, name)), name
    h = ps.host()
    tol = 1e-12 if mode != "nl" else 1e-9
    assert rel(h["z"], zo) <= tol and rel(h["ux"], uxo) <= tol and rel(h["uz"], uzo) <= tol
    # without the builder-defined Dx term the slab would not feel the beam at all: Ex there would equal the beam-free run
    assert (sim.engine == 1) == (engine == "tile")


def test_tile_engine_with_a_static_current_equals_the_per_op_engine(pic):
    """run_time_loop with a prescribed V.Jx (constant in time): the fused tile kernel (k = 64 steps per launch, Jx carried
    as a per-cell input) against the per-op engine, Lorentz mode, bit for bit."""
    from pyfdtd_b200 import Solver_Engine as SE
    from test_host_layer import build_objects
    outs = []
    for engine in ("ops", "tile"):
        SE.ENGINE = engine
        try:
            V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=9e9, dom=0.15, win=[300, 320], source="sine", periods=1000))
            C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=True)
            L = len(V.Ex)
            V.Jx = 1e-3 * np.sin(np.arange(L) * 0.01) * (np.arange(L) > P.pmlWidth) * (np.arange(L) < L - P.pmlWidth)
            tr = SE.run_time_loop(V, P, C_V, C_P, "lorentz", True, Exs, Hys, [P.x2Loc], nsteps=300)
            assert SE.LAST_RUN_INFO["engine"] == engine
            outs.append((V.Ex.copy(), V.Hy.copy(), V.Dx.copy(), V.polarisationCurr.copy(), tr.copy()))
        finally:
            SE.ENGINE = "auto"
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert np.max(np.abs(outs[0][0])) > 0
