"""Vectorised sweep setup (sweep_setup.py + csrc/pf_setup.cu) against the per-member chain -- CPU only.

The per-member chain (Environment_Setup.envSetup -> MasterController.Params/Variables/CPML_* -> Solver_Engine.prepare_pass
-> sweep.Member) is pinned bit for bit to the reference by tests/test_host_layer.py and the goldens; here the table path
must reproduce its numbers exactly: geometry, every PfGrid scalar, the CPML profiles and the source tables.
"""
import numpy as np
import pytest

import pyfdtd_b200  # noqa: F401
from pyfdtd_b200 import Environment_Setup as envDef, MasterController as MC, Solver_Engine as SE, _native as nat
from pyfdtd_b200 import sweep, sweep_setup


def _lib_or_skip():
    try:
        return nat.lib()
    except nat.NativeError:
        pytest.skip("libpyfdtd_b200.so not built")


def per_member_lorentz(f, amp, periods=1000, passes=2):
    tup = envDef.envSetup(f, 0.7, 7000, 8000, LorMed=True)
    P = MC.Params(*tup, False, 0.7, f, 20)
    P.TFSF, P.SineCont, P.Periods, P.LorentzMed, P.FreeSpace = True, True, periods, True, False
    V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
    C_P = MC.CPML_Params(P.dz)
    C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
    out = []
    for i in range(passes):
        C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=True)
        m = sweep.Member(V, P, C_V, C_P, np.asarray(Exs) * amp, np.asarray(Hys) * amp, [P.x1Loc if i == 0 else P.x2Loc])
        m.set_mode("lorentz")
        out.append((tup, dict(m.scalars), m.flags, {k: v.copy() for k, v in m.coef.items()}, m.srcE.copy(), m.srcH.copy(),
                    m.probe_idx))
    return out


def test_envsetup_many_is_envsetup():
    freqs = np.concatenate([np.linspace(6e9, 10.5e9, 97), [9e9, 3e9, 12.3456e9]])
    for kw in (dict(LorMed=True), dict(nonLinMed=True), dict()):
        env = sweep_setup.envSetup_many(freqs, 0.7, 7000, 8000, **kw)
        for i, f in enumerate(freqs):
            tup = envDef.envSetup(float(f), 0.7, 7000, 8000, **kw)
            for k, v in zip(envDef._TUPLE_FIELDS, tup):
                assert env[k][i] == v, (kw, f, k)
    # another window / domain
    env = sweep_setup.envSetup_many(freqs[:16], 0.3, 1000, 1100, nonLinMed=True)
    for i, f in enumerate(freqs[:16]):
        tup = envDef.envSetup(float(f), 0.3, 1000, 1100, nonLinMed=True)
        assert all(env[k][i] == v for k, v in zip(envDef._TUPLE_FIELDS, tup))
    with pytest.raises(ValueError):
        sweep_setup.envSetup_many(np.array([9e9, 1e3]), 0.7, 7000, 8000)


def test_lorentz_tables_and_native_inputs_equal_the_per_member_chain():
    _lib_or_skip()
    freqs = np.array([6e9, 7.3e9, 9e9, 9e9, 10.5e9])
    amps = np.array([1.0, 0.1, 1.0, 3.7, 10.0])
    tables = sweep_setup.lorentz_sweep_tables(freqs, amps, 0.7, 7000, 8000, periods=1000)
    refs = [per_member_lorentz(float(f), float(a)) for f, a in zip(freqs, amps)]
    for i in range(len(freqs)):
        for p in range(2):
            tup, scal, flags, coef, srcE, srcH, probes = refs[i][p]
            t = tables[p]
            d = dict(zip(envDef._TUPLE_FIELDS, tup))
            assert t.L[i] == d["Nz"] + 1 and t.T[i] == d["timeSteps"] and t.pw[i] == d["pmlWidth"]
            assert (t.mf[i], t.mr[i], t.nzsrc[i]) == (scal["mf"], scal["mr"], scal["nzsrc"])
            assert list(t.probes[i]) == probes
            assert t.flags[i] == flags
            for k in sweep_setup.MemberTable.SCALARS:
                assert getattr(t, k)[i] == scal[k], (freqs[i], p, k, getattr(t, k)[i], scal[k])
    assert list(tables[0].share) == [0, 1, 2, 2, 4]
    # native profiles + sources, written at arbitrary offsets of one buffer
    t = tables[1]
    Lp, Tp = (t.L + 31) // 32 * 32, (t.T + 31) // 32 * 32
    sz = 3 * Lp + 2 * Tp
    off = np.concatenate([[0], np.cumsum(sz)])
    out = np.full(off[-1], np.nan)
    sweep_setup.build_inputs(t, out, off[:-1], off[:-1] + Lp, off[:-1] + 2 * Lp, off[:-1] + 3 * Lp, off[:-1] + 3 * Lp + Tp,
                             t.T, threads=3)
    for i in range(len(freqs)):
        _, _, _, coef, srcE, srcH, _ = refs[i][1]
        o = off[i]
        for j, name in enumerate(("beX", "ceX", "cmY")):
            assert np.array_equal(out[o + j * Lp[i]: o + j * Lp[i] + t.L[i]], coef[name]), (freqs[i], name)
        assert np.array_equal(out[o + 3 * Lp[i]: o + 3 * Lp[i] + t.T[i]], srcE), freqs[i]
        assert np.array_equal(out[o + 3 * Lp[i] + Tp[i]: o + 3 * Lp[i] + Tp[i] + t.T[i]], srcH), freqs[i]


def test_short_pulse_sources_and_nonlinear_table():
    """Periods = 1 (the LoopedSim default: the sine is switched off after one period) and the nonlinear run's pump."""
    _lib_or_skip()
    freqs = np.array([6.5e9, 9e9])
    tables = sweep_setup.lorentz_sweep_tables(freqs, 1.0, 0.7, 7000, 8000, periods=1.0)
    t = tables[0]
    out = np.zeros(int((2 * t.T).sum()))
    o = np.concatenate([[0], np.cumsum(2 * t.T)])[:-1]
    neg = np.full(t.n, -1)
    sweep_setup.build_inputs(t, out, neg, neg, neg, o, o + t.T, t.T)
    for i, f in enumerate(freqs):
        _, _, _, _, srcE, srcH, _ = per_member_lorentz(float(f), 1.0, periods=1.0, passes=1)[0]
        assert np.array_equal(out[o[i]: o[i] + t.T[i]], srcE) and np.array_equal(out[o[i] + t.T[i]: o[i] + 2 * t.T[i]], srcH)
        assert np.count_nonzero(srcE) < 500
    # nonlinear (IntegratorNL1D) members
    fr, am = np.array([6e9, 6e9, 10.5e9]), np.array([0.1, 10.0, 0.1])
    tn = sweep_setup.nonlinear_sweep_table(fr, am, 0.7, 7000, 8000)
    pairs, mine, members, share = sweep.nonlinear_sweep_members([6e9, 10.5e9], [0.1, 10.0], 0.7, 7000, 8000)
    assert pairs[:3] == [(6e9, 0.1), (6e9, 10.0), (10.5e9, 0.1)]
    out = np.zeros(int((2 * tn.T).sum()))
    o = np.concatenate([[0], np.cumsum(2 * tn.T)])[:-1]
    neg = np.full(tn.n, -1)
    sweep_setup.build_inputs(tn, out, neg, neg, neg, o, o + tn.T, tn.T)
    for i in range(3):
        m = members[i]
        m.set_mode("nl")
        for k in sweep_setup.MemberTable.SCALARS:
            assert getattr(tn, k)[i] == m.scalars[k], k
        assert tn.flags[i] == m.flags and list(tn.probes[i]) == m.probe_idx and tn.L[i] == m.L and tn.T[i] == m.T
        assert np.array_equal(out[o[i]: o[i] + tn.T[i]], m.srcE) and np.array_equal(out[o[i] + tn.T[i]: o[i] + 2 * tn.T[i]], m.srcH)


def test_numpy_sin_is_the_c_library_sin():
    """pf_setup.cu computes the source tables with libm sin where the reference calls np.sin on scalars: the two must be
    the same function on this machine (NumPy delegates float64 sin / cos to the C library)."""
    L = _lib_or_skip()
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 400, 20000), rng.uniform(0, 2e15, 2000)])
    y = np.empty_like(x)
    assert L.pf_host_sin(x.ctypes.data, y.ctypes.data, len(x), 2) == 0
    assert np.array_equal(y, np.sin(x))
    assert L.pf_host_cos(x.ctypes.data, y.ctypes.data, len(x), 2) == 0
    assert np.array_equal(y, np.cos(x))


def test_member_table_select_rebases_sharing():
    _lib_or_skip()
    freqs = np.array([6e9, 6e9, 7e9, 7e9, 7e9])
    t = sweep_setup.lorentz_sweep_tables(freqs, 1.0, 0.7, 7000, 8000)[1]
    assert list(t.share) == [0, 0, 2, 2, 2]
    s = t.select([1, 3, 4])          # owners 0 and 2 are not selected: the first selected member of each group owns
    assert list(s.share) == [0, 1, 1] and s.n == 3 and list(s.L) == [t.L[1], t.L[3], t.L[4]]


def test_reference_sweep_sizing_chain_without_objects():
    """LoopedSim's quirk (MasterController.py:547): member i is sized from member i-1's twice-corrected medium.  The scalar
    chain of sweep._sweep_chain_env against the per-member object chain (new_member_objects + spatialStab)."""
    _lib_or_skip()
    from pyfdtd_b200 import genericStability as gStab
    from test_host_layer import build_objects
    V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=6e9, dom=0.3, win=[2000, 2200], source="sine", periods=1.0))
    fr, env, wp2 = sweep._sweep_chain_env(V, P, 0.3, 2000, 2200, 5e8, 6)
    prevV, prevP, freq = V, P, P.freq_in
    for i in range(6):
        Vi, Pi, CVi, CPi = sweep.new_member_objects(freq, 0.3, 2000, 2200, prevV, prevP, P)
        wp = Vi.plasmaFreqE
        for _ in range(2):
            wp = gStab.spatialStab(Pi.timeSteps, Pi.Nz, Pi.dz, Pi.freq_in, Pi.delT, wp, Vi.omega_0E, Vi.gammaE)[3]
        assert (Pi.Nz, Pi.timeSteps, Pi.pmlWidth, Pi.nzsrc, Pi.materialFrontEdge, Pi.x1Loc, Pi.x2Loc, Pi.dz, Pi.delT) == (
            env["Nz"][i], env["timeSteps"][i], env["pmlWidth"][i], env["nzsrc"][i], env["materialFrontEdge"][i],
            env["x1Loc"][i], env["x2Loc"][i], env["dz"][i], env["delT"][i]), i
        assert wp == wp2[i] and freq == fr[i]
        sh = MC.Variables(Pi.Nz, 1, 1, 1)
        sh.plasmaFreqE = wp
        prevV, prevP, freq = sh, Pi, Pi.freq_in + 5e8
    assert len(set(env["Nlam"])) > 1                       # the chain really does change the resolution from member to member
    t = sweep_setup.lorentz_sweep_tables(fr, 1.0, 0.3, 2000, 2200, periods=1.0, env=env)
    assert np.array_equal(t[1].wp, wp2)
