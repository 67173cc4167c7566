"""GPU tests of the vectorised sweep path (sweep_setup tables + pf_host_sweep_inputs + MemberBatch.from_table +
sweep.reflection_sweep) against the per-member chain, and of the ABI-v2 argument checks."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import MasterController as MC, Solver_Engine as SE, _native as nat, sweep, sweep_setup, longgrid
    from test_host_layer import build_objects

    class NS:
        pass
    ns = NS()
    ns.MC, ns.SE, ns.nat, ns.sweep, ns.ss, ns.longgrid, ns.build_objects, ns.torch = MC, SE, nat, sweep, sweep_setup, longgrid, build_objects, torch
    return ns


DOM, WIN = 0.15, (300, 320)


def _objs(pk, freqs, periods):
    out = []
    for f in freqs:
        V, P, C_V, C_P = pk.build_objects(dict(mode="lorentz", freq=float(f), dom=DOM, win=WIN, source="sine", periods=periods))
        out.append((V, P, C_V, C_P))
    return out


def test_table_batch_is_bit_identical_to_member_batch(pk):
    """One pass of the Lorentz integrator over heterogeneous members: MemberBatch.from_table (vectorised setup, native
    inputs) against MemberBatch(list of Member) (per-member chain) -- probe traces and final fields bit for bit."""
    freqs = np.array([6e9, 7.7e9, 9e9, 9e9, 10.5e9])
    amps = np.array([1.0, 0.5, 1.0, 2.0, 1.0])
    tables = pk.ss.lorentz_sweep_tables(freqs, amps, DOM, *WIN, periods=1000)
    members = []
    for (V, P, C_V, C_P), a in zip(_objs(pk, freqs, 1000), amps):
        for _ in range(2):
            C_V, Exs, Hys = pk.SE.prepare_pass(V, P, C_V, C_P, lorentz=True)
        members.append(pk.sweep.Member(V, P, C_V, C_P, np.asarray(Exs) * a, np.asarray(Hys) * a, [P.x2Loc]))
    a_batch = pk.sweep.MemberBatch(members, "lorentz")
    b_batch = pk.sweep.MemberBatch.from_table(tables[1], "lorentz")
    assert b_batch.n_in < a_batch.n_in          # members 2 and 3 share their CPML profiles in the table path
    for b in (a_batch, b_batch):
        b.upload()
        b.reset_state()
        b.run(do_pol=True)
    ta, tb = a_batch.download_probes(), b_batch.download_probes()
    for i in range(len(freqs)):
        assert np.array_equal(ta[i], tb[i]), i
        for name in ("Ex", "Hy", "Dx", "P", "Pprev", "psiE", "psiH"):
            assert np.array_equal(a_batch.state(i, name), b_batch.state(i, name)), (i, name)
    assert np.max(np.abs(ta[0])) > 1e-3


def test_reflection_sweep_equals_the_per_member_two_pass_batch(pk):
    freqs = np.linspace(6e9, 10.5e9, 7)
    DOM, WIN = 0.3, (2000, 2200)          # the geometry of the reference's sweep golden: long enough for the reflection to return
    objs = []
    for f in freqs:
        objs.append(pk.build_objects(dict(mode="lorentz", freq=float(f), dom=DOM, win=WIN, source="sine", periods=1.0)))
    _, _, want = pk.sweep.run_two_pass_batch(objs, lorentz=True, device_reflection=True)
    got = pk.sweep.reflection_sweep(freqs, DOM, *WIN, periods=1.0, chunk=3)       # 3 chunks: 3 + 3 + 1 members
    # the time stepping is bit-identical (test above); the reflection figure goes through a batched cuFFT whose plan -- and with
    # it the last bit of the peak magnitude -- depends on the batch size and stride, so the two routes agree to an ulp
    np.testing.assert_allclose(got["measured"], want, rtol=1e-13, atol=0)
    assert list(got["index"]) == list(range(7)) and got["cell_steps"] > 0
    # the analytical figure is results(AnalRefCo=True) with the medium as the two passes leave it
    for i, (V, P, C_V, C_P) in enumerate(objs):
        assert got["analytical"][i] == pk.MC.results(V, P, C_V, C_P, None, AnalRefCo=True)
    # sharded over two ranks: same numbers, dealt round-robin
    r0 = pk.sweep.reflection_sweep(freqs, DOM, *WIN, periods=1.0, rank=0, world_size=2)
    r1 = pk.sweep.reflection_sweep(freqs, DOM, *WIN, periods=1.0, rank=1, world_size=2)
    np.testing.assert_allclose(r0["measured"], want[0::2], rtol=1e-13, atol=0)
    np.testing.assert_allclose(r1["measured"], want[1::2], rtol=1e-13, atol=0)
    assert 0.0 < got["measured"].min() and got["measured"].max() < 1.0


def test_steps_past_the_source_tables_are_rejected(pk):
    """PfGrid.n_src / probe_stride: pf_run_pass, pf_run_batch and pf_run_block return PF_E_ARG instead of reading / writing
    past the caller's tables (CoupledPIC.step, LongGrid.run, run_time_loop(nsteps=...))."""
    V, P, C_V, C_P = pk.build_objects(dict(mode="lorentz", freq=9e9, dom=DOM, win=WIN, source="sine"))
    C_V, Exs, Hys = pk.SE.prepare_pass(V, P, C_V, C_P, lorentz=True)
    with pytest.raises(ValueError, match="source tables|probe rows"):
        pk.SE.run_time_loop(V, P, C_V, C_P, "lorentz", True, Exs, Hys, [P.x2Loc], n0=P.timeSteps - 10, nsteps=20)
    m = pk.sweep.Member(V, P, C_V, C_P, Exs, Hys, [P.x2Loc], nsteps=P.timeSteps + 1)
    b = pk.sweep.MemberBatch([m], "lorentz")
    b.upload()
    b.reset_state()
    with pytest.raises(ValueError, match="source tables|probe rows"):
        b.run(do_pol=True)
    lg, info = pk.longgrid.lorentz_long_grid(20000, T=64, k=32)
    lg.run(64)
    with pytest.raises(ValueError, match="source tables"):
        lg.run(1)
    # straight through the C-ABI as well
    rc = pk.nat.lib().pf_run_block(lg.grids[lg.cur], lg.grids[lg.cur ^ 1], len(lg.mine), lg.mode_id, 1, 64, 8, 32, 0,
                                   lg.scratch.data_ptr(), lg.scratch_bytes, pk.nat.current_stream_ptr())
    assert rc == -1 and b"source tables" in pk.nat.lib().pf_last_error()


def test_block_tables_are_rebuilt_unless_the_caller_vouches_for_them(pk):
    """pf_run_block keeps no record of earlier calls (the process-global table cache): two LongGrid objects that
    end up at the same scratch address, the scratch overwritten in between, still give the undecomposed result."""
    torch = pk.torch
    ref, _ = pk.longgrid.lorentz_long_grid(30000, T=128, k=32, max_piece=1 << 27)
    ref.run(128)
    want = ref.gather_owned("Ex")
    for attempt in range(2):
        lg, _ = pk.longgrid.lorentz_long_grid(30000, T=128, k=32, max_piece=9000)      # several pieces, exchanges between them
        addr = lg.scratch.data_ptr()
        lg.run(64)
        lg.scratch.fill_(0xAB)            # somebody else's data now lives where the tables were ...
        lg._tables_built = False          # ... and the owner says so: no TABLES_VALID promise for the next call
        lg.run(64)
        assert np.array_equal(lg.gather_owned("Ex"), want)
        del lg
        torch.cuda.empty_cache()
    assert np.max(np.abs(want)) > 0
