"""GPU tests of the long-grid / domain-decomposition path: a grid far beyond the reference's Nz cap,
cut into several pieces with ghost exchange every k steps, must equal the CPU oracle's undecomposed
run bit for bit (same operations in the same order on every owned cell)."""
import ctypes

import numpy as np
import pytest

import fdtd_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import longgrid
    return longgrid


def oracle_long(info, Lg, T, nsteps, mode="lorentz"):
    c = fo.Case(mode=mode, freq=9e9, Nz=Lg - 1, T=T, pw=info["pw"], mf=info["mf"], mr=info["mr"], nzsrc=info["nzsrc"],
                x1Loc=info["mf"] - 20, x2Loc=info["nzsrc"] - 100, dz=info["dz"], dt=info["dt"],
                courantNo=info["courantNo"], period=1 / 9e9, source="sine", tfsf=True, Periods=1000.0)
    pa = fo.PassArrays(c, info["V"].plasmaFreqE, info["Exs"], info["Hys"], [info["nzsrc"] - 100], False)
    fo.lib().orc_run(ctypes.byref(pa.g), fo.MODE_ID[mode], 1, 0, nsteps, T)
    return pa


@pytest.mark.parametrize("k,max_piece,split", [(64, 9000, False), (17, 7000, False), (64, 1 << 27, False),
                                               (64, 9000, True), (17, 1500, True), (25, 7000, True)])
def test_decomposed_long_grid_is_bit_identical_to_oracle(lg, k, max_piece, split):
    """split: every block as two launches, inner tiles (PF_BLOCK_F_INNER_TILES) then edge tiles (PF_BLOCK_F_EDGE_TILES) --
    the form LongGrid.run uses to overlap a remote ghost exchange; 1500-cell pieces consist of edge tiles only."""
    Lg, T, nsteps = 40_000, 512, 300
    grid, info = lg.lorentz_long_grid(Lg, T=T, k=k, max_piece=max_piece)
    grid.force_split = split
    assert Lg > 25000, "beyond the reference's grid-size guard"
    if max_piece < Lg:
        assert len(grid.pieces) >= 4
    grid.run(nsteps, do_pol=True)
    pa = oracle_long(info, Lg, T, nsteps)
    for name, want in (("Ex", pa.Ex), ("Hy", pa.Hy), ("Dx", pa.Dx), ("P", pa.P), ("Pprev", pa.Pprev), ("psiH", pa.psiH)):
        got = grid.gather_owned(name)
        if name in ("Dx", "P", "Pprev"):      # only defined on the slab
            sl = slice(info["mf"], info["mr"])
            assert np.array_equal(got[sl], want[sl]), name
        elif name == "psiH":
            pw = info["pw"]
            assert np.array_equal(got[1:pw], want[1:pw]) and np.array_equal(got[Lg - pw:Lg - 1], want[Lg - pw:Lg - 1])
        else:
            assert np.array_equal(got, want), name
    assert np.max(np.abs(pa.Ex)) > 0
    probe = grid.probe_out[0, :nsteps].cpu().numpy()
    assert np.array_equal(probe, pa.probe_out[0, :nsteps])


def test_decomposed_kerr_lorentz_grid_matches_oracle(lg):
    """Config 5's material (PF_LORENTZ_NL: Lorentz ADE + cubic Kerr law) on a decomposed grid."""
    Lg, T, nsteps = 40_000, 512, 300
    grid, info = lg.lorentz_long_grid(Lg, T=T, k=64, max_piece=9000, mode="lorentz_nl")
    assert len(grid.pieces) >= 4
    grid.run(nsteps, do_pol=True)
    pa = oracle_long(info, Lg, T, nsteps, mode="lorentz_nl")
    sl = slice(info["mf"], info["mr"])
    scale = np.max(np.abs(pa.Ex))
    assert scale > 0
    for name, want in (("Ex", pa.Ex), ("Hy", pa.Hy), ("Dx", pa.Dx), ("P", pa.P)):
        got = grid.gather_owned(name)
        if name in ("Dx", "P"):
            got, want = got[sl], want[sl]
        assert np.max(np.abs(got - want)) <= 1e-10 * np.max(np.abs(want)), name


def test_long_free_space_grid(lg):
    Lg, T, nsteps = 30_000, 256, 200
    grid, info = lg.lorentz_long_grid(Lg, T=T, k=32, max_piece=8000, mode="free")
    grid.run(nsteps, do_pol=False)
    c = fo.Case(mode="free", freq=9e9, Nz=Lg - 1, T=T, pw=info["pw"], mf=info["mf"], mr=info["mr"], nzsrc=info["nzsrc"],
                x1Loc=info["mf"] - 20, x2Loc=info["nzsrc"] - 100, dz=info["dz"], dt=info["dt"],
                courantNo=info["courantNo"], period=1 / 9e9, source="sine", tfsf=True, Periods=1000.0)
    pa = fo.PassArrays(c, 0.0, info["Exs"], info["Hys"], [], False)
    fo.lib().orc_run(ctypes.byref(pa.g), 0, 0, 0, nsteps, T)
    assert np.array_equal(grid.gather_owned("Ex"), pa.Ex) and np.array_equal(grid.gather_owned("Hy"), pa.Hy)
