"""CPU tests of the multi-GPU plumbing (no GPU needed): piece planning, the ghost-exchange schedule,
and the exchange itself over torch.distributed with the gloo backend at world_size 2, using tensor
slicing as a stand-in for the pack / unpack kernels.  Also the round-robin sharding of sweep members."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyfdtd_b200  # noqa: F401
from pyfdtd_b200 import longgrid as lgm


@pytest.mark.parametrize("Lg,pw,world,k", [(100_000, 2394, 1, 64), (100_000, 2394, 2, 64), (1_000_000, 2394, 8, 64),
                                           (40_000, 2394, 4, 17), (20_000, 300, 3, 64)])
def test_plan_pieces_covers_grid(Lg, pw, world, k):
    pieces = lgm.plan_pieces(Lg, pw, world, k, max_piece=30_000)
    assert pieces[0]["lo"] == 0 and pieces[-1]["hi"] == Lg
    for a, b in zip(pieces[:-1], pieces[1:]):
        assert a["hi"] == b["lo"] and a["rank"] <= b["rank"]
    assert sorted({p["rank"] for p in pieces}) == list(range(world))
    for p in pieces:
        assert p["hi"] - p["lo"] >= 2 * k or len(pieces) == 1
        assert p["hi"] - p["lo"] <= 30_000 + 1
    for r in range(world):      # every rank owns exactly its contiguous share
        own = [p for p in pieces if p["rank"] == r]
        assert own[0]["lo"] == r * Lg // world and own[-1]["hi"] == (r + 1) * Lg // world
    assert pieces[0]["ghost_l"] == 0 and pieces[-1]["ghost_r"] == 0
    assert all(p["ghost_l"] == k for p in pieces[1:]) and all(p["ghost_r"] == k for p in pieces[:-1])


def test_rank_cuts_balance_work_not_cells():
    Lg, pw, world = 2_000_000, 2394, 4
    mf, mr = int(0.3 * Lg), Lg - 2
    cuts = lgm.balanced_rank_cuts(Lg, pw, world, mf, mr)
    assert cuts[0] == 0 and cuts[-1] == Lg and all(b > a for a, b in zip(cuts[:-1], cuts[1:]))

    def work(a, b):
        z = np.arange(a, b)
        w = np.where((z >= mf) & (z < mr), lgm.CELL_COST["slab"], lgm.CELL_COST["vacuum"])
        w = w + np.where((z < pw) | (z >= Lg - pw), lgm.CELL_COST["cpml"], 0.0)
        return float(w.sum())
    loads = [work(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert max(loads) / min(loads) < 1.01
    assert cuts[1] > Lg // world          # the vacuum-side rank takes more cells
    pieces = lgm.plan_pieces(Lg, pw, world, 64, mf=mf, mr=mr)
    assert [p["lo"] for p in pieces if p["ghost_l"] == 64 and p["rank"] != pieces[p["index"] - 1]["rank"]] == cuts[1:-1]


def test_exchange_schedule_is_symmetric():
    pieces = lgm.plan_pieces(200_000, 2394, 2, 64, max_piece=40_000)
    s0, s1 = lgm.exchange_schedule(pieces, 0), lgm.exchange_schedule(pieces, 1)
    sends0 = [(i, j) for kind, i, side, j in s0 if kind == "send"]
    recvs1 = [(j, i) for kind, i, side, j in s1 if kind == "recv"]
    assert sends0 == recvs1 and len(sends0) == 1      # one rank boundary -> one message each way
    assert sum(1 for kind, *_ in s0 if kind == "local") == sum(1 for p in pieces if p["rank"] == 0) - 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Lg, pw, k, n_arr, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pieces = lgm.plan_pieces(Lg, pw, world, k, max_piece=3000)
    mine = [p for p in pieces if p["rank"] == rank]
    # "state": value of global cell z in array a is a*1e6 + z ; ghosts start as -1
    local = {}
    for p in mine:
        z0 = p["lo"] - p["ghost_l"]
        L = p["hi"] + p["ghost_r"] - z0
        arr = torch.full((n_arr, L), -1.0, dtype=torch.float64)
        own = torch.arange(p["lo"], p["hi"], dtype=torch.float64)
        for a in range(n_arr):
            arr[a, p["ghost_l"]: p["ghost_l"] + len(own)] = a * 1e6 + own
        local[p["index"]] = arr

    def pack(i, side):          # stand-in for pf_halo_pack: the k owned cells next to the edge
        arr, p = local[i], pieces[i]
        L = arr.shape[1]
        sl = slice(k, 2 * k) if side == 0 else slice(L - 2 * k, L - k)
        assert p["ghost_l" if side == 0 else "ghost_r"] == k
        return arr[:, sl].reshape(-1).clone()

    def unpack(i, side, buf):   # stand-in for pf_halo_unpack: the k ghost cells at the edge
        arr = local[i]
        L = arr.shape[1]
        sl = slice(0, k) if side == 0 else slice(L - k, L)
        arr[:, sl] = buf.reshape(n_arr, k)

    sched = lgm.exchange_schedule(pieces, rank)
    lgm.run_exchange(sched, pieces, pack, unpack, dist=dist, make_buffer=lambda i, side: torch.empty(n_arr * k, dtype=torch.float64))
    ok = True
    for p in mine:               # every ghost cell now holds its global cell's value
        z0 = p["lo"] - p["ghost_l"]
        arr = local[p["index"]]
        z = torch.arange(z0, z0 + arr.shape[1], dtype=torch.float64)
        for a in range(n_arr):
            ok &= bool(torch.equal(arr[a], a * 1e6 + z))
    # sweep-member sharding: member % world == rank, results gathered on rank 0 without a data-path collective
    members = list(range(11))
    mine_m = [m for m in members if m % world == rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, {m: m * m for m in mine_m})
    merged = {k_: v for d in gathered for k_, v in d.items()}
    ok &= merged == {m: m * m for m in members}
    np.save(os.path.join(result_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_ghost_exchange_over_gloo_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 24_000, 700, 32, 7, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"ok{r}.npy")[0])


def test_rank_boundaries_follow_the_mode_specific_slab_cost():
    """balanced_rank_cuts: a costlier slab (Kerr-Lorentz) pushes the boundaries towards the vacuum side; estimated work
    per rank is equal for the ratio the cuts were made with."""
    import importlib
    lg = importlib.import_module("pyfdtd_b200.longgrid")
    Lg, pw, mf, mr, world = 8_000_000, 2394, 2_400_000, 7_999_998, 8
    cuts_l = lg.balanced_rank_cuts(Lg, pw, world, mf, mr, lg.SLAB_COST["lorentz"])
    cuts_k = lg.balanced_rank_cuts(Lg, pw, world, mf, mr, lg.SLAB_COST["lorentz_nl"])
    assert cuts_l[0] == cuts_k[0] == 0 and cuts_l[-1] == cuts_k[-1] == Lg
    assert all(b > a for a, b in zip(cuts_k[:-1], cuts_k[1:]))
    assert cuts_k[1] > cuts_l[1]            # rank 0 (all the vacuum) takes more cells when the slab is dearer
    for cuts, r in ((cuts_l, lg.SLAB_COST["lorentz"]), (cuts_k, lg.SLAB_COST["lorentz_nl"])):
        work = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            slab = max(0, min(b, mr) - max(a, mf))
            work.append((b - a - slab) + r * slab)
        assert max(work) / min(work) < 1.02


def test_nonlinear_sweep_members_shard_without_overlap():
    """Host half of sweep.nonlinear_sweep (config 3): frequency-major (f, amp) list dealt round-robin over the ranks;
    one setup chain per distinct frequency on a rank, sources scaled by the amplitude, profile sharing inside the rank."""
    import importlib
    sw = importlib.import_module("pyfdtd_b200.sweep")
    freqs, amps = [9e9, 7.5e9, 6e9], [0.5, 2.0, 8.0]
    world = 2
    seen = []
    for rank in range(world):
        pairs, mine, members, share = sw.nonlinear_sweep_members(freqs, amps, 0.15, 300, 320, nsteps=100, rank=rank,
                                                                 world_size=world)
        assert pairs == [(f, a) for f in freqs for a in amps]
        assert mine == [i for i in range(9) if i % world == rank] and len(members) == len(mine)
        seen += mine
        for j, i in enumerate(mine):
            f, a = pairs[i]
            m = members[j]
            assert m.P.freq_in == f and m.nsteps == 100 and m.probe_idx == [m.P.materialFrontEdge, m.P.materialRearEdge]
            first = share[j]
            assert first <= j and members[first].P is m.P                      # shares the first same-frequency member's grid
            a0 = pairs[mine[first]][1]
            assert np.allclose(m.srcE / a, members[first].srcE / a0, rtol=1e-14, atol=0)
    assert sorted(seen) == list(range(9))
