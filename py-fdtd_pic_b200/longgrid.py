"""Long single grids (1e8 - 1e9 cells) and their 1-D domain decomposition over GPUs (config 5).

The reference caps a run at Nz <= 25000 (BaseFDTD11.py:48); the physics of its integrators does not
depend on the grid length, so the same update rules are applied here to a grid of any length by
cutting it into PIECES.  A piece is a contiguous range of global cells plus k ghost cells on every
interior side, described by one ``PfGrid`` with ``z0``/``Lg`` set (all index-dependent rules are
evaluated on global indices by the kernels).  Time advances in blocks of k steps:

    exchange ghosts (k owned cells next to every interior boundary -> neighbour's ghost cells)
    pf_run_block: every piece of this rank advances k steps, buffer A -> buffer B (tile engine)

After k steps exactly the ghost cells are stale, and the next exchange refreshes them, so the result
is bit-identical to the undecomposed run (same operations in the same order on every owned cell).
Pieces of one rank exchange through device buffers; pieces on different ranks through
``torch.distributed`` point-to-point messages (NCCL over NVLink on GPUs; gloo in the CPU tests of the
plumbing) -- 7*k doubles per side at most, once per k steps.  There is no collective on the data path.

Memory: only the arrays a piece needs are allocated -- Ex/Hy everywhere, psi and the CPML profiles
only where the piece touches the CPML, Dx/P/P^{n-1} only where it touches the slab -- twice (ping-pong).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat

STATE_ALL = ("Ex", "Hy", "psiE", "psiH", "Dx", "P", "Pprev")


# ------------------------------------------------------------------------------------------------ planning (pure host logic)
# relative cost of one cell-update by cell class, measured on B200 with the tile engine (vacuum 60 ps,
# Lorentz slab 131 ps per cell per 64-step block; CPML adds about the cost of a vacuum cell)
CELL_COST = {"vacuum": 1.0, "slab": 2.2, "cpml": 1.0}
# slab cell cost relative to a vacuum cell OF THE SAME KERNEL, by integrator mode, fitted to the measured 8-GPU weak-scaling
# efficiencies of tools/bench_configs.py (Lorentz: 93.1 % with 2.2, 86.7 % with 2.4 -> 2.0; Kerr-Lorentz with the Newton
# root: 88.6 % with 2.2, 88.8 % with 11.5 -> 7): a vacuum cell of the material kernels costs more than one of the
# vacuum-only kernel, so ratios taken from single-GPU rates of different kernels overestimate the slab
SLAB_COST = {"lorentz": 2.0, "lorentz_nl": 7.0, "nl": 7.0}


def balanced_rank_cuts(Lg, pw, world_size, mf=None, mr=None, slab_cost=None):
    """Rank boundaries that equalise the estimated WORK (not the cell count): the slab costs ~2.2x a
    vacuum cell, so equal-length ranges leave the vacuum-side ranks idle in every ghost exchange."""
    if mf is None or mr is None or world_size == 1:
        return [r * Lg // world_size for r in range(world_size + 1)]
    edges = sorted({0, min(pw, Lg), min(max(mf, 0), Lg), min(max(mr, 0), Lg), max(Lg - pw, 0), Lg})
    seg = []
    for a, b in zip(edges[:-1], edges[1:]):
        if b <= a:
            continue
        mid = (a + b) // 2
        w = (slab_cost or CELL_COST["slab"]) if mf <= mid < mr else CELL_COST["vacuum"]
        if mid < pw or mid >= Lg - pw:
            w += CELL_COST["cpml"]
        seg.append((a, b, w))
    total = sum((b - a) * w for a, b, w in seg)
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        acc = 0.0
        for a, b, w in seg:
            if acc + (b - a) * w >= target:
                cuts.append(int(a + (target - acc) / w))
                break
            acc += (b - a) * w
    cuts.append(Lg)
    return cuts


def plan_pieces(Lg, pw, world_size, k, max_piece=1 << 27, mf=None, mr=None, slab_cost=None):
    """Cut [0, Lg) into pieces.  Returns a list of dicts (rank, lo, hi) in global order.

    Rank boundaries balance the estimated work (``balanced_rank_cuts``; equal cell counts when the slab is
    not given).  Inside a rank the CPML zones (plus a 2k margin) become their own pieces so that the long
    interior pieces carry no CPML arrays, and pieces longer than ``max_piece`` are split (index arithmetic
    inside a piece is 32-bit)."""
    if Lg < 4 * k * world_size:
        raise ValueError("grid too short for this decomposition")
    rank_cuts = balanced_rank_cuts(Lg, pw, world_size, mf, mr, slab_cost)
    if any(b - a < 4 * k for a, b in zip(rank_cuts[:-1], rank_cuts[1:])):
        rank_cuts = [r * Lg // world_size for r in range(world_size + 1)]
    extra = [c for c in (pw + 2 * k, Lg - pw - 2 * k)
             if 0 < c < Lg and all(abs(c - rc) >= 2 * k for rc in rank_cuts)]
    if len(extra) == 2 and extra[1] - extra[0] < 2 * k:
        extra = extra[:1]
    cuts = sorted(set(rank_cuts) | set(extra))
    pieces = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        rank = max(r for r in range(world_size) if rank_cuts[r] <= lo)
        n_sub = max(1, -(-(hi - lo) // max_piece))
        if (hi - lo) // n_sub < 2 * k:
            n_sub = max(1, (hi - lo) // (2 * k))
        for i in range(n_sub):
            pieces.append(dict(rank=rank, lo=lo + (hi - lo) * i // n_sub, hi=lo + (hi - lo) * (i + 1) // n_sub))
    for i, p in enumerate(pieces):
        p["index"] = i
        p["ghost_l"] = k if i > 0 else 0
        p["ghost_r"] = k if i < len(pieces) - 1 else 0
    return pieces


def nccl_options_for_overlap():
    """Process-group options under which LongGrid.run's overlap works: NCCL's streams at high priority, so that the
    point-to-point kernels of a ghost exchange are dispatched ahead of the thousands of pending CTAs of the inner-tile
    launch instead of behind them.  Use: ``dist.init_process_group("nccl", pg_options=nccl_options_for_overlap(), ...)``."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


def exchange_schedule(pieces, rank):
    """Messages of one ghost exchange for ``rank``: a list of (kind, piece, side, peer_piece) with kind in
    {"local", "send", "recv"}; side 0/1 = left/right edge of ``piece``.  Deterministic global order."""
    sched = []
    for i in range(len(pieces) - 1):
        a, b = pieces[i], pieces[i + 1]          # boundary between a (left) and b (right)
        if a["rank"] == rank and b["rank"] == rank:
            sched.append(("local", i, 1, i + 1))
        elif a["rank"] == rank:
            sched.append(("send", i, 1, i + 1))
            sched.append(("recv", i, 1, i + 1))
        elif b["rank"] == rank:
            sched.append(("send", i + 1, 0, i))
            sched.append(("recv", i + 1, 0, i))
    return sched


def start_exchange(sched, pieces, pack, unpack, dist=None, make_buffer=None):
    """First half of a ghost exchange: same-rank boundaries are copied at once, remote messages are packed and posted
    as one batch of isend/irecv.  Returns the handle ``finish_exchange`` completes (requests + receive buffers)."""
    ops, recvs = [], []
    for kind, i, side, j in sched:
        if kind == "local":
            left, right = pack(i, 1), pack(j, 0)
            unpack(j, 0, left)
            unpack(i, 1, right)
        elif kind == "send":
            ops.append(dist.P2POp(dist.isend, pack(i, side), pieces[j]["rank"]))
        else:
            buf = make_buffer(i, side)
            recvs.append((i, side, buf))
            ops.append(dist.P2POp(dist.irecv, buf, pieces[j]["rank"]))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return reqs, recvs


def finish_exchange(handle, unpack):
    """Second half: wait for the posted messages, write the received cells into the ghost cells."""
    reqs, recvs = handle
    for req in reqs:
        req.wait()
    for i, side, buf in recvs:
        unpack(i, side, buf)


def run_exchange(sched, pieces, pack, unpack, dist=None, make_buffer=None):
    """Execute one ghost exchange.  ``pack(piece, side) -> tensor`` returns the k owned cells next to
    that edge of every state array; ``unpack(piece, side, tensor)`` writes a neighbour's packed cells
    into the ghost cells at that edge.  Remote messages are posted as one batch of isend/irecv."""
    finish_exchange(start_exchange(sched, pieces, pack, unpack, dist=dist, make_buffer=make_buffer), unpack)


# ------------------------------------------------------------------------------------------------ device side
class LongGrid:
    """A long Lorentz / dielectric / nonlinear grid decomposed into pieces on this rank's GPU."""

    def __init__(self, Lg, *, mode, pw, mf, mr, nzsrc, dz, dt, courantNo, scalars, pml_profiles, srcE, srcH,
                 probes=(), tfsf=True, k=64, rank=0, world_size=1, fma=False, fp32=False, newton=False, max_piece=1 << 27,
                 device=None):
        torch = nat.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.Lg, self.mode, self.k, self.rank, self.world = int(Lg), mode, int(k), rank, world_size
        from . import _device as dev
        self.mode_id = dev.MODE_ID[mode]
        self.pieces = plan_pieces(self.Lg, pw, world_size, self.k, max_piece, mf=mf if mode != "free" else None,
                                  mr=mr if mode != "free" else None, slab_cost=SLAB_COST.get(mode))
        self.mine = [p for p in self.pieces if p["rank"] == rank]
        self.sched = exchange_schedule(self.pieces, rank)
        T = len(srcE)
        self.T = T
        f64 = dict(dtype=torch.float64, device=self.device)
        self.src_t = torch.as_tensor(np.concatenate([np.asarray(srcE, dtype=np.float64), np.asarray(srcH, dtype=np.float64)]), **f64)
        probes = [int(p) for p in probes]
        self.probe_out = torch.zeros((max(len(probes), 1), T), **f64)
        self.probes = probes
        flags = ((nat.PF_F_TFSF if tfsf else 0) | nat.PF_F_CPML_M | nat.PF_F_CPML_P | nat.PF_F_CANONICAL |
                 (nat.PF_F_FMA if fma else 0) | (nat.PF_F_FP32 if fp32 else 0) | (nat.PF_F_NEWTON if newton else 0))
        n_arr = 7 if mode in ("lorentz", "lorentz_nl") else (5 if mode == "nl" else 4)
        self.names = STATE_ALL[:n_arr]
        self.bufs = [[], []]          # [which][piece] -> dict name -> tensor
        self.grids = [(nat.PfGrid * len(self.mine))(), (nat.PfGrid * len(self.mine))()]
        self.coefs = []
        self._keep = []
        be_l, ce_l, cm_l, be_r, ce_r, cm_r = pml_profiles      # each of length pw (left: cells 0..pw-1, right: Lg-pw..Lg-1)
        for m, p in enumerate(self.mine):
            z0 = p["lo"] - p["ghost_l"]
            L = (p["hi"] + p["ghost_r"]) - z0
            has_pml = pw > 0 and (z0 < pw or z0 + L > self.Lg - pw)
            has_slab = mode != "free" and z0 < mr and z0 + L > mf
            coef = None
            if has_pml:
                coef = {n: np.zeros(L) for n in ("beX", "ceX", "cmY")}
                gz = z0 + np.arange(L)
                left = gz < pw
                right = gz >= self.Lg - pw
                for n, lv, rv in (("beX", be_l, be_r), ("ceX", ce_l, ce_r), ("cmY", cm_l, cm_r)):
                    coef[n][left] = lv[gz[left]]
                    coef[n][right] = rv[gz[right] - (self.Lg - pw)]
                coef = {n: torch.as_tensor(v, **f64) for n, v in coef.items()}
            self.coefs.append(coef)
            own = [q for q in probes if p["lo"] <= q < p["hi"]]
            pidx = torch.tensor(own or [0], dtype=torch.int32, device=self.device)
            self._keep.append(pidx)
            for which in (0, 1):
                arrs = {}
                for n in self.names:
                    need = (n in ("Ex", "Hy")) or (n in ("psiE", "psiH") and has_pml) or (n in ("Dx", "P", "Pprev") and has_slab)
                    arrs[n] = torch.zeros(L, **f64) if need else None
                self.bufs[which].append(arrs)
                g = self.grids[which][m]
                g.L, g.pw, g.mf, g.mr, g.nzsrc = L, pw, mf, mr, nzsrc
                g.flags = flags
                g.n_probes, g.probe_stride = len(own), T
                g.n_src = T
                g.z0, g.Lg = z0, self.Lg
                for kk, v in scalars.items():
                    if hasattr(g, kk) and kk not in ("pw", "mf", "mr", "nzsrc"):
                        setattr(g, kk, float(v))
                for n in STATE_ALL:
                    t = arrs.get(n)
                    setattr(g, n, t.data_ptr() if t is not None else None)
                if coef is not None:
                    g.beX, g.ceX, g.cmY = coef["beX"].data_ptr(), coef["ceX"].data_ptr(), coef["cmY"].data_ptr()
                    g.bmY = g.beX
                g.srcE = self.src_t.data_ptr()
                g.srcH = self.src_t.data_ptr() + 8 * T
                g.probe_idx = pidx.data_ptr()
                # probe rows are global: row of probe q is its position in `probes`
                g.probe_out = self.probe_out.data_ptr()
            p["local_L"], p["z0"] = L, z0
        # probe rows: the kernel writes row p of *its own* probe list; give every piece a row map by
        # ordering its probe_out base so that its first owned probe lands in the right global row
        for m, p in enumerate(self.mine):
            own = [q for q in probes if p["lo"] <= q < p["hi"]]
            if own:
                rows = [probes.index(q) for q in own]
                if rows != list(range(rows[0], rows[0] + len(rows))):
                    raise ValueError("probes owned by one piece must be consecutive in the probe list")
                for which in (0, 1):
                    self.grids[which][m].probe_out = self.probe_out.data_ptr() + 8 * T * rows[0]
        lib = nat.lib()
        sb = lib.pf_run_block_scratch_bytes(self.grids[0], len(self.mine), self.k)
        self.scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device=self.device)
        self.scratch_bytes = sb
        self.cur = 0
        self.n_done = 0
        self._tables_built = False    # the tile tables in self.scratch were written by this object's last pf_run_block
        self.time_exchange = False    # True: every ghost exchange is bracketed by CUDA events (exchange_ms())
        self.overlap = False          # True: remote exchanges overlap the inner tiles of the block (run()); measured no gain
        self.force_split = False      # tests: split every block into inner / edge launches even without remote neighbours
        self._xch_events = []
        self.halo_bufs = {}
        self.cells_owned = sum(p["hi"] - p["lo"] for p in self.mine)
        self._local_of = {p["index"]: i for i, p in enumerate(self.mine)}

    # -- ghost exchange ------------------------------------------------------------------------
    def _buffer(self, kind, piece_index, side):
        key = (kind, piece_index, side)
        if key not in self.halo_bufs:
            self.halo_bufs[key] = self.torch.empty(len(self.names) * self.k, dtype=self.torch.float64, device=self.device)
        return self.halo_bufs[key]

    def _pack(self, piece_index, side):
        m = self._local(piece_index)
        buf = self._buffer("send", piece_index, side)
        n = nat.lib().pf_halo_pack(ctypes.byref(self.grids[self.cur][m]), self.mode_id, side, self.k, buf.data_ptr(),
                                  nat.current_stream_ptr())
        nat.check(int(min(n, 0)), "pf_halo_pack")
        return buf

    def _unpack(self, piece_index, side, buf):
        m = self._local(piece_index)
        n = nat.lib().pf_halo_unpack(ctypes.byref(self.grids[self.cur][m]), self.mode_id, side, self.k, buf.data_ptr(),
                                    nat.current_stream_ptr())
        nat.check(int(min(n, 0)), "pf_halo_unpack")

    def _local(self, piece_index):
        return self._local_of[piece_index]

    def exchange(self):
        import torch.distributed as dist
        run_exchange(self.sched, self.pieces, self._pack, self._unpack, dist=dist if self.world > 1 else None,
                     make_buffer=lambda i, side: self._buffer("recv", i, side))

    # -- time stepping -------------------------------------------------------------------------
    def run(self, nsteps, do_pol=True):
        """Advance the grid ``nsteps`` steps, k at a time.  With remote neighbours and ``overlap`` switched on a block is two
        launches: the inner tiles, which read no ghost cell, run while the messages of the exchange are in flight; the edge
        tiles follow once the ghost cells are written.  Same tiles, same arithmetic: the result does not depend on it
        (tests/test_gpu_longgrid.py, tests/test_gpu_multirank.py).  Off by default: on 2 B200s the exchange is ~0.1 ms of a
        7.5 ms block and the extra launch for the edge tiles (a 64-step kernel is >= 30 us however few tiles it has) costs what
        the hidden messages save -- 1673-1676 against 1680-1684 Gcell-updates/s, vacuum 3470-3489 against 3491-3508."""
        lib = nat.lib()
        torch = self.torch
        if self.n_done + nsteps > self.T:
            raise ValueError(f"LongGrid.run: steps {self.n_done}..{self.n_done + nsteps - 1} run past the source tables (T = {self.T})")
        import torch.distributed as dist
        remote = self.world > 1 and any(kind != "local" for kind, *_ in self.sched)
        split = (remote and self.overlap) or self.force_split

        def block(ks, part):
            # the tables hold buffer set 0 as `src`; this object owns the scratch, so after the first call they stay valid
            bflags = part
            if self._tables_built:
                bflags |= nat.PF_BLOCK_F_TABLES_VALID | (nat.PF_BLOCK_F_SWAPPED if self.cur != self._tables_src else 0)
            nat.check(lib.pf_run_block(self.grids[self.cur], self.grids[self.cur ^ 1], len(self.mine), self.mode_id,
                                       int(do_pol), self.n_done, ks, self.k, bflags, self.scratch.data_ptr(), self.scratch_bytes,
                                       nat.current_stream_ptr()), "pf_run_block")
            if not self._tables_built:
                self._tables_built, self._tables_src = True, self.cur

        done = 0
        while done < nsteps:
            ks = min(self.k, nsteps - done)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if (self.time_exchange and len(self.pieces) > 1) else None
            if len(self.pieces) > 1:
                if ev:
                    ev[0].record()
                handle = start_exchange(self.sched, self.pieces, self._pack, self._unpack, dist=dist if self.world > 1 else None,
                                        make_buffer=lambda i, side: self._buffer("recv", i, side))
                if ev:
                    ev[1].record()
                if split:
                    block(ks, nat.PF_BLOCK_F_INNER_TILES)
                if ev:
                    ev[2].record()
                finish_exchange(handle, self._unpack)
                if ev:
                    ev[3].record()
                    self._xch_events.append(ev)
            block(ks, nat.PF_BLOCK_F_EDGE_TILES if split else 0)
            self.cur ^= 1
            self.n_done += ks
            done += ks

    def exchange_ms(self):
        """Summed device time [ms] the ghost exchanges kept the compute stream busy or waiting since time_exchange was
        switched on: pack kernels + posting, and -- after the inner tiles, when the block is split -- the rest of the wait
        for the messages + unpack kernels; clears the list."""
        self.torch.cuda.synchronize()
        ms = sum(e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e in self._xch_events)
        self._xch_events = []
        return float(ms)

    def gather_owned(self, name):
        """Owned cells of one state array of this rank, concatenated in global order (host numpy)."""
        out = []
        for m, p in enumerate(self.mine):
            t = self.bufs[self.cur][m].get(name)
            L_own = p["hi"] - p["lo"]
            if t is None:
                out.append(np.zeros(L_own))
            else:
                out.append(t[p["ghost_l"]: p["ghost_l"] + L_own].cpu().numpy())
        return np.concatenate(out) if out else np.zeros(0)


def lorentz_long_grid(Lg, freq=9e9, *, T=1024, slab_fraction=0.7, k=64, rank=0, world_size=1, mode="lorentz", fma=False,
                      probes=None, max_piece=1 << 27, fp32=False, newton=False):
    """Config 5 of BASELINE.json: a grid of Lg cells with the reference's default cell size / CPML, the
    slab filling the right ``slab_fraction`` of the domain.  Returns (LongGrid, info dict)."""
    from . import BaseFDTD11, Environment_Setup as envDef, MasterController as MC, Solver_Engine as SE
    tup = envDef.envSetup(freq, 0.7, 7000, 8000, LorMed=(mode in ("lorentz", "lorentz_nl")), nonLinMed=(mode == "nl"))
    Pp = MC.Params(*tup, False, 0.7, freq, 20)
    pw = int(Pp.pmlWidth)
    # proxy grid: the CPML profiles and update scalars do not depend on the grid length
    Pp.Nz = 2 * pw + 200
    Pp.timeSteps = int(T)
    Pp.materialFrontEdge, Pp.materialRearEdge = pw + 50, Pp.Nz - 1
    Pp.TFSF, Pp.SineCont, Pp.Periods = True, True, 1000
    Pp.LorentzMed, Pp.nonLinMed, Pp.FreeSpace = mode in ("lorentz", "lorentz_nl"), mode == "nl", mode == "free"
    V = MC.Variables(Pp.Nz, Pp.timeSteps, Pp.vidInterval, 1)
    C_P = MC.CPML_Params(Pp.dz)
    C_V = MC.CPML_Variables(Pp.Nz, Pp.timeSteps)
    guards = BaseFDTD11.LIFT_SIZE_GUARDS
    BaseFDTD11.LIFT_SIZE_GUARDS = True
    try:
        V.Ex = np.zeros(Pp.Nz + 1)
        V.Hy = np.zeros(Pp.Nz + 1)
        V.UpHyMat, V.UpExMat = BaseFDTD11.EmptySpaceCalc(V, Pp)
        C_V = BaseFDTD11.CPML_FieldInit(V, Pp, C_V, C_P)
        C_V = SE.boundCondManager(V, Pp, C_V, C_P)
        Exs, Hys = SE.SourceManager(V, Pp, C_V, C_P)
    finally:
        BaseFDTD11.LIFT_SIZE_GUARDS = guards
    Lp = Pp.Nz + 1
    prof = (C_V.beX[:pw].copy(), C_V.ceX[:pw].copy(), C_V.cmY[:pw].copy(),
            C_V.beX[Lp - pw:].copy(), C_V.ceX[Lp - pw:].copy(), C_V.cmY[Lp - pw:].copy())
    mf = int(Lg * (1.0 - slab_fraction))
    mr = Lg - 2
    nzsrc = pw + int(0.05 / Pp.dz)
    Pp.materialFrontEdge, Pp.materialRearEdge, Pp.nzsrc = mf, mr, nzsrc
    scal = BaseFDTD11.grid_scalars(V, Pp, kerr_lorentz=(mode == "lorentz_nl"))
    scal.update(cE0=float(V.UpExMat[0]), cE1=float(V.UpExMat[0]), cH0=float(V.UpHyMat[0]), cH1=float(V.UpHyMat[0]),
                c2_pml=float(C_V.C2[1]))
    lg = LongGrid(Lg, mode=mode, pw=pw, mf=mf, mr=mr, nzsrc=nzsrc, dz=Pp.dz, dt=Pp.delT, courantNo=Pp.courantNo,
                  scalars=scal, pml_profiles=prof, srcE=np.asarray(Exs) / Pp.courantNo, srcH=np.asarray(Hys) / Pp.courantNo,
                  probes=probes if probes is not None else [nzsrc - 100], k=k, rank=rank, world_size=world_size, fma=fma,
                  fp32=fp32, newton=newton, max_piece=max_piece)
    info = dict(pw=pw, mf=mf, mr=mr, nzsrc=nzsrc, dz=Pp.dz, dt=Pp.delT, courantNo=Pp.courantNo, scalars=scal, profiles=prof,
                Exs=Exs, Hys=Hys, P=Pp, V=V)
    return lg, info
