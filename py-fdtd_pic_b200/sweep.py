"""Batched sweeps: many independent 1-D runs advanced together by the fused tile engine.

The reference's sweep driver is ``MasterController.LoopedSim(loop=True)`` (:533-569): 20 sequential
frequency points, each rebuilding Params/Variables and calling ``Controller``.  Members share nothing,
so here every member becomes one ``PfGrid`` and one ``pf_run_batch`` call advances all of them
(one CTA per tile of a member, k time steps per launch).  Members may differ in every parameter:
grid size, step count, dz/dt, medium, source.

Data movement per pass: ONE host->device copy (CPML profiles + source tables of all members; the
state starts from FieldInit zeros and is cleared on the device) and ONE device->host copy (probe
traces, optionally final fields).  With several GPUs members are dealt out round-robin
(``member % world_size``); there is no data-path collective.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import BaseFDTD11, Environment_Setup as envDef, Solver_Engine as SE
from . import _device as dev
from . import _native as nat

STATE_NAMES = ("Ex", "Hy", "psiE", "psiH", "Dx", "P", "Pprev", "Acubic")
COEF_NAMES = ("beX", "ceX", "cmY")


class Member:
    """Host-side description of one sweep member for one pass."""

    def __init__(self, V, P, C_V, C_P, Exs, Hys, probe_idx, nsteps=None):
        self.V, self.P, self.C_V, self.C_P = V, P, C_V, C_P
        self.L = len(V.Ex)
        self.T = int(P.timeSteps)
        self.nsteps = self.T if nsteps is None else int(nsteps)
        self.srcE = np.asarray(Exs) / P.courantNo
        self.srcH = np.asarray(Hys) / P.courantNo
        self.probe_idx = [int(p) for p in probe_idx]
        arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
        canon = dev.canonical_form(P, arrs, None)
        if canon is None or not dev.probes_ok_for_tiles(self.probe_idx):
            raise ValueError("sweep member is not in the tile engine's canonical form; run it through Controller")
        self._canon = canon
        self.scalars = None
        self.set_mode("lorentz_nl" if SE.KERR_LORENTZ else None)
        self.flags = BaseFDTD11.grid_flags(P, SE.USE_FMA, SE.USE_FP32, SE.CUBIC == "newton") | nat.PF_F_CANONICAL
        self.coef = {"beX": C_V.beX, "ceX": C_V.ceX, "cmY": C_V.cmY}


    def set_mode(self, mode):
        """Scalars of the PfGrid for integrator ``mode``: only "lorentz_nl" (the Kerr-Lorentz composition) uses its own
        cubic coefficients; MemberBatch calls this with the mode the batch actually runs, so the module-level switch
        Solver_Engine.KERR_LORENTZ cannot leak into a batch of another mode."""
        kerr = mode == "lorentz_nl"
        if self.scalars is None or kerr != self._kerr:
            canon = self._canon
            self._kerr = kerr
            self.scalars = BaseFDTD11.grid_scalars(self.V, self.P, kerr_lorentz=kerr)
            self.scalars.update(cE0=canon[0], cE1=canon[1], cH0=canon[2], cH1=canon[3], c2_pml=canon[4])


def _round32(x):
    return (np.asarray(x, dtype=np.int64) + 31) // 32 * 32


_GRID_DTYPE = np.dtype(nat.PfGrid)
_SCALAR_KEYS = ("dt_over_dz", "eps0", "polA", "polB", "polC", "cub_a", "cub_b", "cub_c", "nl_den0", "nl_den1",
                "cE0", "cE1", "cH0", "cH1", "c2_pml")


class MemberBatch:
    """Device-resident batch of members + the PfGrid array handed to pf_run_batch.

    Two ways in: a list of ``Member`` objects (each built by the per-member setup chain), or
    ``MemberBatch.from_table`` with a ``sweep_setup.MemberTable`` -- the setup of the whole sweep vectorised over members,
    with the CPML profiles and source tables written into the pinned upload buffer by one native call."""

    def __init__(self, members, mode, device=None, share_coef=None):
        for m in members:
            m.set_mode(mode)
        M = len(members)
        self.members = members
        share = np.asarray(share_coef if share_coef is not None else np.arange(M), dtype=np.int64)
        n_probes = np.array([len(m.probe_idx) for m in members], dtype=np.int64)
        pidx = np.concatenate([np.asarray(m.probe_idx or [0], dtype=np.int32) for m in members]) if M else np.zeros(0, np.int32)
        self._setup(mode, device, L=np.array([m.L for m in members], dtype=np.int64),
                    T=np.array([m.T for m in members], dtype=np.int64),
                    nsteps=np.array([m.nsteps for m in members], dtype=np.int64), n_probes=n_probes, pidx=pidx, share=share,
                    n_src=np.array([min(len(m.srcE), len(m.srcH)) for m in members], dtype=np.int64))
        # host staging: per member (the per-member chain produced the arrays)
        hv = self.host_in.numpy()
        o0 = self.n_state
        for i, m in enumerate(members):
            if share[i] == i:
                for a, name in enumerate(COEF_NAMES):
                    o = int(self.off_coef[i] + a * self.Lp[i]) - o0
                    hv[o:o + m.L] = m.coef[name]
            o = int(self.off_src[i]) - o0
            hv[o:o + len(m.srcE)] = m.srcE
            o = int(self.off_src[i] + self.Tp[i]) - o0
            hv[o:o + len(m.srcH)] = m.srcH
        scal = {k: np.array([float(m.scalars[k]) for m in members]) for k in _SCALAR_KEYS}
        geo = {k: np.array([m.scalars[k] for m in members], dtype=np.int64) for k in ("pw", "mf", "mr", "nzsrc")}
        self._fill_descriptors(geo, scal, np.array([m.flags for m in members], dtype=np.int64))

    @classmethod
    def from_table(cls, table, mode, device=None, T_alloc=None, threads=0, pinned_tag=None):
        """Batch of every row of a ``sweep_setup.MemberTable``.  T_alloc (optional, >= nsteps): rows of the source / probe
        tables to allocate instead of the member's full timeSteps (benchmarks of truncated passes)."""
        from . import sweep_setup
        self = cls.__new__(cls)
        self.members = None
        self.table = table
        T = table.T if T_alloc is None else np.full(table.n, int(T_alloc), dtype=np.int64)
        if np.any(T < table.nsteps):
            raise ValueError("MemberBatch.from_table: T_alloc is shorter than nsteps")
        n_probes = np.full(table.n, table.probes.shape[1], dtype=np.int64)
        self._setup(mode, device, L=table.L, T=T, nsteps=table.nsteps, n_probes=n_probes,
                    pidx=table.probes.astype(np.int32).reshape(-1), share=table.share, n_src=T, pinned_tag=pinned_tag)
        own = table.share == np.arange(table.n)
        o0 = self.n_state
        sweep_setup.build_inputs(table, self.host_in.numpy(), np.where(own, self.off_coef - o0, -1),
                                 np.where(own, self.off_coef + self.Lp - o0, -1),
                                 np.where(own, self.off_coef + 2 * self.Lp - o0, -1), self.off_src - o0,
                                 self.off_src + self.Tp - o0, T, threads=threads)
        self._fill_descriptors({k: getattr(table, k) for k in ("pw", "mf", "mr", "nzsrc")},
                               {k: getattr(table, k) for k in _SCALAR_KEYS}, table.flags)
        return self

    # -- layout + descriptors, vectorised over members --------------------------------------------
    def _setup(self, mode, device, *, L, T, nsteps, n_probes, pidx, share, n_src, pinned_tag=None):
        torch = nat.require_cuda()
        self.torch = torch
        self.mode = mode
        self.mode_id = dev.MODE_ID[mode]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        M = len(L)
        self.M = M
        self.L, self.T, self.nsteps_arr, self.n_probes, self.n_src = (np.asarray(x, dtype=np.int64)
                                                                       for x in (L, T, nsteps, n_probes, n_src))
        self.share = np.asarray(share, dtype=np.int64)
        self.Lp, self.Tp = _round32(self.L), _round32(self.T)
        rows = np.maximum(self.n_probes, 1)
        # layout (doubles): [state of all members][coef (owners only) + src of all members][probes of all members]
        sz_state = len(STATE_NAMES) * self.Lp
        self.off_state = np.concatenate([[0], np.cumsum(sz_state)])[:-1]
        self.n_state = int(sz_state.sum())
        own = self.share == np.arange(M)
        sz_in = np.where(own, len(COEF_NAMES) * self.Lp, 0) + 2 * self.Tp
        off_in = self.n_state + np.concatenate([[0], np.cumsum(sz_in)])[:-1]
        self.n_in = int(sz_in.sum())
        self.off_coef = off_in[self.share]                       # a sharing member points at its owner's profiles
        self.off_src = off_in + np.where(own, len(COEF_NAMES) * self.Lp, 0)
        sz_probe = rows * self.Tp
        self.off_probe = self.n_state + self.n_in + np.concatenate([[0], np.cumsum(sz_probe)])[:-1]
        self.n_probe = int(sz_probe.sum())
        self.n_total = self.n_state + self.n_in + self.n_probe
        self.pool = torch.empty(self.n_total, dtype=torch.float64, device=self.device)
        # upload staging: an own pinned buffer, or (pinned_tag) a slot of the grow-only cache in _device -- pinning hundreds of
        # megabytes costs more than filling them, so a chunked sweep reuses two slots; the trace staging buffer is only
        # allocated if traces are ever downloaded (a device-side reflection extraction never does)
        self.host_in = (torch.zeros(self.n_in, dtype=torch.float64).pin_memory() if pinned_tag is None
                        else dev.pinned_buffer(self.n_in, tag=pinned_tag))
        self._host_probe = None
        self.pidx_off = np.concatenate([[0], np.cumsum(rows)])
        self.host_pidx = torch.from_numpy(np.ascontiguousarray(pidx, dtype=np.int32)).pin_memory()
        self.pidx = torch.empty(len(pidx), dtype=torch.int32, device=self.device)
        self.grids = (nat.PfGrid * M)()
        self.nsteps = (ctypes.c_int * M)(*[int(v) for v in self.nsteps_arr])
        self.h2d_bytes = self.n_in * 8 + len(pidx) * 4
        self.d2h_bytes = self.n_probe * 8
        self.cell_steps = int((self.L * self.nsteps_arr).sum())

    def _fill_descriptors(self, geo, scal, flags):
        M = self.M
        g = np.frombuffer(self.grids, dtype=_GRID_DTYPE, count=M) if M else np.zeros(0, dtype=_GRID_DTYPE)
        base = self.pool.data_ptr()
        g["L"], g["pw"], g["mf"], g["mr"], g["nzsrc"] = self.L, geo["pw"], geo["mf"], geo["mr"], geo["nzsrc"]
        g["flags"] = flags
        g["n_probes"], g["probe_stride"], g["n_src"] = self.n_probes, self.Tp, self.n_src
        g["z0"], g["Lg"] = 0, self.L
        for k in _SCALAR_KEYS:
            g[k] = scal[k]
        for a, name in enumerate(STATE_NAMES):
            g[name] = base + 8 * (self.off_state + a * self.Lp)
        Lp_owner = self.Lp[self.share]
        for a, name in enumerate(COEF_NAMES):
            g[name] = base + 8 * (self.off_coef + a * Lp_owner)
        g["bmY"] = g["beX"]
        g["srcE"] = base + 8 * self.off_src
        g["srcH"] = base + 8 * (self.off_src + self.Tp)
        g["probe_idx"] = self.pidx.data_ptr() + 4 * self.pidx_off[:-1]
        g["probe_out"] = base + 8 * self.off_probe
        self.scratch_bytes = nat.lib().pf_run_scratch_bytes(self.grids, M, nat.PF_ENGINE_TILE)
        self.scratch = self.torch.empty(self.scratch_bytes, dtype=self.torch.uint8, device=self.device)

    @property
    def host_probe(self):
        if self._host_probe is None:
            self._host_probe = self.torch.zeros(self.n_probe, dtype=self.torch.float64).pin_memory()
        return self._host_probe

    def set_pass(self, table):
        """Re-point the descriptors at another pass of the same members (same grids and inputs; other update scalars and
        probe cells) -- e.g. pass 0 / pass 1 of the two-pass Lorentz integrator."""
        if table.n != self.M or np.any(table.L != self.L):
            raise ValueError("set_pass: the table describes other members")
        self.pidx = self.torch.as_tensor(table.probes.astype(np.int32).reshape(-1), device=self.device)
        g = np.frombuffer(self.grids, dtype=_GRID_DTYPE, count=self.M)
        for k in _SCALAR_KEYS:
            g[k] = getattr(table, k)
        g["flags"] = table.flags
        g["probe_idx"] = self.pidx.data_ptr() + 4 * self.pidx_off[:-1]
        self.table = table

    # -- device side --------------------------------------------------------------------------
    def upload(self):
        """H2D of every member's CPML profiles + source tables (one copy) and the probe indices."""
        self.pool[self.n_state:self.n_state + self.n_in].copy_(self.host_in, non_blocking=True)
        self.pidx.copy_(self.host_pidx, non_blocking=True)

    def reset_state(self, template=False):
        """FieldInit / CPML_FieldInit on the device: zero fields, polarisation, psi and probe traces.
        template=True starts from the synthetic state made by randomize_state() instead of zeros."""
        if template:
            self.pool[: self.n_state].copy_(self.state_template)
        else:
            self.pool[: self.n_state].zero_()
        self.pool[self.n_state + self.n_in:].zero_()

    # physically plausible magnitudes: E ~ 1 V/m, H ~ E/377, D and P ~ eps0*E, psi_E ~ 1e-7, psi_H ~ 1
    STATE_SCALE = {"Ex": 1.0, "Hy": 1.0 / 376.73, "psiE": 1e-7, "psiH": 1.0, "Dx": 8.85e-12, "P": 8.85e-12,
                   "Pprev": 8.85e-12, "Acubic": 0.0}

    def randomize_state(self, seed=1234):
        """Synthetic non-zero state for benchmarks: every cell of every state array gets a uniform
        random value of the field's natural magnitude, so no kernel path is favoured by zeros."""
        torch = self.torch
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed)
        self.state_template = torch.rand(self.n_state, dtype=torch.float64, device=self.device, generator=gen) * 2 - 1
        scale = torch.tensor([self.STATE_SCALE[name] for name in STATE_NAMES], dtype=torch.float64,
                             device=self.device)[:, None]
        for i in range(self.M):      # a member's state arrays are one contiguous [8, Lp] block
            Lp = int(self.Lp[i])
            o = int(self.off_state[i])
            self.state_template[o:o + len(STATE_NAMES) * Lp].view(len(STATE_NAMES), Lp).mul_(scale)
        return self.state_template

    def run(self, do_pol, n0=0, k_block=0):
        nat.check(nat.lib().pf_run_batch(self.grids, self.M, self.mode_id, int(do_pol), int(n0),
                                         self.nsteps, int(k_block), self.scratch.data_ptr(), self.scratch_bytes,
                                         nat.current_stream_ptr()), "pf_run_batch")

    def _unpack_probes(self, hv):
        rows = np.maximum(self.n_probes, 1)
        if self.M and len({(int(r), int(p), int(t), int(T)) for r, p, t, T in zip(rows, self.n_probes, self.Tp, self.T)}) == 1:
            # uniform batch (the sweep case): one strided copy instead of one per member
            return list(hv.reshape(self.M, int(rows[0]), int(self.Tp[0]))[:, :int(self.n_probes[0]), :int(self.T[0])].copy())
        out = []
        o0 = self.n_state + self.n_in
        for i in range(self.M):
            Tp, n_p, r = int(self.Tp[i]), int(self.n_probes[i]), int(rows[i])
            a = hv[int(self.off_probe[i]) - o0: int(self.off_probe[i]) - o0 + r * Tp]
            out.append(a.reshape(r, Tp)[:n_p, : int(self.T[i])].copy())
        return out

    def download_probes(self):
        """D2H of all probe traces (one copy) -> list of arrays [n_probes, T] per member."""
        self.host_probe.copy_(self.pool[self.n_state + self.n_in:], non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        return self._unpack_probes(self.host_probe.numpy())

    def download_probes_async(self):
        """Stream-ordered D2H of all probe traces into this batch's pinned buffer; returns a function that
        waits for the copy and unpacks it like download_probes() (call it after other work was queued)."""
        torch = self.torch
        self.host_probe.copy_(self.pool[self.n_state + self.n_in:], non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())

        def finish():
            done.synchronize()
            return self._unpack_probes(self.host_probe.numpy())
        return finish

    def download_probe(self, i):
        """D2H of the probe traces of member i only -> array [n_probes, T]."""
        Tp, n_p = int(self.Tp[i]), int(self.n_probes[i])
        o = int(self.off_probe[i])
        a = self.pool[o: o + max(n_p, 1) * Tp].cpu().numpy()
        return a.reshape(max(n_p, 1), Tp)[:n_p, : int(self.T[i])].copy()

    def probe_rows(self, idxs, probe, T):
        """Device tensor [len(idxs), T] of probe row ``probe`` of the given members (which share T): a strided view of
        the pool when the members are evenly spaced (the uniform sweep), a gather otherwise."""
        torch = self.torch
        Tp = (T + 31) // 32 * 32
        offs = self.off_probe[np.asarray(idxs)] + probe * Tp
        if len(offs) > 1 and np.all(np.diff(offs) == offs[1] - offs[0]):
            return torch.as_strided(self.pool, (len(offs), T), (int(offs[1] - offs[0]), 1), int(offs[0]))
        return torch.stack([self.pool[int(o): int(o) + T] for o in offs])

    def probe_peaks(self, probe=0, keep_from=None, keep_to=None, on_device=False):
        """RefTester's scalar of probe ``probe`` of every member, computed on the device (batched FFT per
        distinct timeSteps; nothing but the result, one double per member, is copied to the host).
        keep_from / keep_to: per-member first / last sample index of the read window (others are zeroed)."""
        from . import TransformHandler as transH
        torch = self.torch
        kf = None if keep_from is None else np.asarray(keep_from, dtype=np.int64)
        kt = None if keep_to is None else np.asarray(keep_to, dtype=np.int64)
        out = torch.empty(self.M, dtype=torch.float64, device=self.device)
        bins = torch.empty(self.M, dtype=torch.int64, device=self.device)
        for T in np.unique(self.T):
            idxs = np.flatnonzero(self.T == T)
            rows = self.probe_rows(idxs, probe, int(T))
            val, idx = transH.ref_tester_batch(rows, int(T), None if kf is None else kf[idxs], None if kt is None else kt[idxs],
                                               check=not on_device)
            sel = torch.as_tensor(idxs, device=self.device)
            out[sel] = val
            bins[sel] = idx
        if on_device:        # no host synchronisation: (values, peak bins) stay on the device; the caller checks bins != 0
            return out, bins
        return out.cpu().numpy()

    def state(self, i, name):
        """Device -> host copy of one state array of member i (tests / final fields)."""
        o = int(self.off_state[i] + STATE_NAMES.index(name) * self.Lp[i])
        return self.pool[o:o + int(self.L[i])].cpu().numpy()


class BatchPipeline:
    """A sweep larger than one batch (e.g. 4096 members in batches of 1024): the passes of successive
    MemberBatch objects with the host-to-device copy of batch i+1's inputs and the device-to-host copy of
    batch i-1's probe traces overlapped with the time stepping of batch i (three CUDA streams, events
    between them; every batch owns its pinned staging buffers and its device pool, so nothing is shared).
    A batch may appear several times in ``order`` (new inputs each time it is uploaded)."""

    def __init__(self, batches):
        torch = nat.require_cuda()
        self.torch = torch
        self.batches = list(batches)
        self.s_in, self.s_run, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def run(self, order, do_pol, *, template=False, k_block=0, on_result=None):
        """Process batches[order[0]], batches[order[1]], ...; returns the list of probe-trace lists (or the
        values of on_result(index_in_order, traces) if given)."""
        torch = self.torch
        cur = torch.cuda.current_stream()
        for st in (self.s_in, self.s_run, self.s_out):
            st.wait_stream(cur)
        free = {}            # batch index -> event after which its pool / pinned input may be overwritten
        pending, results = [], []

        def drain(keep, batch=None):
            # finish() copies out of the batch's pinned buffer: results are consumed in order, and always before the
            # same batch's next device-to-host copy is queued
            while len(pending) > keep or (batch is not None and any(p[2] == batch for p in pending)):
                j, fin, _ = pending.pop(0)
                tr = fin()
                results.append(on_result(j, tr) if on_result else tr)

        for j, bi in enumerate(order):
            b = self.batches[bi]
            with torch.cuda.stream(self.s_in):
                if bi in free:
                    self.s_in.wait_event(free[bi])     # the previous use of this batch has left the device pool
                b.upload()
                up = torch.cuda.Event()
                up.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(up)
                b.reset_state(template=template)
                b.run(do_pol=do_pol, k_block=k_block)
                ran = torch.cuda.Event()
                ran.record(self.s_run)
            drain(keep=len(pending), batch=bi)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ran)
                fin = b.download_probes_async()
                out_done = torch.cuda.Event()
                out_done.record(self.s_out)
            free[bi] = out_done
            pending.append((j, fin, bi))
            drain(keep=1)          # unpack the previous batch on the host while this one runs
        drain(keep=0)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        return results


# ------------------------------------------------------------------------------------------------
def new_member_objects(freq_in, domainSize, lowLimTim, highLimTim, prevV, prevP, template_P):
    """One iteration of the reference sweep's object construction (MasterController.py:545-551):
    the grid is sized from the PREVIOUS member's (dispersion-corrected) medium."""
    from . import MasterController as MC
    import types
    # the reference has already advanced prevP.freq_in by Interval when it sizes the next member
    # (MasterController.py:559); the previous member's own P must not change before it has run
    probe_P = types.SimpleNamespace(freq_in=freq_in, permit_0=prevP.permit_0)
    tup = envDef.envSetup(freq_in, domainSize, lowLimTim, highLimTim, VExists=True, V=prevV, P=probe_P)
    P = MC.Params(*tup, template_P.MORmode, domainSize, freq_in, 20, LorentzMed=template_P.LorentzMed,
                  SineCont=template_P.SineCont, Gaussian=template_P.Gaussian, TFSF=template_P.TFSF)
    V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
    C_P = MC.CPML_Params(P.dz)
    C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
    return V, P, C_V, C_P


def run_two_pass_batch(objs, lorentz=True, k_block=0, rank=0, world_size=1, device_reflection=False):
    """Run the two-pass (incident / with medium) integrator for every (V,P,C_V,C_P) in ``objs`` as a
    batch.  Fills V.x1ColBe / V.x1ColAf and the final fields of every member owned by this rank
    (member % world_size == rank).  Returns (owned member indices, their source tables, reflection).

    device_reflection=True: the reflection coefficient of every owned member (``results(RefCo=True)``:
    RefTester peak of the windowed x1ColAf over that of x1ColBe) is computed on the device and returned as
    an array; only the LAST owned member's traces are copied back into its V (the reference's sweep never
    looks at the others again).  Otherwise reflection is None and every member's traces are downloaded."""
    mode = ("lorentz_nl" if SE.KERR_LORENTZ else "lorentz") if lorentz else "free"
    mine = [i for i in range(len(objs)) if i % world_size == rank]
    srcs = {}
    peaks = [None, None]
    for pass_idx in range(2):
        members = []
        for i in mine:
            V, P, C_V, C_P = objs[i]
            C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz)
            objs[i] = (V, P, C_V, C_P)
            srcs[i] = (Exs, Hys)
            members.append(Member(V, P, C_V, C_P, Exs, Hys, [P.x1Loc if pass_idx == 0 else P.x2Loc]))
        if not members:
            continue
        batch = MemberBatch(members, mode)
        batch.upload()
        batch.reset_state()
        batch.run(do_pol=(lorentz and pass_idx == 1), k_block=k_block)
        fin_be = [int(objs[i][1].timeSteps * 0.7) for i in mine]        # Solver_Engine.py:360-368 read windows
        start_af = [int(objs[i][1].timeSteps * 0.05) for i in mine]
        if device_reflection:
            peaks[pass_idx] = (batch.probe_peaks(keep_to=fin_be) if pass_idx == 0
                               else batch.probe_peaks(keep_from=start_af))
            fetch = [len(mine) - 1]
            traces = {fetch[0]: batch.download_probe(fetch[0])}
        else:
            fetch = range(len(mine))
            traces = dict(enumerate(batch.download_probes()))
        for j in fetch:
            V, P, C_V, C_P = objs[mine[j]]
            n = np.arange(P.timeSteps)
            if pass_idx == 0:
                V.x1ColBe = np.where(n <= fin_be[j], traces[j][0], 0.0)
            else:
                V.x1ColAf = np.where(n >= start_af[j], traces[j][0], 0.0)
                V.Ex, V.Hy = batch.state(j, "Ex"), batch.state(j, "Hy")
    refl = peaks[1] / peaks[0] if device_reflection and mine else None
    return mine, srcs, refl


def nonlinear_sweep_members(freqs, amplitudes, domainSize, lowLimTim, highLimTim, *, nsteps=None, rank=0, world_size=1):
    """Host-side half of nonlinear_sweep (no device needed): the frequency-major (f, amp) list, the indices this rank owns
    (index % world_size == rank), their Member objects (one setup chain per distinct frequency, sources scaled by the
    amplitude) and, per member, the position of the member whose CPML profiles it shares."""
    from . import MasterController as MC
    pairs = [(float(f), float(a)) for f in freqs for a in amplitudes]
    mine = [i for i in range(len(pairs)) if i % world_size == rank]
    base, members, share = {}, [], []
    for j, i in enumerate(mine):
        f, amp = pairs[i]
        if f not in base:
            tup = envDef.envSetup(f, domainSize, lowLimTim, highLimTim, nonLinMed=True)
            P = MC.Params(*tup, False, domainSize, f, 20)
            P.TFSF, P.SineCont, P.Periods, P.nonLinMed, P.FreeSpace, P.LorentzMed = True, True, 1000, True, False, False
            V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
            C_P = MC.CPML_Params(P.dz)
            C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
            C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False, nonlinear=True)
            base[f] = (V, P, C_V, C_P, np.asarray(Exs), np.asarray(Hys), j)
        V, P, C_V, C_P, Exs, Hys, first = base[f]
        members.append(Member(V, P, C_V, C_P, Exs * amp, Hys * amp, [P.materialFrontEdge, P.materialRearEdge], nsteps=nsteps))
        share.append(first)
    return pairs, mine, members, share


def nonlinear_sweep(freqs, amplitudes, domainSize, lowLimTim, highLimTim, *, nsteps=None, rank=0, world_size=1,
                    harmonics=(1, 3), download_traces=False, k_block=0):
    """BASELINE config 3: the cubic nonlinear integrator (Solver_Engine.IntegratorNL1D, Solver_Engine.py:220-271) over
    every (frequency, amplitude) pair as ONE batch -- the amplitude scales the member's source tables Exs / Hys, the
    grid is the one ``envSetup(f, ..., nonLinMed=True)`` gives for its frequency (members of one frequency share their
    CPML profiles).  Members are dealt round-robin over ``world_size`` ranks (member % world_size == rank), no collective.

    Returns a dict for the members this rank owns: ``index`` (position in the frequency-major (f, amp) list), ``freq``,
    ``amp``, ``harmonic_amplitude`` [n_owned, len(harmonics), 2] = 2|FFT|/T of Port1 (slab front) and Port2 (slab rear) at
    the bins nearest k*f -- computed on the device (batched FFT per distinct timeSteps; what the reference's missing
    ``transH.CZT`` call, MasterController.py:669, was after), and ``Port1`` / ``Port2`` traces if ``download_traces``."""
    torch = nat.require_cuda()
    pairs, mine, members, share = nonlinear_sweep_members(freqs, amplitudes, domainSize, lowLimTim, highLimTim, nsteps=nsteps,
                                                          rank=rank, world_size=world_size)
    out = dict(index=np.asarray(mine, dtype=np.int64), freq=np.asarray([pairs[i][0] for i in mine]),
               amp=np.asarray([pairs[i][1] for i in mine]))
    if not members:
        out["harmonic_amplitude"] = np.zeros((0, len(harmonics), 2))
        return out
    batch = MemberBatch(members, "nl", share_coef=share)
    batch.upload()
    batch.reset_state()
    batch.run(do_pol=False, k_block=k_block)
    # harmonic content of the two port traces, on the device
    groups = {}
    for j, m in enumerate(members):
        groups.setdefault(m.T, []).append(j)
    H = torch.zeros((len(members), len(harmonics), 2), dtype=torch.float64, device=batch.device)
    for T, idxs in groups.items():
        Tp = (T + 31) // 32 * 32
        rows = torch.stack([batch.pool[batch.off_probe[j]: batch.off_probe[j] + 2 * Tp].view(2, Tp)[:, :T] for j in idxs])
        spec = torch.fft.rfft(rows, dim=-1).abs() * (2.0 / T)                     # [n, 2, T//2+1]
        dt = torch.tensor([members[j].P.delT for j in idxs], dtype=torch.float64, device=batch.device)
        f0 = torch.tensor([members[j].P.freq_in for j in idxs], dtype=torch.float64, device=batch.device)
        for h, k in enumerate(harmonics):
            bins = torch.clamp(torch.round(k * f0 * dt * T).long(), 0, spec.shape[-1] - 1)      # nearest bin to k*f
            H[torch.as_tensor(idxs, device=batch.device), h] = spec[torch.arange(len(idxs), device=batch.device), :, bins]
    out["harmonic_amplitude"] = H.cpu().numpy()
    if download_traces:
        tr = batch.download_probes()
        out["Port1"] = [t[0] for t in tr]
        out["Port2"] = [t[1] for t in tr]
    out["cell_steps"] = batch.cell_steps
    return out


def _sweep_chain_env(V, P, domainSize, lowLimTim, highLimTim, Interval, points):
    """The sizing chain of the reference's sweep (MasterController.py:543-563) without the objects: member i's points per
    wavelength come from member i-1's medium AS ITS TWO PASSES LEAVE IT (twice dispersion-corrected plasma frequency; the
    first member is sized from the V that was passed in), every member then starts from the default medium again.
    Returns (freqs, env dict of arrays, twice-corrected plasma frequency per member)."""
    from . import sweep_setup, genericStability as gStab
    med = sweep_setup.default_medium(1)
    wp0, w0d, gamd = float(med["wp"][0]), float(med["w0"][0]), float(med["gam"][0])
    prev = (float(V.plasmaFreqE), float(V.omega_0E), float(V.gammaE))
    freq = float(P.freq_in)
    freqs, envs, wps = [], [], []
    for _ in range(points):
        w = 2 * np.pi * freq                                        # Environment_Setup.py:23-46 with VExists=True
        eps = 1 + (prev[0] * prev[0]) / (prev[1] * prev[1] - (w * w) + 1j * prev[2] * w)
        Nlam = int(60 * (np.real(eps)) ** 1.05)
        env = sweep_setup.envSetup_many(np.array([freq]), domainSize, lowLimTim, highLimTim, LorMed=True, Nlam=np.array([Nlam]))
        wp = wp0
        for _k in range(2):                                         # Solver_Engine.py:286, once per pass
            wp = float(gStab.spatialStab(int(env["timeSteps"][0]), int(env["Nz"][0]), float(env["dz"][0]), freq,
                                         float(env["delT"][0]), wp, w0d, gamd)[3])
        freqs.append(freq)
        envs.append(env)
        wps.append(wp)
        prev = (wp, w0d, gamd)
        freq = freq + Interval
    env = {k: np.concatenate([e[k] for e in envs]) for k in envs[0]}
    return np.array(freqs), env, np.array(wps)


def frequency_sweep(V, P, domainSize, lowLimTim, highLimTim, Low=3e9, Interval=1e8, points=20, batched=True,
                    device_postproc=True):
    """MasterController.LoopedSim(loop=True) :533-569.  Returns (freqs, measured R, analytical R,
    (V,P,C_V,C_P,Exs,Hys) of the last member).

    batched + device_postproc (default) with a Lorentz medium and the sine source: the whole sweep runs through the
    vectorised path -- the reference's sizing chain on scalars (``_sweep_chain_env``), tables, native input building, one
    batch, reflection extraction on the device; only the LAST member is also built as objects (the tuple the reference
    returns), with its traces and final fields copied back.  Otherwise: per-member objects, batched (``run_two_pass_batch``)
    or one Controller call after another like the reference."""
    from . import MasterController as MC
    freqs = np.arange(Low, points * Interval + Low, Interval)[:points]
    fast = (batched and device_postproc and bool(P.LorentzMed) and bool(P.SineCont) and not bool(P.Gaussian)
            and not SE.KERR_LORENTZ and points > 0)
    if fast:
        from . import sweep_setup
        _, env, wp2 = _sweep_chain_env(V, P, domainSize, lowLimTim, highLimTim, Interval, points)
        fr = np.empty(points)                       # member frequencies as the reference accumulates them (freq_in += Interval)
        f = float(P.freq_in)
        for i in range(points):
            fr[i] = f
            f = f + Interval
        tables = sweep_setup.lorentz_sweep_tables(fr, 1.0, domainSize, lowLimTim, highLimTim, periods=P.Periods,
                                                  tfsf=bool(P.TFSF), fma=SE.USE_FMA, fp32=SE.USE_FP32, env=env)
        assert np.array_equal(tables[1].wp, wp2)
        measured, _, _, last = _run_reflection_tables(tables, np.arange(points), keep_last=True)
        analytical = _analytical_reflection(fr, wp2)
        # the last member as objects: sized from the previous member's medium exactly like the chain above
        prevV, prevP = V, P
        if points > 1:
            prevV = MC.Variables(1, 1, 1, 1)
            prevV.plasmaFreqE = float(wp2[points - 2])
        Vi, Pi, CVi, CPi = new_member_objects(float(fr[-1]), domainSize, lowLimTim, highLimTim, prevV, prevP, P)
        for _k in range(2):
            CVi, Exs, Hys = SE.prepare_pass(Vi, Pi, CVi, CPi, True)
        n = np.arange(Pi.timeSteps)
        Vi.x1ColBe = np.where(n <= int(Pi.timeSteps * 0.7), last[0], 0.0)
        Vi.x1ColAf = np.where(n >= int(Pi.timeSteps * 0.05), last[1], 0.0)
        Vi.Ex, Vi.Hy = last[2], last[3]
        Pi.freq_in = Pi.freq_in + Interval          # the reference leaves the last P advanced (:559)
        return freqs, measured, analytical, (Vi, Pi, CVi, CPi, Exs, Hys)
    # -- setup chain (sequential by construction: member i is sized from member i-1's corrected medium;
    #    the correction itself is host-side setup, so it does not need member i-1's fields)
    objs = []
    prevV, prevP = V, P
    freq = P.freq_in
    for _ in range(points):
        Vi, Pi, CVi, CPi = new_member_objects(freq, domainSize, lowLimTim, highLimTim, prevV, prevP, P)
        objs.append((Vi, Pi, CVi, CPi))
        if batched:
            # what Controller will leave in V.plasmaFreqE after its two passes (Solver_Engine.py:286)
            shadow = MC.Variables(Pi.Nz, 1, 1, 1)
            wp = Vi.plasmaFreqE
            if Pi.LorentzMed:
                from . import genericStability as gStab
                for _k in range(2):
                    wp = gStab.spatialStab(Pi.timeSteps, Pi.Nz, Pi.dz, Pi.freq_in, Pi.delT, wp, Vi.omega_0E, Vi.gammaE)[3]
            shadow.plasmaFreqE = wp
            prevV, prevP = shadow, Pi
        else:
            MC.Controller(Vi, Pi, CVi, CPi)
            prevV, prevP = Vi, Pi
        freq = Pi.freq_in + Interval
    srcs, refl = {}, None
    if batched:
        _, srcs, refl = run_two_pass_batch(objs, lorentz=bool(P.LorentzMed), k_block=0,
                                           device_reflection=bool(device_postproc))
    measured = np.zeros(points)
    analytical = np.zeros(points)
    for i, (Vi, Pi, CVi, CPi) in enumerate(objs):
        t = np.arange(0, len(Vi.x1ColBe)) * Pi.delT
        measured[i] = refl[i] if refl is not None else MC.results(Vi, Pi, CVi, CPi, t, RefCo=True)
        analytical[i] = MC.results(Vi, Pi, CVi, CPi, t, AnalRefCo=True)
    Vi, Pi, CVi, CPi = objs[-1]
    Pi.freq_in = Pi.freq_in + Interval          # the reference leaves the last P advanced (:559)
    Exs, Hys = srcs.get(points - 1, (None, None))
    return freqs, measured, analytical, (Vi, Pi, CVi, CPi, Exs, Hys)


def _run_reflection_tables(tables, mine, *, chunk=256, k_block=0, threads=0, keep_last=False):
    """Two passes of IntegratorLinLor1D + results(RefCo=True) for the rows ``mine`` of a (pass 0, pass 1) table pair, chunk by
    chunk, pipelined (while the GPU steps chunk i the host builds the inputs of chunk i+1).  Returns (R per member,
    cell-updates, seconds spent building inputs, and -- keep_last -- the last member's (x1ColBe trace, x1ColAf trace, Ex, Hy))."""
    import time
    torch = nat.require_cuda()
    t_build = 0.0
    out_R, out_bins, uploaded, cell_steps, keep, last = [], [], {}, 0, [], None
    for c, lo in enumerate(range(0, len(mine), chunk)):
        idx = mine[lo:lo + chunk]
        t0c, t1c = tables[0].select(idx), tables[1].select(idx)
        slot = f"sweep_in_{c % 2}"
        if slot in uploaded:
            uploaded[slot].synchronize()           # the H2D copy that last read this pinned slot has finished
        tb = time.perf_counter()
        batch = MemberBatch.from_table(t0c, "lorentz", pinned_tag=slot, threads=threads)
        t_build += time.perf_counter() - tb
        batch.upload()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        uploaded[slot] = ev
        T = batch.T
        is_last = keep_last and lo + chunk >= len(mine)
        batch.reset_state()
        batch.run(do_pol=False, k_block=k_block)
        p0, b0 = batch.probe_peaks(keep_to=np.trunc(T * 0.7).astype(np.int64), on_device=True)
        tr0 = batch.download_probe(batch.M - 1)[0] if is_last else None
        batch.set_pass(t1c)
        batch.reset_state()
        batch.run(do_pol=True, k_block=k_block)
        p1, b1 = batch.probe_peaks(keep_from=np.trunc(T * 0.05).astype(np.int64), on_device=True)
        if is_last:
            last = (tr0, batch.download_probe(batch.M - 1)[0], batch.state(batch.M - 1, "Ex"), batch.state(batch.M - 1, "Hy"))
        out_R.append(p1 / p0)
        out_bins.append(torch.minimum(b0, b1))
        cell_steps += 2 * batch.cell_steps
        keep.append(batch)                          # descriptors / pools stay alive until the stream has drained
    if out_R:
        R = torch.cat(out_R).cpu().numpy()
        if bool((torch.cat(out_bins) == 0).any()):
            raise ValueError("Could not find non-DC freq")
    else:
        R = np.zeros(0)
    del keep
    return R, cell_steps, t_build, last


def _analytical_reflection(freqs, wp):
    """BaseFDTD11.AnalyticalReflectionE per member (Python-float / Python-complex arithmetic as there), default medium."""
    from . import sweep_setup
    med = sweep_setup.default_medium(1)
    w0, gam = float(med["w0"][0]), float(med["gam"][0])
    out = np.empty(len(freqs))
    for j in range(len(freqs)):
        wpj, wj = float(wp[j]), 2 * np.pi * float(freqs[j])
        eps = 1 + (wpj * wpj) / (w0 * w0 - (wj * wj) + 1j * gam * wj)
        n2 = np.real(np.sqrt(eps))
        out[j] = abs((n2 - 1) / (1 + n2))
    return out


def reflection_sweep(freqs, domainSize, lowLimTim, highLimTim, *, amps=1.0, periods=1.0, tfsf=True, chunk=256, rank=0,
                     world_size=1, k_block=0, fma=False, fp32=False, nsteps=None, threads=0):
    """Reflection coefficient of the Lorentz half-space at every frequency of ``freqs``: what the reference's sweep
    (MasterController.LoopedSim, :543-563) computes per point -- ``Controller`` (two passes of IntegratorLinLor1D) and
    ``results(RefCo=True)`` -- for grids ``envSetup(f, domainSize, lowLimTim, highLimTim, LorMed=True)``.

    Everything per member is batched: the host setup is vectorised (sweep_setup), the CPML profiles and source tables are
    built natively straight into pinned memory, all members of a chunk advance together (pf_run_batch), the reflection
    figure is extracted on the device (batched FFT).  Chunks are pipelined: while the GPU steps chunk i the host builds the
    inputs of chunk i+1.  Members are dealt round-robin over ranks (member % world_size == rank); no collective.

    Returns a dict: ``index`` (positions in ``freqs`` this rank owns), ``freq``, ``measured`` (R per owned member),
    ``analytical`` (BaseFDTD11.AnalyticalReflectionE with the medium as Controller leaves it), ``cell_steps`` and
    ``timing`` (seconds: vectorised setup, native input building, total wall)."""
    import time
    from . import sweep_setup
    nat.require_cuda()
    if threads == 0 and world_size > 1:
        import os
        threads = max(1, (os.cpu_count() or 1) // world_size)      # one process per GPU on one host: share its cores
    t_start = time.perf_counter()
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    mine = np.arange(rank, len(freqs), world_size)
    amps_mine = np.broadcast_to(np.asarray(amps, dtype=np.float64), freqs.shape)[mine]
    # (only this rank's members are set up: the host chain is per member, nothing is shared between ranks)
    tables = sweep_setup.lorentz_sweep_tables(freqs[mine], amps_mine, domainSize, lowLimTim, highLimTim, periods=periods, tfsf=tfsf,
                                              nsteps=nsteps, fma=fma, fp32=fp32)
    t_setup = time.perf_counter() - t_start
    analytical = _analytical_reflection(freqs[mine], tables[1].wp)     # medium as Controller leaves it
    R, cell_steps, t_build, _ = _run_reflection_tables(tables, np.arange(len(mine)), chunk=chunk, k_block=k_block, threads=threads)
    return dict(index=mine, freq=freqs[mine], measured=R, analytical=analytical, cell_steps=cell_steps,
                timing=dict(setup_s=t_setup, build_inputs_s=t_build, total_s=time.perf_counter() - t_start))
