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


class MemberBatch:
    """Device-resident batch of members + the PfGrid array handed to pf_run_batch."""

    def __init__(self, members, mode, device=None, share_coef=None):
        torch = nat.require_cuda()
        self.torch = torch
        self.members = members
        self.mode = mode
        self.mode_id = dev.MODE_ID[mode]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        for m in members:
            m.set_mode(mode)
        M = len(members)
        # ---- layout (doubles): [state of all members][coef+src of all members][probes of all members]
        self.off_state, self.off_in, self.off_probe = [], [], []
        cur = 0
        for m in members:
            Lp = (m.L + 31) // 32 * 32
            self.off_state.append(cur)
            cur += len(STATE_NAMES) * Lp
        self.n_state = cur
        share_coef = share_coef or list(range(M))      # member -> index of the member whose coef it reuses
        self.share = share_coef
        for i, m in enumerate(members):
            Lp = (m.L + 31) // 32 * 32
            Tp = (m.T + 31) // 32 * 32
            self.off_in.append(cur)
            cur += (len(COEF_NAMES) * Lp if share_coef[i] == i else 0) + 2 * Tp
        self.n_in = cur - self.n_state
        for m in members:
            Tp = (m.T + 31) // 32 * 32
            self.off_probe.append(cur)
            cur += max(len(m.probe_idx), 1) * Tp
        self.n_probe = cur - self.n_state - self.n_in
        self.n_total = cur
        self.pool = torch.empty(cur, dtype=torch.float64, device=self.device)
        self.host_in = torch.zeros(self.n_in, dtype=torch.float64).pin_memory()
        self.host_probe = torch.zeros(self.n_probe, dtype=torch.float64).pin_memory()
        pidx = np.concatenate([np.asarray(m.probe_idx or [0], dtype=np.int32) for m in members])
        self.pidx_off = np.concatenate([[0], np.cumsum([max(len(m.probe_idx), 1) for m in members])])
        self.host_pidx = torch.from_numpy(pidx).pin_memory()
        self.pidx = torch.empty(len(pidx), dtype=torch.int32, device=self.device)
        self.grids = (nat.PfGrid * M)()
        self.nsteps = (ctypes.c_int * M)(*[m.nsteps for m in members])
        self._fill_host_inputs()
        self._fill_descriptors()
        self.scratch_bytes = nat.lib().pf_run_scratch_bytes(self.grids, M, nat.PF_ENGINE_TILE)
        self.scratch = torch.empty(self.scratch_bytes, dtype=torch.uint8, device=self.device)
        self.h2d_bytes = self.n_in * 8 + len(pidx) * 4
        self.d2h_bytes = self.n_probe * 8
        self.cell_steps = sum(m.L * m.nsteps for m in members)

    # -- host staging -------------------------------------------------------------------------
    def _coef_offset(self, i, name):
        j = self.share[i]
        Lp = (self.members[j].L + 31) // 32 * 32
        return self.off_in[j] + COEF_NAMES.index(name) * Lp

    def _src_offset(self, i, which):
        m = self.members[i]
        Lp = (m.L + 31) // 32 * 32
        Tp = (m.T + 31) // 32 * 32
        base = self.off_in[i] + (len(COEF_NAMES) * Lp if self.share[i] == i else 0)
        return base + which * Tp

    def _fill_host_inputs(self):
        hv = self.host_in.numpy()
        o0 = self.n_state
        for i, m in enumerate(self.members):
            if self.share[i] == i:
                for name in COEF_NAMES:
                    o = self._coef_offset(i, name) - o0
                    hv[o:o + m.L] = m.coef[name]
            o = self._src_offset(i, 0) - o0
            hv[o:o + len(m.srcE)] = m.srcE
            o = self._src_offset(i, 1) - o0
            hv[o:o + len(m.srcH)] = m.srcH

    def _fill_descriptors(self):
        base = self.pool.data_ptr()
        for i, m in enumerate(self.members):
            g = self.grids[i]
            Lp = (m.L + 31) // 32 * 32
            Tp = (m.T + 31) // 32 * 32
            s = m.scalars
            g.L, g.pw, g.mf, g.mr, g.nzsrc = m.L, s["pw"], s["mf"], s["mr"], s["nzsrc"]
            g.flags = m.flags
            g.n_probes, g.probe_stride = len(m.probe_idx), Tp
            g.n_src = min(len(m.srcE), len(m.srcH))
            g.z0, g.Lg = 0, m.L
            for k in ("dt_over_dz", "eps0", "polA", "polB", "polC", "cub_a", "cub_b", "cub_c", "nl_den0", "nl_den1",
                      "cE0", "cE1", "cH0", "cH1", "c2_pml"):
                setattr(g, k, float(s[k]))
            for a, name in enumerate(STATE_NAMES):
                setattr(g, name, base + 8 * (self.off_state[i] + a * Lp))
            for name in COEF_NAMES:
                setattr(g, name, base + 8 * self._coef_offset(i, name))
            g.bmY = g.beX
            g.srcE = base + 8 * self._src_offset(i, 0)
            g.srcH = base + 8 * self._src_offset(i, 1)
            g.probe_idx = self.pidx.data_ptr() + 4 * int(self.pidx_off[i])
            g.probe_out = base + 8 * self.off_probe[i]

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
        for i, m in enumerate(self.members):      # a member's state arrays are one contiguous [8, Lp] block
            Lp = (m.L + 31) // 32 * 32
            o = self.off_state[i]
            self.state_template[o:o + len(STATE_NAMES) * Lp].view(len(STATE_NAMES), Lp).mul_(scale)
        return self.state_template

    def run(self, do_pol, n0=0, k_block=0):
        nat.check(nat.lib().pf_run_batch(self.grids, len(self.members), self.mode_id, int(do_pol), int(n0),
                                         self.nsteps, int(k_block), self.scratch.data_ptr(), self.scratch_bytes,
                                         nat.current_stream_ptr()), "pf_run_batch")

    def download_probes(self):
        """D2H of all probe traces (one copy) -> list of arrays [n_probes, T] per member."""
        self.host_probe.copy_(self.pool[self.n_state + self.n_in:], non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        hv = self.host_probe.numpy()
        shapes = {(max(len(m.probe_idx), 1), len(m.probe_idx), (m.T + 31) // 32 * 32, m.T) for m in self.members}
        if len(shapes) == 1:        # uniform batch (the sweep case): one strided copy instead of one per member
            rows, n_p, Tp, T = next(iter(shapes))
            return list(hv.reshape(len(self.members), rows, Tp)[:, :n_p, :T].copy())
        out = []
        o0 = self.n_state + self.n_in
        for i, m in enumerate(self.members):
            Tp = (m.T + 31) // 32 * 32
            n_p = len(m.probe_idx)
            a = hv[self.off_probe[i] - o0: self.off_probe[i] - o0 + max(n_p, 1) * Tp]
            out.append(a.reshape(max(n_p, 1), Tp)[:n_p, : m.T].copy())
        return out

    def download_probes_async(self):
        """Stream-ordered D2H of all probe traces into this batch's pinned buffer; returns a function that
        waits for the copy and unpacks it like download_probes() (call it after other work was queued)."""
        torch = self.torch
        self.host_probe.copy_(self.pool[self.n_state + self.n_in:], non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())

        def finish():
            done.synchronize()
            hv = self.host_probe.numpy()
            shapes = {(max(len(m.probe_idx), 1), len(m.probe_idx), (m.T + 31) // 32 * 32, m.T) for m in self.members}
            if len(shapes) == 1:
                rows, n_p, Tp, T = next(iter(shapes))
                return list(hv.reshape(len(self.members), rows, Tp)[:, :n_p, :T].copy())
            o0 = self.n_state + self.n_in
            out = []
            for i, m in enumerate(self.members):
                Tp = (m.T + 31) // 32 * 32
                n_p = len(m.probe_idx)
                a = hv[self.off_probe[i] - o0: self.off_probe[i] - o0 + max(n_p, 1) * Tp]
                out.append(a.reshape(max(n_p, 1), Tp)[:n_p, : m.T].copy())
            return out
        return finish

    def download_probe(self, i):
        """D2H of the probe traces of member i only -> array [n_probes, T]."""
        m = self.members[i]
        Tp = (m.T + 31) // 32 * 32
        n_p = len(m.probe_idx)
        a = self.pool[self.off_probe[i]: self.off_probe[i] + max(n_p, 1) * Tp].cpu().numpy()
        return a.reshape(max(n_p, 1), Tp)[:n_p, : m.T].copy()

    def probe_peaks(self, probe=0, keep_from=None, keep_to=None):
        """RefTester's scalar of probe ``probe`` of every member, computed on the device (batched FFT per
        distinct timeSteps; nothing but the result, one double per member, is copied to the host).
        keep_from / keep_to: per-member first / last sample index of the read window (others are zeroed)."""
        from . import TransformHandler as transH
        torch = self.torch
        M = len(self.members)
        kf = None if keep_from is None else np.asarray(keep_from, dtype=np.int64)
        kt = None if keep_to is None else np.asarray(keep_to, dtype=np.int64)
        groups = {}
        for i, m in enumerate(self.members):
            groups.setdefault(m.T, []).append(i)
        out = torch.empty(M, dtype=torch.float64, device=self.device)
        for T, idxs in groups.items():
            Tp = (T + 31) // 32 * 32
            rows = torch.stack([self.pool[self.off_probe[i] + probe * Tp: self.off_probe[i] + probe * Tp + T] for i in idxs])
            val, _ = transH.ref_tester_batch(rows, T, None if kf is None else kf[idxs], None if kt is None else kt[idxs])
            out[torch.as_tensor(idxs, device=self.device)] = val
        return out.cpu().numpy()

    def state(self, i, name):
        """Device -> host copy of one state array of member i (tests / final fields)."""
        m = self.members[i]
        Lp = (m.L + 31) // 32 * 32
        o = self.off_state[i] + STATE_NAMES.index(name) * Lp
        return self.pool[o:o + m.L].cpu().numpy()


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


def frequency_sweep(V, P, domainSize, lowLimTim, highLimTim, Low=3e9, Interval=1e8, points=20, batched=True,
                    device_postproc=True):
    """MasterController.LoopedSim(loop=True) :533-569.  Returns (freqs, measured R, analytical R,
    (V,P,C_V,C_P,Exs,Hys) of the last member).  With ``batched`` and ``device_postproc`` the reflection
    extraction (results(RefCo=True) -> RefTester) runs on the device too and only the last member's traces
    come back to the host."""
    from . import MasterController as MC
    freqs = np.arange(Low, points * Interval + Low, Interval)[:points]
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
