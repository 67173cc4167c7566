#!/usr/bin/env python
"""bench.py -- throughput of the batched dispersive (Lorentz ADE + CPML) 1-D FDTD hot path.

Workload (BASELINE.json configs[1], batched as its `metric` says): a frequency/amplitude sweep of
M independent Lorentz-slab runs at the reference's default geometry family (9 GHz-class, 0.7 m
domain, Nlam = 400 -> ~10-15 k cells per member, CPML both sides, TF/SF sine source).  One bench
"step" = one sweep pass of S time steps for every member with the polarisation update on (pass 1 of
IntegratorLinLor1D) and the reflection probe recorded every step.  Every step restarts from the same
synthetic NON-ZERO state (uniform random values of each field's natural magnitude in every cell) so
that the measured rate is the steady-state rate of a run whose wave has filled the grid, not the rate
of a mostly-quiescent grid.

  value : Gcell-updates/s with inputs (CPML profiles, source tables) already resident in HBM
  e2e   : the same step through the host API (sweep.MemberBatch): pinned-host -> device copy of the
          inputs, on-device state clear, S steps, device -> host copy of the probe traces
  roofline : the tile kernel's algorithmic HBM bytes (k = 1 figure of SURVEY 8d) / its CUDA-event
          time, against the measured HBM peak.  The kernel is temporally blocked (k steps per HBM
          round trip), so a fraction above 1 is the design goal, not an error: `temporal_block_k`
          and `hbm_bytes_per_launch_model` say how many bytes a launch really moves.
  cpu_baseline : the oracle's C port of the reference loop on the host cores (bounded sample)

--impl reference times that CPU port with every host thread, same workload family, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "Gcell-updates/s batched dispersive 1D FDTD"
UNIT = "Gcell-updates/s"


# ------------------------------------------------------------------------------------------------ workload
def member_specs(n_members, n_freq=64):
    """(frequency, amplitude) of every member: n_freq frequencies x amplitudes."""
    freqs = np.linspace(6e9, 10.5e9, n_freq)
    n_amp = max(1, (n_members + n_freq - 1) // n_freq)
    amps = np.linspace(0.1, 10.0, n_amp) if n_amp > 1 else np.array([1.0])
    out = []
    for a in amps:
        for f in freqs:
            out.append((float(f), float(a)))
    return out[:n_members]


def alg_bytes_per_cell_step(L, pw, mf, mr):
    """SURVEY 8(d) state-only bytes per cell-update, summed over one member's cells (k = 1):
    32 B per cell (Ex, Hy r+w) + 64 B per CPML cell (psi_E, psi_H r+w, 4 profile reads)
    + 40 B per Lorentz slab cell (Dx r+w, P r+w, Pprev r)."""
    cpml = max(0, pw - 1) + pw
    slab = max(0, mr - mf)
    return 32 * L + 64 * cpml + 40 * slab


def dp_instr_per_cell_step(L, pw, mf, mr):
    """Separately rounded fp64 instructions of the reference's own arithmetic per time step, summed over one
    member's cells (exact mode, pass 1 of the Lorentz integrator): vacuum cell 6 (Ex 3 + Hy 3); CPML cell +10
    (psi_E 3 + correction 2 + psi_H 3 + correction 2); Lorentz slab cell 15 (P 5, dH 1, Dx 2, Dx-P 1, /eps0 3, Hy 3;
    its Ex += ... is dead because ADE_ExCreate overwrites it) and +6 where the slab lies inside the CPML."""
    cpml_left = max(0, pw - 1)
    slab = max(0, mr - mf)
    slab_in_cpml = max(0, mr - max(mf, L - pw))
    vac = L - slab
    return 6 * vac + 10 * (cpml_left + 2) + 15 * slab + 6 * slab_in_cpml


class ProductWorkload:
    def __init__(self, n_members, steps_per_pass, n_freq):
        import pyfdtd_b200  # noqa: F401
        from pyfdtd_b200 import MasterController as MC, Environment_Setup as envDef, Solver_Engine as SE, sweep
        self.sweep = sweep
        specs = member_specs(n_members, n_freq)
        first_of_freq = {}
        members, share = [], []
        for i, (f, amp) in enumerate(specs):
            if f in first_of_freq:
                j = first_of_freq[f]
                base = members[j]
                V, P, C_V, C_P = base.V, base.P, base.C_V, base.C_P
                m = sweep.Member(V, P, C_V, C_P, base._Exs * amp, base._Hys * amp, [P.x2Loc], nsteps=steps_per_pass)
                share.append(j)
            else:
                tup = envDef.envSetup(f, 0.7, 7000, 8000, LorMed=True)
                P = MC.Params(*tup, False, 0.7, f, 20)
                P.TFSF, P.SineCont, P.Periods, P.LorentzMed, P.FreeSpace = True, True, 1000, True, False
                V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
                C_P = MC.CPML_Params(P.dz)
                C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
                for _ in range(2):   # pass 1 state of the setup chain (twice-corrected plasma frequency)
                    C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=True)
                m = sweep.Member(V, P, C_V, C_P, Exs * amp, Hys * amp, [P.x2Loc], nsteps=steps_per_pass)
                m._Exs, m._Hys = Exs, Hys
                first_of_freq[f] = i
                share.append(i)
            # only the first steps_per_pass entries of the source tables are used
            m.T = steps_per_pass
            m.srcE, m.srcH = m.srcE[:steps_per_pass], m.srcH[:steps_per_pass]
            members.append(m)
        self.members, self.share = members, share
        self.batch = sweep.MemberBatch(members, "lorentz", share_coef=share)
        self.cell_steps = self.batch.cell_steps
        self.dp_instr_per_step = sum(dp_instr_per_cell_step(m.L, m.scalars["pw"], m.scalars["mf"], m.scalars["mr"])
                                     for m in members)
        self.alg_bytes_per_step = sum(alg_bytes_per_cell_step(m.L, m.scalars["pw"], m.scalars["mf"], m.scalars["mr"])
                                      for m in members)          # per time step, all members


# ------------------------------------------------------------------------------------------------ other BASELINE configs
def _time_cuda(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def build_nl_batch(M, S):
    """Config 3 workload: M members of the nonlinear (cubic) sweep, S time steps each (16 distinct grids)."""
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import MasterController as MC, Environment_Setup as envDef, Solver_Engine as SE, sweep
    freqs = np.linspace(6e9, 10.5e9, 16)
    members, share, first = [], [], {}
    for i in range(M):
        f = float(freqs[i % 16])
        if f in first:
            b = members[first[f]]
            m = sweep.Member(b.V, b.P, b.C_V, b.C_P, b._Exs * (1 + i / M), b._Hys * (1 + i / M), [b.P.materialFrontEdge], nsteps=S)
            share.append(first[f])
        else:
            tup = envDef.envSetup(f, 0.7, 7000, 8000, nonLinMed=True)
            P = MC.Params(*tup, False, 0.7, f, 20)
            P.TFSF, P.SineCont, P.Periods, P.nonLinMed, P.FreeSpace, P.LorentzMed = True, True, 1000, True, False, False
            V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
            C_P = MC.CPML_Params(P.dz)
            C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
            C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False, nonlinear=True)
            m = sweep.Member(V, P, C_V, C_P, Exs, Hys, [P.materialFrontEdge], nsteps=S)
            m._Exs, m._Hys = Exs, Hys
            first[f] = i
            share.append(i)
        m.T = S
        m.srcE, m.srcH = m.srcE[:S], m.srcH[:S]
        members.append(m)
    batch = sweep.MemberBatch(members, "nl", share_coef=share)
    batch.upload()
    batch.randomize_state()
    return batch, members


def extras(torch, peak_gbs, quick=False):
    """Short measurements of the other BASELINE.json configs (device-resident, synthetic non-zero state)."""
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import MasterController as MC, Environment_Setup as envDef, Solver_Engine as SE, longgrid, pic, sweep
    out = {}
    # --- config 1 / 5: one long grid, streaming with k-step temporal blocking
    big = (1 << 24) if quick else 100_000_000
    for name, mode, cells, alg, kw in (("free_long_grid", "free", 1 << 24, 32.0, {}),
                                       ("lorentz_long_grid", "lorentz", big, 32.0 + 0.7 * 40.0, {}),
                                       # config 5 as BASELINE words it: dispersive AND nonlinear (PF_LORENTZ_NL, the
                                       # Lorentz ADE + the cubic Kerr law on Dx - P, converged Newton root)
                                       ("kerr_lorentz_long_grid", "lorentz_nl", big, 32.0 + 0.7 * 40.0, {}),
                                       ("lorentz_long_grid_fp32", "lorentz", big, 32.0 + 0.7 * 40.0, {"fp32": True})):
        steps = 128
        grid, info = longgrid.lorentz_long_grid(cells, T=steps + 64, k=64, mode=mode, **kw)
        for which in (0, 1):
            for arrs in grid.bufs[which]:
                for n, t in arrs.items():
                    if t is not None:
                        t.copy_((torch.rand_like(t) * 2 - 1) * sweep.MemberBatch.STATE_SCALE[n])
        for arrs0, arrs1 in zip(*grid.bufs):
            for n in arrs0:
                if arrs0[n] is not None:
                    arrs1[n].copy_(arrs0[n])
        sec = _time_cuda(torch, lambda: grid.run(steps, do_pol=(mode != "free")), 2)
        rate = cells * steps / sec / 1e9
        out[name] = {"cells": cells, "steps": steps, "pieces": len(grid.pieces), "Gcell_updates_per_s": rate,
                     "algorithmic_GBps_k1": rate * alg, "frac_of_hbm_peak_k1": rate * alg / peak_gbs, "temporal_block_k": 64}
        del grid
        torch.cuda.empty_cache()
    # --- configs 1/2 as the reference runs them: ONE default-geometry run through Controller (host API, e2e)
    import time as _t
    for name, lor in (("single_run_free_default", False), ("single_run_lorentz_default", True)):
        tup = envDef.envSetup(9e9, 0.7, 7000, 8000, LorMed=lor)
        P = MC.Params(*tup, False, 0.7, 9e9, 20)
        P.TFSF, P.SineCont, P.Periods, P.LorentzMed, P.FreeSpace = True, True, 1000, lor, not lor
        V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 10)
        C_P = MC.CPML_Params(P.dz)
        C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
        best = None
        for _ in range(2):
            t0 = _t.perf_counter()
            MC.Controller(V, P, C_V, C_P)
            torch.cuda.synchronize()
            dtc = _t.perf_counter() - t0
            best = dtc if best is None else min(best, dtc)
        cu = 2 * P.timeSteps * (P.Nz + 1)
        out[name] = {"Nz": P.Nz, "timeSteps": P.timeSteps, "passes": 2, "seconds_e2e": best, "Mcell_updates_per_s": cu / best / 1e6,
                     "reference_seconds_build_container": 130.2 if lor else 7.3,
                     "note": "Controller() incl. host setup, H2D/D2H, Ex_History snapshots every 50 steps; reference time = "
                             "unmodified reference on the build container CPU (tests/golden/*_default_full.npz: ref_wall_seconds)"}
    # --- config 3: nonlinear (cubic solve per slab cell per step) sweep batch
    M, S = (64, 64) if quick else (256, 128)
    batch, members = build_nl_batch(M, S)

    def nl_step():
        batch.reset_state(template=True)
        batch.run(do_pol=False)
    sec = _time_cuda(torch, nl_step, 2)
    slab_cells = sum(m.scalars["mr"] - m.scalars["mf"] for m in members)
    out["nl_cubic_sweep"] = {"members": M, "steps": S, "Gcell_updates_per_s": batch.cell_steps / sec / 1e9,
                             "cubic_solves_per_s": slab_cells * S / sec, "cubic": "closed form (reference algorithm)"}
    del batch
    torch.cuda.empty_cache()
    # the same sweep with the optional arithmetic modes (not parity modes; tolerances in tests/test_gpu_parity.py)
    for label, fp32, cubic in (("nl_cubic_sweep_newton", False, "newton"), ("nl_cubic_sweep_fp32", True, "closed")):
        SE.USE_FP32, SE.CUBIC = fp32, cubic
        try:
            batch, members = build_nl_batch(M, S)
            sec = _time_cuda(torch, nl_step, 2)
            out[label] = {"members": M, "steps": S, "Gcell_updates_per_s": batch.cell_steps / sec / 1e9,
                          "cubic_solves_per_s": slab_cells * S / sec}
        finally:
            SE.USE_FP32, SE.CUBIC = False, "closed"
        del batch
        torch.cuda.empty_cache()
    # --- the headline workload (config 2 sweep) in the optional arithmetic modes
    if not quick:
        modes = {}
        for label, fma, fp32 in (("fma_contracted", True, False), ("fp32", False, True)):
            SE.USE_FMA, SE.USE_FP32 = fma, fp32
            try:
                wl2 = ProductWorkload(1024, 512, 64)
                b2 = wl2.batch
                b2.upload()
                b2.randomize_state(seed=1234)

                def sweep_step():
                    b2.reset_state(template=True)
                    b2.run(do_pol=True)
                sec = _time_cuda(torch, sweep_step, 3)
                modes[label] = wl2.cell_steps / sec / 1e9
            finally:
                SE.USE_FMA, SE.USE_FP32 = False, False
            del b2, wl2
            torch.cuda.empty_cache()
        out["lorentz_sweep_optional_modes_Gcell_updates_per_s"] = dict(
            modes, note="same 1024-member workload as `value`; fma: PF_F_FMA (<= 1e-10 relative, tested); fp32: PF_F_FP32 "
                        "(stated tolerance 1e-5 of the trace peak, profiles/r1l_fp32_accuracy.json)")
    # --- config 4: PIC push + cell sort + deterministic deposit
    L, dz, dt = 13194, 8.3276e-5, 2.6389e-13
    n = 2_000_000 if quick else 20_000_000
    z, ux, uz, w = pic.make_beam(n, L, dz, seed=1)
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    Ex = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 1e5
    Hy = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 3e2

    def pic_step_radix():
        ps.push(Ex, Hy)
        ps.deposit()       # sorts first (radix sort: particles left their cells), then deposits
    sec_radix = _time_cuda(torch, pic_step_radix, 3)

    def pic_step():
        ps.push_sorted(Ex, Hy)   # fused push + counting re-sort (particles move < 1 cell per step)
        ps.deposit()
    sec = _time_cuda(torch, pic_step, 5)
    sec_push = _time_cuda(torch, lambda: ps.push(Ex, Hy), 5)
    ps.sort()
    sec_fused = _time_cuda(torch, lambda: ps.step_sorted(Ex, Hy), 5)   # deposit fused into the move pass
    # the same particle step COUPLED to the field grid: deposit -> one FDTD step subtracting Jx (ADE_ExUpdate's slot,
    # BaseFDTD11.py:667; per-op engine, the one that carries per-cell arrays) -> push in the new fields
    from pyfdtd_b200 import BaseFDTD11, _device as dev
    tup = envDef.envSetup(9e9, 0.7, 7000, 8000)
    P = MC.Params(*tup, False, 0.7, 9e9, 20)
    P.TFSF, P.SineCont, P.Periods, P.LorentzMed, P.FreeSpace = True, True, 1000, False, True
    V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 1)
    C_P = MC.CPML_Params(P.dz)
    C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
    C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=False)
    Lc = len(V.Ex)
    zc, uxc, uzc, wc = pic.make_beam(n, Lc, P.dz, seed=2)
    psc = pic.ParticleSet(zc, uxc, uzc, wc, Lc, P.dz, P.delT)
    gdev = dev.DeviceGrid(L=Lc, T=P.timeSteps, arrays=BaseFDTD11._host_arrays(V, C_V, V.tempVarPol),
                          scalars=BaseFDTD11.grid_scalars(V, P), srcE=np.asarray(Exs) / P.courantNo,
                          srcH=np.asarray(Hys) / P.courantNo, probe_idx=[], flags=BaseFDTD11.grid_flags(P))
    sim = pic.CoupledPIC(gdev, psc, mode="free", fused=True)
    for _ in range(3):
        sim.step()
    sec_coupled = _time_cuda(torch, sim.step, 10)
    out["pic"] = {"particles": n, "particle_steps_per_s": n / sec_fused,
                  "coupled_to_fdtd_particle_steps_per_s": n / sec_coupled, "coupled_grid_cells": Lc, "push_only_particles_per_s": n / sec_push,
                  "algorithmic_GBps": 60.0 * n / sec_fused / 1e9, "frac_of_hbm_peak": 60.0 * n / sec_fused / 1e9 / peak_gbs,
                  "separate_deposit_pass_particle_steps_per_s": n / sec,
                  "radix_sort_variant_particle_steps_per_s": n / sec_radix,
                  "note": "step = pf_pic_step_sorted: Boris push + stable counting re-sort by cell (count/scan/move) with "
                          "the deterministic deposit fused into the move pass; 60 B/particle-step algorithmic; coupled = the same "
                          "step plus one FDTD step of the reference's default grid with the deposited Jx subtracted"}
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(n_members, steps, threads, warmup=1, reps=1):
    """Time the oracle's C restatement of IntegratorLinLor1D's pass-1 loop on `threads` host threads
    over `n_members` members of the bench workload family (bounded sample)."""
    import ctypes
    import fdtd_oracle as fo
    specs = member_specs(n_members, n_freq=min(64, n_members))
    passes = []
    for f, amp in specs:
        c = fo.make_case("lorentz", f, 0.7, 7000, 8000, source="sine", periods=1000)
        wp = c.medium["wp"]
        for _ in range(2):
            wp, _ = fo.spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp, c.medium["w0"], c.medium["gam"])
        Exs, Hys = fo.sources(c)
        passes.append((c, fo.PassArrays(c, wp, Exs * amp, Hys * amp, [c.x2Loc], False)))
    Grids = fo.OrcGrid * len(passes)
    T_tot = (ctypes.c_int * len(passes))(*[c.T for c, _ in passes])
    cells = sum(c.L for c, _ in passes) * steps
    lib = fo.lib()
    times = []
    for it in range(warmup + reps):
        rng = np.random.default_rng(1234)
        for _, pa in passes:     # same synthetic non-zero state as the GPU arm (sweep.MemberBatch.STATE_SCALE)
            for a, sc in ((pa.Ex, 1.0), (pa.Hy, 1.0 / 376.73), (pa.Dx, 8.85e-12), (pa.P, 8.85e-12),
                          (pa.Pprev, 8.85e-12), (pa.psiE, 1e-7), (pa.psiH, 1.0)):
                a[:] = rng.uniform(-1.0, 1.0, len(a)) * sc
        grids = Grids(*[pa.g for _, pa in passes])
        t0 = time.perf_counter()
        lib.orc_run_batch(grids, len(passes), 1, 1, 0, steps, T_tot, threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return cells / np.mean(times) / 1e9, float(np.mean(times)), cells


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_members = max(threads, min(args.members, 4 * threads))
    steps = args.cpu_steps
    rate, sec, cells = cpu_port_rate(n_members, steps, threads, warmup=args.warmup, reps=args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batched Lorentz-ADE+CPML 1D FDTD sweep (reference IntegratorLinLor1D pass-1 loop), "
                               f"CPU sample: {n_members} members x {steps} steps, Nz~10-15k cells/member"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_members} members x {steps} steps ({cells/1e9:.2f} Gcell-updates/step), C port of the "
                                   "reference loop (oracle/fdtd_oracle.c), one member per thread"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=1024, help="sweep members per GPU")
    ap.add_argument("--pass-steps", type=int, default=512, help="time steps per bench step (S)")
    ap.add_argument("--k-block", type=int, default=0, help="time steps per launch (0 = library default)")
    ap.add_argument("--n-freq", type=int, default=64)
    ap.add_argument("--cpu-steps", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fma", action="store_true", help="PF_F_FMA kernels (not bit-identical; reported in config)")
    ap.add_argument("--fp32", action="store_true", help="PF_F_FP32 kernels (optional single-precision mode; reported as dtype f32)")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--quick-extras", action="store_true", help="smaller extras (smoke)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import Solver_Engine as SE, _native as nat
    SE.USE_FMA = bool(args.fma)
    SE.USE_FP32 = bool(args.fp32)
    lib = nat.lib()
    wl = ProductWorkload(args.members, args.pass_steps, args.n_freq)
    batch = wl.batch
    kcfg = [nat.c_int(), nat.c_int(), nat.c_int()]
    lib.pf_tile_config(*kcfg)
    k_block = args.k_block or 64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch.randomize_state(seed=1234 + rank)

    def step_resident():
        batch.reset_state(template=True)
        batch.run(do_pol=True, k_block=args.k_block)

    def step_e2e():
        batch.upload()
        batch.reset_state(template=True)
        batch.run(do_pol=True, k_block=args.k_block)
        return batch.download_probes()

    batch.upload()
    for _ in range(args.warmup):
        step_resident()
    # ---- timed region: K steps, device-timed, max over ranks -------------------------------
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.pf_launch_count()
    lib.pf_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    lib.pf_profile_enable(0)
    clocks = sampler.stop()
    launches = lib.pf_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    kms, kn = nat.c_double(), nat.c_int()
    lib.pf_profile_collect(kms, kn)

    # ---- e2e: host buffers, H2D + D2H inside the timed region -------------------------------
    # (a) one batch at a time: upload -> run -> download, strictly serial
    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        traces = step_e2e()
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t0
    # (b) the public pipelined runner (sweep.BatchPipeline): consecutive steps alternate between two batch pools, so the
    # H2D of step i+1's inputs and the D2H + host unpacking of step i-1's traces overlap the time stepping of step i;
    # every step still uploads its own inputs from pinned memory and returns its own traces to the host
    batch_b = wl.sweep.MemberBatch(wl.members, "lorentz", share_coef=wl.share)
    batch_b.upload()
    batch_b.randomize_state(seed=4321 + rank)
    pipe = wl.sweep.BatchPipeline([batch, batch_b])
    pipe.run([0, 1], True, template=True, k_block=args.k_block)
    barrier()
    t0 = time.perf_counter()
    results = pipe.run([i % 2 for i in range(args.steps)], True, template=True, k_block=args.k_block)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    probe_checksum = float(sum(np.abs(t).sum() for t in results[0]))
    assert len(results) == args.steps
    if abs(probe_checksum - float(sum(np.abs(t).sum() for t in traces))) > 1e-9 * max(1.0, abs(probe_checksum)):
        raise SystemExit("bench: pipelined and serial e2e steps disagree")
    del batch_b, pipe

    t_dev = torch.tensor([ms_total, e2e_s * 1e3, e2e_serial_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_serial_ms = [float(x) for x in t_dev.cpu()]
    total_cell_steps = wl.cell_steps * world * args.steps
    value = total_cell_steps / (ms_total * 1e-3) / 1e9
    e2e_value = total_cell_steps / (e2e_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_launch_per_step = max(1, kn.value // max(1, args.steps))
    avg_kernel_ms = kms.value / max(1, kn.value)
    alg_bytes_per_launch = wl.alg_bytes_per_step * args.pass_steps / n_launch_per_step
    achieved = alg_bytes_per_launch / (avg_kernel_ms * 1e-3) / 1e9
    tile_cells, halo = kcfg[0].value, k_block
    # modelled real HBM traffic of one launch: every tile reads tile_cells of state, writes its interior
    hbm_model = wl.alg_bytes_per_step * (1.0 + 2.0 * halo / (tile_cells - 2 * halo))
    fp64_peak, fp64_src = 1.85e13, "fallback (round-1 probe on this pool)"
    probe_path = os.path.join(ROOT, "profiles", "r1_fp64_probe.json")
    if os.path.exists(probe_path):
        fp64_peak = 2.0 * json.load(open(probe_path))["dmul_dadd_pairs_per_s_t1024"]
        fp64_src = "measured (tools/fp64_probe.cu, profiles/r1_fp64_probe.json: separately rounded DMUL+DADD stream)"
    dp_per_launch = wl.dp_instr_per_step * args.pass_steps / n_launch_per_step
    fp64 = {"achieved_dp_instr_per_s": dp_per_launch / (avg_kernel_ms * 1e-3), "peak_dp_instr_per_s": fp64_peak,
            "frac": dp_per_launch / (avg_kernel_ms * 1e-3) / fp64_peak, "peak_source": fp64_src,
            "note": "algorithmic fp64 instructions of the reference arithmetic (exact mode, no FMA contraction) / kernel time; "
                    "this, not HBM, is the pipe that bounds the temporally blocked kernel"}
    # DRAM bytes of one launch of this kernel from the committed ncu --set full capture of the DEFAULT workload
    # (profiles/r1_final_k_tile_ncu.txt: dram__bytes_read.sum 539.5 MB + dram__bytes_write.sum 469.8 MB)
    traffic = 537.556736e6 + 462.760448e6 if (args.members == 1024 and k_block == 64 and args.n_freq == 64) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if (args.fp32 or args.fma) else traffic, "kernel": "k_tile<PF_LORENTZ, POL=1, C=2, %s>" % ("Fast32" if args.fp32 else "Fused" if args.fma else "Exact"), "peak_source": peak_src,
                "kernel_ms_avg": avg_kernel_ms, "kernel_launches_timed": kn.value,
                "kernel_share_of_step": kms.value / ms_total if ms_total else None,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "temporal_block_k": k_block,
                "hbm_bytes_per_launch_model": hbm_model, "fp64_pipe": None if args.fp32 else fp64,
                "note": "on-chip temporally blocked: algorithmic (k=1) bytes / time exceeds the HBM roofline by design"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_cpu = max(threads, min(args.members, 4 * threads))
        rate, sec, cells = cpu_port_rate(n_cpu, args.cpu_steps, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} members x {args.cpu_steps} steps of the same sweep family on {threads} threads "
                         f"({sec:.2f} s wall)"}

    cfg_cells, cfg_state_mb, h2d_b, d2h_b = wl.cell_steps, batch.n_state * 8 / 1e6, batch.h2d_bytes, batch.d2h_bytes
    extra = None
    if rank == 0 and world == 1 and not args.no_extras:
        del wl, batch
        torch.cuda.empty_cache()
        try:
            extra = extras(torch, peak, quick=args.quick_extras)
        except Exception as e:   # the headline line must survive a failing extra
            extra = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": f"batched Lorentz-ADE+CPML 1D FDTD sweep, {args.members} members/GPU x "
                                   f"{args.pass_steps} time steps per step, Nz~10-15k cells/member "
                                   f"({cfg_cells/1e9:.2f} Gcell-updates/step/GPU), probes recorded every step",
                       "members_per_gpu": args.members, "time_steps_per_step": args.pass_steps,
                       "k_block": k_block, "arithmetic": ("fp32 on chip (PF_F_FP32, stated tolerance 1e-5; not a parity mode)" if args.fp32 else
                                      "fma-contracted" if args.fma else "exact (bit-identical to reference order)"),
                       "l2_policy": f"state {cfg_state_mb:.0f} MB per GPU > 126 MB L2, restored from a random template every step",
                       "parallelism": f"members sharded over {world} GPU(s), no collectives"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                    "ms_per_step": e2e_ms / args.steps, "probe_checksum": probe_checksum,
                    "how": "sweep.BatchPipeline: steps alternate between two batch pools; H2D of the next step's inputs and D2H of "
                           "the previous step's traces overlap the time stepping (3 streams)",
                    "serial_value": total_cell_steps / (e2e_serial_ms * 1e-3) / 1e9,
                    "serial_how": "MemberBatch.upload -> run -> download_probes, one step at a time"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["other_configs"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
