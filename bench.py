#!/usr/bin/env python
"""bench.py -- throughput of the batched dispersive (Lorentz ADE + CPML) 1-D FDTD hot path, plus one measured leg per
BASELINE.json config.

Headline workload (BASELINE.json configs[1], batched as its `metric` says): a frequency/amplitude sweep of M independent
Lorentz-slab runs at the reference's default geometry family (6-10.5 GHz, 0.7 m domain, Nlam = 400 -> ~10-15 k cells per
member, CPML both sides, TF/SF sine source).  One bench "step" = one sweep pass of S time steps for every member with the
polarisation update on (pass 1 of IntegratorLinLor1D) and the reflection probe recorded every step.  Every step restarts
from the same synthetic NON-ZERO state (uniform random values of each field's natural magnitude in every cell) so that the
measured rate is the steady-state rate of a run whose wave has filled the grid, not that of a mostly-quiescent grid.

  value    : Gcell-updates/s with inputs (CPML profiles, source tables) already resident in HBM
  e2e      : the same step through the host API (sweep.BatchPipeline over sweep.MemberBatch): pinned-host -> device copy of
             the inputs, on-device state restore, S steps, device -> host copy of the probe traces
  e2e_full_sweep : a whole reflection-vs-frequency sweep of 1024 DISTINCT grids through the product call
             sweep.reflection_sweep -- vectorised host setup, native input building, H2D, 2 passes x full timeSteps,
             reflection extraction on the device -- wall clock from the call to the result on the host
  roofline : the dominant kernel (k_tile) against the pipe that bounds it.  The kernel advances k = 64 steps per HBM round
             trip, so it is bound by the FP64 pipe, not HBM: achieved = the reference arithmetic's own fp64 instructions
             (separately rounded, no contraction) per second over the kernel's CUDA-event time, peak = the DMUL+DADD
             instruction rate of this GPU measured in the same run (pf_probe_fp64).  The HBM view (algorithmic k = 1 bytes,
             modelled real traffic, ncu-measured DRAM bytes when a capture of this workload is committed) is kept beside it.
  configs  : one entry per other BASELINE config, each with kernel name, kernel ms and both roofline fractions
  long_grid: config 5 (the only communicating config) at every N: weak (1e8 cells/GPU) and strong (one fixed 1e9-cell grid)
  cpu_baseline : the oracle's C port of the reference loop on the host cores, same member family

--impl reference times that CPU port with every host thread on the SAME workload (members x steps of the headline).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "Gcell-updates/s batched dispersive 1D FDTD"
UNIT = "Gcell-updates/s"
DOM, WIN = 0.7, (7000, 8000)


# ------------------------------------------------------------------------------------------------ workload
def member_specs(n_members, n_freq=64):
    """(frequency, amplitude) of every member: n_freq frequencies x amplitudes, amplitude-major."""
    freqs = np.linspace(6e9, 10.5e9, n_freq)
    n_amp = max(1, (n_members + n_freq - 1) // n_freq)
    amps = np.linspace(0.1, 10.0, n_amp) if n_amp > 1 else np.array([1.0])
    f = np.tile(freqs, n_amp)[:n_members]
    a = np.repeat(amps, n_freq)[:n_members]
    return f, a


def alg_bytes_per_step(L, pw, mf, mr, mode="lorentz"):
    """SURVEY 8(d) state-only bytes per time step, summed over members (k = 1): 32 B per cell (Ex, Hy r+w) + 64 B per CPML
    cell (psi_E, psi_H r+w, 4 profile reads) + per slab cell 40 B (Lorentz: Dx r+w, P r+w, Pprev r) or 16 B (cubic: Dx r+w)."""
    L, pw, mf, mr = (np.asarray(x, dtype=np.float64) for x in (L, pw, mf, mr))
    cpml = np.maximum(0, pw - 1) + pw
    slab = np.maximum(0, mr - mf)
    per_slab = {"lorentz": 40.0, "lorentz_nl": 40.0, "nl": 16.0, "free": 0.0}[mode]
    return float(np.sum(32 * L + 64 * cpml + per_slab * slab))


def dp_instr_per_step(L, pw, mf, mr, mode="lorentz"):
    """Separately rounded fp64 instructions of the reference's OWN arithmetic per time step, summed over members (exact
    mode): vacuum cell 6 (Ex 3 + Hy 3); CPML cell +10 (psi_E 3 + correction 2 + psi_H 3 + correction 2); Lorentz slab cell 15
    (P 5, dH 1, Dx 2, Dx-P 1, /eps0 3, Hy 3; its Ex += ... is dead because ADE_ExCreate overwrites it) and +6 where the slab
    lies inside the CPML.  Cubic slab cell (mode "nl"): 50, the DADD / DMUL / DFMA instructions the closed-form law executes
    per slab cell-step (ncu, profiles/r2e_k_tile_nl_closed_ncu.txt: 2.78e8 fp64 warp-instructions per 256-member x 64-step
    launch, halo recomputation and the non-slab cells taken out; sqrt, two cube roots and four divisions expanded) -- an
    EXECUTED count: the reference's own expression has ~20 simple operations + 4 divisions + 1 sqrt + 2 pow."""
    L, pw, mf, mr = (np.asarray(x, dtype=np.float64) for x in (L, pw, mf, mr))
    cpml_left = np.maximum(0, pw - 1)
    slab = np.maximum(0, mr - mf)
    slab_in_cpml = np.maximum(0, mr - np.maximum(mf, L - pw))
    vac = L - slab
    if mode == "free":
        return float(np.sum(6 * L + 10 * (cpml_left + pw)))
    per_slab = 50.0 if mode == "nl" else 15.0
    return float(np.sum(6 * vac + 10 * (cpml_left + 2) + per_slab * slab + 6 * slab_in_cpml))


def lorentz_sweep_batch(n_members, steps, n_freq, *, distinct=False, fma=False, fp32=False):
    """The headline batch through the vectorised setup path (sweep_setup tables + native input builder)."""
    from pyfdtd_b200 import sweep, sweep_setup
    if distinct:
        f, a = np.linspace(6e9, 10.5e9, n_members), np.ones(n_members)
    else:
        f, a = member_specs(n_members, n_freq)
    t = sweep_setup.lorentz_sweep_tables(f, a, DOM, *WIN, periods=1000, nsteps=steps, fma=fma, fp32=fp32)[1]
    batch = sweep.MemberBatch.from_table(t, "lorentz", T_alloc=steps)
    return batch, t


def nl_sweep_batch(n_members, steps, *, fp32=False, newton=False, rank=0, world=1):
    """Config 3: 64 frequencies x 64 amplitudes of the cubic integrator, members dealt round-robin over ranks."""
    from pyfdtd_b200 import sweep, sweep_setup
    f = np.repeat(np.linspace(6e9, 10.5e9, 64), 64)
    a = np.tile(np.linspace(0.1, 10.0, 64), 64)
    t = sweep_setup.nonlinear_sweep_table(f, a, DOM, *WIN, nsteps=steps, fp32=fp32, newton=newton)
    t = t.select(np.arange(rank, len(f), world)[:n_members])
    t.probes = t.probes[:, :1].copy()
    batch = sweep.MemberBatch.from_table(t, "nl", T_alloc=steps)
    return batch, t


# ------------------------------------------------------------------------------------------------ measurement helpers
def time_cuda(torch, fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


_time_cuda = time_cuda   # name used by tools/*_profile.py


def timed_with_kernels(torch, nat, fn, reps, warm=1):
    """(seconds per call, {kernel: (launches, ms)}) -- CUDA events around the calls and around every profiled launch."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    lib = nat.lib()
    lib.pf_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    lib.pf_profile_enable(0)
    return e0.elapsed_time(e1) * 1e-3 / reps, nat.profile_report()


def roofline_entry(kernels, reps, dp_instr, alg_bytes, k, tile_cells, fp64_peak, hbm_peak, fp64_executed_note=None):
    """Roofline of the dominant kernel of one config.  dp_instr / alg_bytes: per call of the timed function."""
    if not kernels:
        return None
    name, (n, ms) = max(kernels.items(), key=lambda kv: kv[1][1])
    sec = ms * 1e-3 / reps
    real_bytes = alg_bytes / k * (1.0 + 2.0 * k / (tile_cells - 2 * k)) if k else alg_bytes
    out = {"kernel": name, "kernel_ms": ms / n, "kernel_launches_per_call": n / reps, "kernel_ms_per_call": ms / reps,
           "fp64": {"achieved_dp_instr_per_s": dp_instr / sec, "peak_dp_instr_per_s": fp64_peak, "frac": dp_instr / sec / fp64_peak},
           "hbm": {"algorithmic_GBps_k1": alg_bytes / sec / 1e9, "frac_k1": alg_bytes / sec / 1e9 / hbm_peak,
                   "modelled_real_GBps": real_bytes / sec / 1e9, "frac_real": real_bytes / sec / 1e9 / hbm_peak,
                   "temporal_block_k": k}}
    if fp64_executed_note:
        out["fp64"]["note"] = fp64_executed_note
    return out


def committed_ncu_traffic(workload_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of THIS workload
    (profiles/ncu_traffic.json: written by tools/ncu_summary.py from the .ncu-rep, keyed by workload); None if absent."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        ent = json.load(open(path)).get(workload_key)
        return (ent["dram_bytes_read"] + ent["dram_bytes_write"], ent["source"]) if ent else (None, None)
    except Exception:
        return None, None


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
def cpu_port_rate(n_members, steps, threads, n_freq=64, warmup=1, reps=1):
    """Time the oracle's C restatement of IntegratorLinLor1D's pass-1 loop on `threads` host threads over the first
    `n_members` members of the headline workload (same frequencies / amplitudes / synthetic state)."""
    import ctypes
    import fdtd_oracle as fo
    fr, am = member_specs(n_members, n_freq)
    passes, cache = [], {}
    for f, amp in zip(fr, am):
        f = float(f)
        if f not in cache:
            c = fo.make_case("lorentz", f, DOM, *WIN, source="sine", periods=1000)
            wp = c.medium["wp"]
            for _ in range(2):
                wp, _ = fo.spatial_stab(c.Nz, c.dz, c.freq, c.dt, wp, c.medium["w0"], c.medium["gam"])
            Exs, Hys = fo.sources(c)
            cache[f] = (c, wp, Exs, Hys)
        c, wp, Exs, Hys = cache[f]
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


def workload_text(members, steps, n_freq):
    return (f"batched Lorentz-ADE+CPML 1D FDTD sweep (IntegratorLinLor1D pass-1 loop), {members} members/GPU "
            f"({n_freq} frequencies 6-10.5 GHz x amplitudes) x {steps} time steps per step, Nz~10-15k cells/member, "
            "probes recorded every step, synthetic non-zero state")


def workload_config(args, arithmetic="exact (bit-identical to reference order)"):
    """The `config` object: the WORKLOAD only, identical for both arms (how an arm runs it goes under `run`)."""
    return {"workload": workload_text(args.members, args.pass_steps, args.n_freq), "members_per_gpu": args.members,
            "time_steps_per_step": args.pass_steps, "n_freq": args.n_freq, "arithmetic": arithmetic}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, sec, cells = cpu_port_rate(args.members, args.pass_steps, threads, args.n_freq, warmup=args.warmup, reps=args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "run": {"cell_updates_per_step": cells, "host_threads": threads, "what": "C port of the reference loop (oracle/fdtd_oracle.c), "
                "one member per thread, the same members, steps, sources and synthetic state as the GPU arm"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"the full bench step: {args.members} members x {args.pass_steps} steps "
                                   f"({cells/1e9:.2f} Gcell-updates), C port of the reference loop (oracle/fdtd_oracle.c), "
                                   f"one member per thread, {threads} threads; the reference itself is Python and is not on "
                                   "the GPU box (its own rate for this loop: 0.005 Gcell-updates/s on one core, SURVEY 6)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ per-config legs
def leg_long_grid(torch, nat, dist, rank, world, *, cells_total, steps, mode, fp64_peak, hbm_peak, label):
    """Config 5 / 1: one long grid of `cells_total` cells cut into contiguous z-ranges over the ranks (work-balanced),
    k = 64 ghost cells exchanged point-to-point every 64 steps.  Returns rank 0's dict (max over ranks for the times)."""
    from pyfdtd_b200 import longgrid, sweep
    k = 64
    grid, info = longgrid.lorentz_long_grid(cells_total, T=steps * 3 + k, k=k, mode=mode, rank=rank, world_size=world)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99 + rank)
    for arrs0, arrs1 in zip(*grid.bufs):
        for n in arrs0:
            if arrs0[n] is not None:
                arrs0[n].copy_((torch.rand(arrs0[n].shape, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1)
                               * sweep.MemberBatch.STATE_SCALE[n])
                arrs1[n].copy_(arrs0[n])
    lib = nat.lib()
    grid.run(steps, do_pol=(mode != "free"))          # warm-up (also builds the tile tables)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    grid.time_exchange = True
    lib.pf_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    grid.run(steps, do_pol=(mode != "free"))
    e1.record()
    torch.cuda.synchronize()
    lib.pf_profile_enable(0)
    kern = nat.profile_report()
    ms = e0.elapsed_time(e1)
    kms = sum(v[1] for v in kern.values())
    xms = grid.exchange_ms()
    t = torch.tensor([ms, kms, xms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kms_max, xms_max = [float(v) for v in t.cpu()]
    own = grid.cells_owned
    slab_own = sum(max(0, min(p["hi"], info["mr"]) - max(p["lo"], info["mf"])) for p in grid.mine) if mode != "free" else 0
    pml_own = sum(max(0, min(p["hi"], info["pw"]) - p["lo"]) + max(0, p["hi"] - max(p["lo"], cells_total - info["pw"])) for p in grid.mine)
    per_slab = {"lorentz": (15.0, 40.0), "lorentz_nl": (15.0, 40.0), "free": (6.0, 0.0)}[mode]
    dp = (6.0 * (own - slab_own) + per_slab[0] * slab_own + 10.0 * pml_own) * steps if mode != "free" else (6.0 * own + 10.0 * pml_own) * steps
    ab = (32.0 * own + per_slab[1] * slab_own + 64.0 * pml_own) * steps
    tc = [nat.c_int(), nat.c_int(), nat.c_int()]
    lib.pf_tile_config(*tc)
    out = {"label": label, "mode": mode, "cells": cells_total, "cells_per_gpu": cells_total // world, "steps": steps, "n_gpus": world,
           "pieces_this_rank": len(grid.mine), "Gcell_updates_per_s": cells_total * steps / (ms * 1e-3) / 1e9,
           "ms": ms, "kernel_ms_max_over_ranks": kms_max, "exchange_ms_max_over_ranks": xms_max,
           "exchange_share": xms_max / ms if ms else None,
           "rank0_roofline": roofline_entry(kern, 1, dp, ab, k, tc[0].value, fp64_peak, hbm_peak),
           "exchange": "k = 64 ghost cells of every state array per side, pf_halo_pack -> NCCL isend/irecv -> pf_halo_unpack, "
                       "once per 64 steps, serial with the block launch" if world > 1 else "single rank: device-local ghost copies only"}
    del grid
    torch.cuda.empty_cache()
    return out


def leg_pic(torch, nat, n, hbm_peak):
    from pyfdtd_b200 import pic
    L, dz, dt = 13194, 8.3276e-5, 2.6389e-13
    z, ux, uz, w = pic.make_beam(n, L, dz, seed=1)
    ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
    del z, ux, uz, w
    Ex = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 1e5
    Hy = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 3e2
    ps.sort()
    reps = 5 if n <= 20_000_000 else 3
    # the rate from a plain event-bracketed run; the per-kernel times from a second, profiled one (an event pair around every
    # launch costs a 1e6-particle step, three 20-us kernels, a quarter of its time)
    sec = time_cuda(torch, lambda: ps.step_sorted(Ex, Hy), 4 * reps, warm=2)
    _, kern = timed_with_kernels(torch, nat, lambda: ps.step_sorted(Ex, Hy), reps, warm=0)
    out = {"particles": n, "grid_cells": L, "particle_steps_per_s": n / sec, "ms_per_step": sec * 1e3,
           "algorithmic_bytes_per_particle_step": 60,
           "hbm": {"algorithmic_GBps": 60.0 * n / sec / 1e9, "frac": 60.0 * n / sec / 1e9 / hbm_peak},
           "kernels": {k: {"ms": v[1] / v[0], "launches_per_step": v[0] / reps} for k, v in kern.items()},
           "step": "pf_pic_step_sorted: Boris push + stable counting re-sort by cell + deterministic deposit (fused)"}
    del ps
    torch.cuda.empty_cache()
    return out


def other_configs(torch, nat, dist, rank, world, args, fp64_peak, hbm_peak):
    """Short measured legs of the other BASELINE configs.  All ranks take part (config 3 is sharded, config 5 decomposed);
    rank 0 returns the dict."""
    from pyfdtd_b200 import MasterController as MC, Environment_Setup as envDef, Solver_Engine as SE
    lib = nat.lib()
    out = {}
    quick = args.quick_extras
    tc = [nat.c_int(), nat.c_int(), nat.c_int()]
    lib.pf_tile_config(*tc)
    tile_cells = tc[0].value

    # ---- config 5 (and 1) : long grids, every N -------------------------------------------------
    per_gpu = 10_000_000 if quick else 100_000_000
    fixed = 40_000_000 if quick else 1_000_000_000
    lg = {}
    lg["weak_lorentz"] = leg_long_grid(torch, nat, dist, rank, world, cells_total=per_gpu * world, steps=128, mode="lorentz",
                                       fp64_peak=fp64_peak, hbm_peak=hbm_peak, label="config 5 weak: 1e8 cells per GPU, Lorentz slab on the right 70 %")
    lg["strong_lorentz_1e9"] = leg_long_grid(torch, nat, dist, rank, world, cells_total=fixed, steps=128, mode="lorentz",
                                             fp64_peak=fp64_peak, hbm_peak=hbm_peak, label="config 5 strong: one fixed 1e9-cell Lorentz grid")
    lg["weak_kerr_lorentz"] = leg_long_grid(torch, nat, dist, rank, world, cells_total=per_gpu * world, steps=64, mode="lorentz_nl",
                                            fp64_peak=fp64_peak, hbm_peak=hbm_peak,
                                            label="config 5 as worded (dispersive AND nonlinear, PF_LORENTZ_NL), weak: 1e8 cells per GPU")
    lg["weak_vacuum"] = leg_long_grid(torch, nat, dist, rank, world, cells_total=per_gpu * world, steps=128, mode="free",
                                      fp64_peak=fp64_peak, hbm_peak=hbm_peak, label="config 1 scaled: vacuum/dielectric grid, 1e8 cells per GPU")
    out["long_grid"] = lg

    # ---- config 3: cubic sweep, 4096 members sharded over the ranks --------------------------------
    M_total, S = (256, 64) if quick else (4096, 128)
    M = M_total // world
    for label, kw in (("nl_cubic_sweep_closed_form", {}), ("nl_cubic_sweep_newton", {"newton": True})):
        batch, t = nl_sweep_batch(M, S, rank=rank, world=world, **kw)
        batch.upload()
        batch.randomize_state(seed=77 + rank)

        def nl_step():
            batch.reset_state(template=True)
            batch.run(do_pol=False)
        sec, kern = timed_with_kernels(torch, nat, nl_step, 2)
        tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec_max = float(tt.cpu()[0])
        slab = float(np.sum(t.mr - t.mf))
        dp = dp_instr_per_step(t.L, t.pw, t.mf, t.mr, "nl") * S
        ab = alg_bytes_per_step(t.L, t.pw, t.mf, t.mr, "nl") * S
        out[label] = {"members_total": M * world, "members_per_gpu": M, "steps": S, "n_gpus": world,
                      "Gcell_updates_per_s": batch.cell_steps * world / sec_max / 1e9,
                      "cubic_solves_per_s": slab * S * world / sec_max,
                      "cubic": "Newton root (PF_F_NEWTON; <= 1e-10 absolute on Acubic)" if kw else "closed form (the reference's algorithm)",
                      "rank0_roofline": roofline_entry(kern, 2, dp, ab, 64, tile_cells, fp64_peak, hbm_peak,
                                                       "slab cells counted at 50 EXECUTED fp64 instructions per cell-step (closed-form law, ncu r2e); Newton variant executes fewer")}
        del batch
        torch.cuda.empty_cache()

    if rank != 0:
        return None

    # ---- config 4: PIC push + cell-sorted deposit at 1e6 / 2e7 / 1e8 particles (replicas only: rank 0 measures) ----
    out["pic"] = {f"{n:.0e}".replace("+0", ""): leg_pic(torch, nat, n, hbm_peak)
                  for n in ((1_000_000, 4_000_000) if quick else (1_000_000, 20_000_000, 100_000_000))}

    # ---- configs 1/2 as the reference runs them: ONE default-geometry run through Controller (host API, e2e) -----
    for name, lor in (("single_run_free_default", False), ("single_run_lorentz_default", True)):
        tup = envDef.envSetup(9e9, 0.7, 7000, 8000, LorMed=lor)
        P = MC.Params(*tup, False, 0.7, 9e9, 20)
        P.TFSF, P.SineCont, P.Periods, P.LorentzMed, P.FreeSpace = True, True, 1000, lor, not lor
        V = MC.Variables(P.Nz, P.timeSteps, P.vidInterval, 10)
        C_P = MC.CPML_Params(P.dz)
        C_V = MC.CPML_Variables(P.Nz, P.timeSteps)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            MC.Controller(V, P, C_V, C_P)
            torch.cuda.synchronize()
            dtc = time.perf_counter() - t0
            best = dtc if best is None else min(best, dtc)
        lib.pf_profile_enable(1)                      # once more with every k_tile launch bracketed by events
        MC.Controller(V, P, C_V, C_P)
        torch.cuda.synchronize()
        lib.pf_profile_enable(0)
        kern = nat.profile_report()
        kname, (kn, kms) = max(kern.items(), key=lambda kv: kv[1][1])
        cu = 2 * P.timeSteps * (P.Nz + 1)
        dp = dp_instr_per_step([P.Nz + 1], [P.pmlWidth], [P.materialFrontEdge], [P.materialRearEdge], "lorentz" if lor else "free")
        out[name] = {"Nz": P.Nz, "timeSteps": P.timeSteps, "passes": 2, "seconds_e2e": best, "Mcell_updates_per_s": cu / best / 1e6,
                     "launches": SE.LAST_RUN_INFO.get("launches"), "kernel": kname, "kernel_launches": kn, "kernel_ms_total": kms,
                     "kernel_ms": kms / kn, "fp64_frac_while_in_kernel": dp * 2 * P.timeSteps / (kms * 1e-3) / fp64_peak,
                     "note": "Controller() incl. host setup, H2D/D2H; snapshots every 50 steps stay on the device until "
                             "V.Ex_History is read; one grid = 15 CTAs: bound by the per-step dependency chain, not by any pipe"}

    # ---- the headline workload in the optional arithmetic modes ------------------------------------------------
    if not quick:
        modes = {}
        for label, fma, fp32 in (("fma_contracted", True, False), ("fp32", False, True)):
            b2, t2 = lorentz_sweep_batch(args.members, args.pass_steps, args.n_freq, fma=fma, fp32=fp32)
            b2.upload()
            b2.randomize_state(seed=1234)

            def sweep_step():
                b2.reset_state(template=True)
                b2.run(do_pol=True)
            sec, kern = timed_with_kernels(torch, nat, sweep_step, 3)
            kname, (kn, kms) = max(kern.items(), key=lambda kv: kv[1][1])
            modes[label] = {"Gcell_updates_per_s": b2.cell_steps / sec / 1e9, "kernel": kname, "kernel_ms": kms / kn}
            del b2
            torch.cuda.empty_cache()
        out["lorentz_sweep_optional_modes_Gcell_updates_per_s"] = dict(
            modes, note="same workload as `value`; fma: PF_F_FMA (<= 1e-10 relative, tested); fp32: PF_F_FP32 (optional mode, "
                        "tolerance stated in include/pyfdtd_b200.h)")
    return out


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fma", action="store_true", help="PF_F_FMA kernels (not bit-identical; reported in config)")
    ap.add_argument("--fp32", action="store_true", help="PF_F_FP32 kernels (optional single-precision mode; reported as dtype f32)")
    ap.add_argument("--no-extras", action="store_true", help="skip the legs of the other BASELINE configs and the full sweep")
    ap.add_argument("--quick-extras", action="store_true", help="smaller extras (smoke)")
    ap.add_argument("--full-sweep-members", type=int, default=1024, help="distinct grids per GPU of the e2e_full_sweep leg")
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
    import pyfdtd_b200  # noqa: F401
    from pyfdtd_b200 import _native as nat, sweep
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = nat.lib()
    t_setup0 = time.perf_counter()
    batch, table = lorentz_sweep_batch(args.members, args.pass_steps, args.n_freq, fma=args.fma, fp32=args.fp32)
    setup_s = time.perf_counter() - t_setup0
    kcfg = [nat.c_int(), nat.c_int(), nat.c_int()]
    lib.pf_tile_config(*kcfg)
    tile_cells = kcfg[0].value
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
    # the pipe this kernel is bound by, measured now on this GPU
    fp64_peak, dfma_peak = nat.probe_fp64()
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
    kernels = nat.profile_report()

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
    batch_b = sweep.MemberBatch.from_table(table, "lorentz", T_alloc=args.pass_steps)
    batch_b.upload()
    batch_b.state_template = batch.state_template
    pipe = sweep.BatchPipeline([batch, batch_b])
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
    cell_steps = batch.cell_steps
    total_cell_steps = cell_steps * world * args.steps
    value = total_cell_steps / (ms_total * 1e-3) / 1e9
    e2e_value = total_cell_steps / (e2e_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"]
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    peak_instr = dfma_peak if (args.fma or args.fp32) else fp64_peak
    dp_step = dp_instr_per_step(table.L, table.pw, table.mf, table.mr, "lorentz") * args.pass_steps
    ab_step = alg_bytes_per_step(table.L, table.pw, table.mf, table.mr, "lorentz") * args.pass_steps
    rl = roofline_entry(kernels, args.steps, dp_step, ab_step, k_block, tile_cells, peak_instr, hbm_peak)
    n_launch_per_step = rl["kernel_launches_per_call"]
    wl_key = f"lorentz_sweep_m{args.members}_f{args.n_freq}_k{k_block}_" + ("fp32" if args.fp32 else "fma" if args.fma else "exact")
    traffic, traffic_src = committed_ncu_traffic(wl_key)
    roofline = {
        "bound": "fp64", "achieved": rl["fp64"]["achieved_dp_instr_per_s"] / 1e9, "peak": peak_instr / 1e9, "unit": "G fp64 instr/s",
        "frac": rl["fp64"]["frac"], "traffic": traffic, "traffic_source": traffic_src,
        "kernel": rl["kernel"], "kernel_ms_avg": rl["kernel_ms"], "kernel_launches_timed": int(round(n_launch_per_step * args.steps)),
        "kernel_share_of_step": rl["kernel_ms_per_call"] * args.steps / ms_total if ms_total else None,
        "peak_source": "measured in this run on this GPU (pf_probe_fp64: independent separately rounded DMUL+DADD streams, all SMs)"
                       if not (args.fma or args.fp32) else "measured in this run (pf_probe_fp64: DFMA streams)",
        "algorithmic_dp_instr_per_launch": dp_step / n_launch_per_step,
        "how": "achieved = fp64 instructions of the reference's own arithmetic (exact mode: every multiply and add separately "
               "rounded; halo recomputation NOT counted) per launch / CUDA-event launch time",
        "hbm": dict(rl["hbm"], peak_GBps=hbm_peak, peak_source=hbm_src, algorithmic_bytes_per_launch=ab_step / n_launch_per_step,
                    note="the kernel is temporally blocked (k steps per HBM round trip): k = 1 algorithmic bytes / time exceed the "
                         "HBM roofline by design; modelled_real = algorithmic/k x (1 + 2k/(tile - 2k)) is what a launch moves"),
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sec, cells = cpu_port_rate(args.members, args.pass_steps, threads, args.n_freq)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"one full bench step ({args.members} members x {args.pass_steps} steps, {cells/1e9:.2f} Gcell-updates) of "
                         f"the C port of the reference loop on {threads} threads ({sec:.2f} s wall, after one warm-up step)"}

    cfg_state_mb, h2d_b, d2h_b = batch.n_state * 8 / 1e6, batch.h2d_bytes, batch.d2h_bytes
    del batch
    torch.cuda.empty_cache()

    # ---- the whole product path on distinct grids: setup + H2D + 2 passes x full T + reflection ------------
    full = None
    extra = None
    if not args.no_extras:
        nfull = 16 if args.quick_extras else args.full_sweep_members
        freqs = np.linspace(6e9, 10.5e9, nfull * world)
        best = None
        for rep in range(2):       # the second call reuses the pinned staging slots (steady state of a sweep service)
            barrier()
            t0 = time.perf_counter()
            res = sweep.reflection_sweep(freqs, DOM, *WIN, periods=1.0, rank=rank, world_size=world)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw.cpu()[0])
            if best is None or wall < best[0]:
                best = (wall, res, rep)
        wall, res, rep = best
        full = {"value": res["cell_steps"] * world / wall / 1e9, "unit": UNIT, "seconds": wall, "members_per_gpu": nfull,
                "distinct_grids": True, "host_setup_s": res["timing"]["setup_s"], "build_inputs_s": res["timing"]["build_inputs_s"],
                "host_threads": os.cpu_count(), "cell_updates": res["cell_steps"] * world,
                "R_first_last": [float(res["measured"][0]), float(res["measured"][-1])],
                "how": "sweep.reflection_sweep(freqs): envSetup_many + spatialStab chain + tables (host, vectorised), "
                       "pf_host_sweep_inputs (CPML profiles + source tables into pinned memory, all host threads), H2D, "
                       "2 passes x full timeSteps per member (pf_run_batch, chunks of 256 members pipelined against the host "
                       "work of the next chunk), batched FFT reflection extraction on the device, one D2H of R per member; "
                       "wall clock of the whole call, best of 2 calls"}
        try:
            extra = other_configs(torch, nat, dist, rank, world, args, fp64_peak, hbm_peak)
        except Exception as e:   # the headline line must survive a failing extra
            import traceback
            extra = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": workload_config(args, "fp32 on chip (PF_F_FP32; not a parity mode)" if args.fp32 else
                                      "fma-contracted" if args.fma else "exact (bit-identical to reference order)"),
            "run": {"cell_updates_per_step_per_gpu": cell_steps, "k_block": k_block,
                    "l2_policy": f"state {cfg_state_mb:.0f} MB per GPU > 126 MB L2, restored from a random template every step",
                    "parallelism": f"members sharded over {world} GPU(s), no collectives", "batch_setup_s": setup_s},
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
        if full is not None:
            line["e2e_full_sweep"] = full
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["long_grid"] = extra.pop("long_grid", None) if isinstance(extra, dict) else None
            line["other_configs"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
