"""Electron-beam PIC coupled to the 1-D FDTD grid through the reference's Jx slot.

The reference has no particle code (only the title and TODOs) but its E update already
subtracts a current slot: ``Ex += (Hy[nz]-Hy[nz-1]-Jx[nz]) * UpExMat * den`` (BaseFDTD11.py:667) with
``V.Jx`` identically zero.  This module supplies what fills that slot: a relativistic Boris push with
linear gather, a stable sort by cell, and a deterministic cell-sorted CIC deposition (csrc/pf_pic.cu;
model spec in DESIGN.md, CPU definition in oracle/pic_oracle.py).  Units: positions in metres,
momenta u = gamma*v in m/s, weights = real particles per macro-particle per unit transverse area, so
that with ``jx_scale = q`` the deposited array is J_phys*dz -- the unit the slot is subtracted in.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat

ELECTRON_Q = -1.602176634e-19
ELECTRON_Q_OVER_M = -1.75882001076e11
C0 = 299792458.0
MU0 = 1.25663706127e-06


def make_beam(n, L, dz, *, gamma=1.2, thermal=0.01, c=C0, seed=1234, zlo=0.05, zhi=0.95, weight=1.0e10):
    """Synthetic electron beam of SURVEY 8(d) config 4: n macro-particles uniform in z over [zlo, zhi] of the grid,
    drifting along z with Lorentz factor ``gamma`` and a relative thermal spread (normal) in both momenta.
    Returns host arrays (z, ux, uz, w)."""
    rng = np.random.default_rng(seed)
    zmax = (L - 1) * dz
    z = rng.uniform(zlo * zmax, zhi * zmax, n)
    u0 = c * np.sqrt(gamma * gamma - 1.0)
    uz = u0 * (1.0 + thermal * rng.standard_normal(n))
    ux = u0 * thermal * rng.standard_normal(n)
    return z, ux, uz, np.full(n, weight)


class ParticleSet:
    """Device-resident SoA particle arrays (double-buffered for the sort) + the PfPic descriptor."""

    def __init__(self, z, ux, uz, w, L, dz, dt, *, q_over_m=ELECTRON_Q_OVER_M, jx_scale=ELECTRON_Q, c=C0, mu0=MU0,
                 device=None):
        torch = nat.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.n, self.L, self.dz, self.dt = int(len(z)), int(L), float(dz), float(dt)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.cur = {k: torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64), **f64)
                    for k, v in (("z", z), ("ux", ux), ("uz", uz), ("w", w))}
        self.alt = {k: torch.empty_like(v) for k, v in self.cur.items()}
        cell = np.clip(np.floor(np.asarray(z) * (1.0 / dz)).astype(np.int64), 0, L - 2).astype(np.int32)
        self.cell = torch.as_tensor(cell, device=self.device)
        self.cell_alt = torch.empty_like(self.cell)
        self.Jx = torch.zeros(L, **f64)
        self.p = nat.PfPic()
        self.p.n, self.p.L = self.n, self.L
        self.p.dz, self.p.dt, self.p.q_over_m, self.p.c, self.p.mu0, self.p.jx_scale = dz, dt, q_over_m, c, mu0, jx_scale
        self._bind()
        sb = nat.lib().pf_pic_scratch_bytes(ctypes.byref(self.p))
        self.scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
        self.scratch_bytes = sb
        self.sorted = False
        self._offsets_valid = False    # the scratch holds the cell offsets of the current arrays (PF_PIC_F_OFFSETS_VALID)

    def _bind(self):
        p = self.p
        p.z, p.ux, p.uz, p.w = (self.cur[k].data_ptr() for k in ("z", "ux", "uz", "w"))
        p.z_alt, p.ux_alt, p.uz_alt, p.w_alt = (self.alt[k].data_ptr() for k in ("z", "ux", "uz", "w"))
        p.cell, p.cell_alt = self.cell.data_ptr(), self.cell_alt.data_ptr()
        p.Jx = self.Jx.data_ptr()

    def push(self, Ex, Hy):
        """Advance every particle one step in the fields Ex, Hy (device tensors of length L)."""
        self.p.Ex, self.p.Hy = Ex.data_ptr(), Hy.data_ptr()
        nat.check(nat.lib().pf_pic_push(ctypes.byref(self.p), nat.current_stream_ptr()), "pf_pic_push")
        self.sorted = False
        self._offsets_valid = False

    def push_sorted(self, Ex, Hy):
        """Push + stable re-sort in one (requires a cell-sorted set; sorts once if it is not).  Faster than
        push() + sort(): no radix sort, no gather -- particles only ever move to a neighbouring cell."""
        if not self.sorted:
            self.sort()
        self.p.Ex, self.p.Hy = Ex.data_ptr(), Hy.data_ptr()
        self.p.flags = nat.PF_PIC_F_OFFSETS_VALID if self._offsets_valid else 0
        nat.check(nat.lib().pf_pic_push_sorted(ctypes.byref(self.p), self.scratch.data_ptr(), self.scratch_bytes,
                                               nat.current_stream_ptr()), "pf_pic_push_sorted")
        self.cur, self.alt = self.alt, self.cur
        self.cell, self.cell_alt = self.cell_alt, self.cell
        self._bind()
        self.sorted = True
        self._offsets_valid = True

    def step_sorted(self, Ex, Hy):
        """One whole particle step: push + stable re-sort + deposition of the new state, fused (pf_pic_step_sorted).
        Returns the Jx tensor.  Same particles as push_sorted(); Jx equals deposit()'s to rounding (different, equally
        deterministic summation tree)."""
        if not self.sorted:
            self.sort()
        self.p.Ex, self.p.Hy = Ex.data_ptr(), Hy.data_ptr()
        self.p.flags = nat.PF_PIC_F_OFFSETS_VALID if self._offsets_valid else 0
        nat.check(nat.lib().pf_pic_step_sorted(ctypes.byref(self.p), self.scratch.data_ptr(), self.scratch_bytes,
                                               nat.current_stream_ptr()), "pf_pic_step_sorted")
        self.cur, self.alt = self.alt, self.cur
        self.cell, self.cell_alt = self.cell_alt, self.cell
        self._bind()
        self.sorted = True
        self._offsets_valid = True
        return self.Jx

    def sub_warps(self):
        return int(nat.lib().pf_pic_sub_warps(ctypes.byref(self.p)))

    def cfl_violated(self):
        return bool(nat.check(nat.lib().pf_pic_check(ctypes.byref(self.p), self.scratch.data_ptr(), self.scratch_bytes,
                                                     nat.current_stream_ptr()), "pf_pic_check"))

    def sort(self):
        """Stable sort by cell; swaps the double buffers."""
        nat.check(nat.lib().pf_pic_sort(ctypes.byref(self.p), self.scratch.data_ptr(), self.scratch_bytes,
                                        nat.current_stream_ptr()), "pf_pic_sort")
        self.cur, self.alt = self.alt, self.cur
        self.cell, self.cell_alt = self.cell_alt, self.cell
        self._bind()
        self.sorted = True
        self._offsets_valid = False

    def deposit(self):
        """Deposit Jx (requires cell-sorted particles; sorts first if needed).  Returns the Jx tensor."""
        if not self.sorted:
            self.sort()
        nat.check(nat.lib().pf_pic_deposit(ctypes.byref(self.p), self.scratch.data_ptr(), self.scratch_bytes,
                                           nat.current_stream_ptr()), "pf_pic_deposit")
        return self.Jx

    def host(self):
        return {k: v.cpu().numpy() for k, v in self.cur.items()} | {"cell": self.cell.cpu().numpy()}


def coupled_grid(V, P, C_V, C_P, Exs, Hys, mode="free", probe_idx=()):
    """DeviceGrid of one prepared pass (Solver_Engine.prepare_pass has run) for CoupledPIC: per-cell coefficient arrays for
    the per-op engine plus, when they have the canonical piecewise form, the scalars and flag the fused tile engine needs."""
    from . import BaseFDTD11, _device as dev
    arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
    scal = BaseFDTD11.grid_scalars(V, P, kerr_lorentz=(mode == "lorentz_nl"))
    flags = BaseFDTD11.grid_flags(P)
    canon = dev.canonical_form(P, arrs)
    slab_src_clash = mode != "free" and P.TFSF and (P.materialFrontEdge - 1 <= P.nzsrc - 1 < P.materialRearEdge)
    if canon is not None and not slab_src_clash and dev.probes_ok_for_tiles(list(probe_idx)):
        scal.update(cE0=canon[0], cE1=canon[1], cH0=canon[2], cH1=canon[3], c2_pml=canon[4])
        flags |= nat.PF_F_CANONICAL
    return dev.DeviceGrid(L=len(V.Ex), T=int(P.timeSteps), arrays=arrs, scalars=scal, srcE=np.asarray(Exs) / P.courantNo,
                          srcH=np.asarray(Hys) / P.courantNo, probe_idx=list(probe_idx), flags=flags)


class CoupledPIC:
    """PIC step coupled to one FDTD grid: deposit -> field step -> push.

    The deposited current enters the field step through the Jx slot: ADE_ExUpdate subtracts it (BaseFDTD11.py:667) and,
    inside the slab [mf, mr) -- where the reference's loops overwrite Ex from Dx -- ADE_DxUpdate subtracts it the same way
    (dD/dt = curl H - J; builder-defined, see include/pyfdtd_b200.h), so a beam inside the Lorentz / cubic medium drives
    the fields there too.  ``grid`` is a _device.DeviceGrid (see ``coupled_grid``) whose descriptor gets its Jx pointer from
    the particle set.  engine "auto": the fused tile kernel (one launch per field step, k_tile<..., JX>) when the grid
    carries PF_F_CANONICAL and the mode is free / lorentz / nl, else the per-op engine (one kernel per leaf op)."""

    def __init__(self, grid, particles, mode="free", fused=False, engine="auto"):
        from . import _device as dev
        self.grid, self.particles = grid, particles
        self.mode_id = dev.MODE_ID[mode]
        self.grid.g.Jx = particles.Jx.data_ptr()
        self.n = 0
        self.fused = fused          # True: push + re-sort + deposit in one pass over the particles (step_sorted)
        self._have_J = False
        tile_ok = bool(grid.g.flags & nat.PF_F_CANONICAL) and mode in ("free", "lorentz", "nl")
        if engine == "tile" and not tile_ok:
            raise ValueError("CoupledPIC: engine='tile' needs a canonical grid (pic.coupled_grid) in mode free / lorentz / nl")
        self.engine = nat.PF_ENGINE_TILE if (tile_ok and engine != "ops") else nat.PF_ENGINE_OPS
        self.scratch, self.scratch_bytes = None, 0
        if self.engine == nat.PF_ENGINE_TILE:
            self.scratch_bytes = nat.lib().pf_run_scratch_bytes(grid.ref(), 1, nat.PF_ENGINE_TILE)
            self.scratch = particles.torch.empty(self.scratch_bytes, dtype=particles.torch.uint8, device=particles.device)

    def step(self, do_pol=False):
        lib = nat.lib()
        if self.n >= self.grid.T:
            raise ValueError(f"CoupledPIC.step: step {self.n} is past the grid's source tables (T = {self.grid.T})")
        if not (self.fused and self._have_J):
            self.particles.deposit()
        nat.check(lib.pf_run_pass(self.grid.ref(), self.mode_id, int(do_pol), self.n, 1, self.engine, None, 0, 0,
                                  self.scratch.data_ptr() if self.scratch is not None else None, self.scratch_bytes,
                                  nat.current_stream_ptr()), "pf_run_pass")
        if self.fused:               # the deposit of the pushed state is the current of the next field step
            self.particles.step_sorted(self.grid.tensor_view("Ex"), self.grid.tensor_view("Hy"))
            self._have_J = True
        else:
            self.particles.push_sorted(self.grid.tensor_view("Ex"), self.grid.tensor_view("Hy"))
        self.n += 1
