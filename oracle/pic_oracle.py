"""CPU ORACLE for the PIC push / sort / deposit -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Parity status: UNPINNED by the reference -- the reference contains no particle code at all (SURVEY.md
F2); its only PIC contract is the current slot V.Jx that ADE_ExUpdate subtracts (BaseFDTD11.py:667).
This NumPy file is therefore the *definition* of the model the CUDA kernels (csrc/pf_pic.cu) implement
(spec in DESIGN.md, section PIC); the GPU is compared with it at 1e-12 relative (bit-exact expected:
every operation below is a single IEEE fp64 op evaluated in the same order as in the kernel, which is
built with --fmad=false).
"""
from __future__ import annotations

import numpy as np


def push(z, ux, uz, Ex, Hy, *, dz, dt, q_over_m, c, mu0):
    """Relativistic Boris push with linear gather; specular walls.  Returns new (z, ux, uz, cell)."""
    L = len(Ex)
    inv_dz = 1.0 / dz
    zmax = float(L - 1) * dz
    s = z * inv_dz
    ci = np.clip(np.floor(s).astype(np.int64), 0, L - 2)
    f = s - ci
    E = (1.0 - f) * Ex[ci] + f * Ex[ci + 1]
    sh = s - 0.5
    ch = np.clip(np.floor(sh).astype(np.int64), 0, L - 2)
    fh = np.minimum(np.maximum(sh - ch, 0.0), 1.0)
    By = mu0 * ((1.0 - fh) * Hy[ch] + fh * Hy[ch + 1])
    qmdt2 = q_over_m * dt * 0.5
    inv_c2 = 1.0 / (c * c)
    uxm = ux + qmdt2 * E
    uzm = uz
    gm = np.sqrt(1.0 + (uxm * uxm + uzm * uzm) * inv_c2)
    t = qmdt2 * By / gm
    sfac = 2.0 * t / (1.0 + t * t)
    uxp = uxm - uzm * t
    uzp = uzm + uxm * t
    uxn = uxm - uzp * sfac
    uzn = uzm + uxp * sfac
    uxn = uxn + qmdt2 * E
    g = np.sqrt(1.0 + (uxn * uxn + uzn * uzn) * inv_c2)
    zn = z + (uzn / g) * dt
    lo = zn < 0.0
    zn = np.where(lo, -zn, zn)
    uzn = np.where(lo, -uzn, uzn)
    hi = zn > zmax
    zn = np.where(hi, 2.0 * zmax - zn, zn)
    uzn = np.where(hi, -uzn, uzn)
    zn = np.minimum(np.maximum(zn, 0.0), zmax)
    cell = np.clip(np.floor(zn * inv_dz).astype(np.int64), 0, L - 2).astype(np.int32)
    return zn, uxn, uzn, cell


def sort_by_cell(z, ux, uz, w, cell):
    """Stable sort by cell index."""
    order = np.argsort(cell, kind="stable")
    return z[order], ux[order], uz[order], w[order], cell[order]


def deposit(z, ux, uz, w, cell, L, *, dz, c, jx_scale):
    """Deterministic cell-sorted CIC deposition, same summation tree as k_pic_cell_sums/k_pic_flush:
    lane l of the cell's warp adds particles l, l+32, ... in order; lanes are combined by the xor
    butterfly 16,8,4,2,1; node nz = jx_scale * (left share of cell nz + right share of cell nz-1)."""
    assert np.all(np.diff(cell) >= 0), "particles must be sorted by cell"
    inv_dz = 1.0 / dz
    inv_c2 = 1.0 / (c * c)
    g = np.sqrt(1.0 + (ux * ux + uz * uz) * inv_c2)
    wv = w * (ux / g)
    f = z * inv_dz - cell
    t0 = wv * (1.0 - f)
    t1 = wv * f
    acc = np.zeros((L, 2))
    starts = np.searchsorted(cell, np.arange(L), side="left")
    ends = np.searchsorted(cell, np.arange(L) + 1, side="left")
    lanes = np.arange(32)
    for cidx in range(L):
        a, b = starts[cidx], ends[cidx]
        if a == b:
            continue
        n = b - a
        pad = (-n) % 32
        for k, t in enumerate((t0, t1)):
            rows = np.concatenate([t[a:b], np.zeros(pad)]).reshape(-1, 32)
            lane = np.zeros(32)
            for r in rows:
                lane = lane + r
            for off in (16, 8, 4, 2, 1):
                lane = lane + lane[lanes ^ off]
            acc[cidx, k] = lane[0]
    J = acc[:, 0].copy()
    J[1:] = J[1:] + acc[:-1, 1]
    return jx_scale * J


def sub_warps(n, L, sub_max=8):
    """pic_sub_warps() of csrc/pf_pic.cu: pieces (warps) per cell of the fused push + re-sort (+ deposit)."""
    per_cell = n // max(1, L)
    return int(min(sub_max, max(1, (per_cell + 255) // 256)))


def deposit_fused(z, ux, uz, w, cell_old, cell_new, L, S, *, dz, c, jx_scale):
    """Summation tree of pf_pic_step_sorted (k_pic_move<true> + k_pic_flush4).  Arrays are the PUSHED particles in their
    OLD order (sorted by cell_old).  Sub-warp s of old cell cc owns a contiguous piece of the cell's particles (a multiple
    of 32 long); lane l adds, in chunk order, the CIC shares of particles l, l+32, ... of the piece at the four nodes
    cc-1 .. cc+2; lanes are combined by the xor butterfly 16,8,4,2,1; node nz = jx_scale * (sum over old cells
    nz-2 .. nz+1 ascending, sub-warps ascending, of the partial sum that cell holds for nz)."""
    assert np.all(np.diff(cell_old) >= 0), "particles must be sorted by their old cell"
    inv_dz = 1.0 / dz
    inv_c2 = 1.0 / (c * c)
    g = np.sqrt(1.0 + (ux * ux + uz * uz) * inv_c2)
    wv = w * (ux / g)
    f = z * inv_dz - cell_new
    t0 = wv * (1.0 - f)
    t1 = wv * f
    starts = np.searchsorted(cell_old, np.arange(L + 1), side="left")
    part = np.zeros((L, S, 4))
    lanes = np.arange(32)
    for cc in range(L):
        a0, b0 = int(starts[cc]), int(starts[cc + 1])
        if a0 == b0:
            continue
        q = (((b0 - a0) + S - 1) // S + 31) // 32 * 32
        for s in range(S):
            lo = min(b0, a0 + s * q)
            hi = min(b0, lo + q)
            if lo == hi:
                continue
            d = np.clip(cell_new[lo:hi].astype(np.int64) - cc, -1, 1)
            n = hi - lo
            con = np.zeros((n, 4))
            for dd in (-1, 0, 1):
                m = d == dd
                con[m, dd + 1] = t0[lo:hi][m]
                con[m, dd + 2] = t1[lo:hi][m]
            rows = np.concatenate([con, np.zeros(((-n) % 32, 4))]).reshape(-1, 32, 4)
            lane = np.zeros((32, 4))
            for r in rows:
                lane = lane + r
            for off in (16, 8, 4, 2, 1):
                lane = lane + lane[lanes ^ off]
            part[cc, s] = lane[0]
    v = np.zeros(L)
    for dc in (-2, -1, 0, 1):
        k = 1 - dc
        src = np.arange(L) + dc
        ok = (src >= 0) & (src < L)
        for s in range(S):
            add = np.zeros(L)
            add[ok] = part[src[ok], s, k]
            v = v + add
    return jx_scale * v


def make_beam(n, L, dz, *, gamma=1.2, thermal=0.01, c=299792458.0, seed=1234, zlo=0.05, zhi=0.95):
    """uniform in z, beam gamma along z with a relative thermal spread."""
    rng = np.random.default_rng(seed)
    zmax = (L - 1) * dz
    z = rng.uniform(zlo * zmax, zhi * zmax, n)
    u0 = c * np.sqrt(gamma * gamma - 1.0)
    uz = u0 * (1.0 + thermal * rng.standard_normal(n))
    ux = u0 * thermal * rng.standard_normal(n)
    w = np.full(n, 1.0e10)
    cell = np.clip(np.floor(z * (1.0 / dz)).astype(np.int64), 0, L - 2).astype(np.int32)
    return z, ux, uz, w, cell
