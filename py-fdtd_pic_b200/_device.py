"""Host <-> device marshalling for one grid: builds the C-ABI ``PfGrid`` descriptor from the
reference-style state objects (V, P, C_V, C_P).

All arrays of a grid are packed into ONE pinned host staging buffer and ONE device pool, so a grid
costs one H2D and one D2H transfer (torch tensors are only the buffer carrier; kernels see raw
pointers).  ``canonical_form`` performs the bit-for-bit check that lets the fused tile engine replace
the per-cell coefficient arrays by scalars (PF_F_CANONICAL in include/pyfdtd_b200.h).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat

STATE = ("Ex", "Hy", "Dx", "P", "Pprev", "psiE", "psiH", "Acubic")
COEF = ("UpExMat", "denE", "UpHySelf", "UpHyMat", "denH", "beX", "ceX", "Cb", "bmY", "cmY", "C2")
MODE_ID = {"free": nat.PF_FREE, "lorentz": nat.PF_LORENTZ, "nl": nat.PF_NL, "lorentz_nl": nat.PF_LORENTZ_NL}


def canonical_form(P, arrs, Jx=None):
    """Return (cE0, cE1, cH0, cH1, c2) if the coefficient arrays have the tile engine's piecewise
    form, else None.  Every comparison is exact (bitwise for finite values)."""
    L = len(arrs["Ex"])
    pw, mf, mr = int(P.pmlWidth), int(P.materialFrontEdge), int(P.materialRearEdge)
    # (a current slot Jx does not affect the form of the coefficient arrays: the tile engine carries it as one more
    #  per-cell input, see k_tile<..., JX>)
    for k in ("denE", "denH", "UpHySelf"):
        if not np.all(arrs[k] == 1.0):
            return None
    if not np.array_equal(arrs["bmY"], arrs["beX"]):
        return None
    inside = np.zeros(L, dtype=bool)
    inside[max(mf, 0):max(min(mr, L), 0)] = True
    out = []
    for k in ("UpExMat", "UpHyMat"):
        a = arrs[k]
        v0 = a[~inside][0] if np.any(~inside) else a[0]
        v1 = a[inside][0] if np.any(inside) else v0
        if not (np.all(a[~inside] == v0) and np.all(a[inside] == v1)):
            return None
        out += [float(v0), float(v1)]
    c2 = float(arrs["C2"][1]) if pw > 1 else 0.0
    if (P.CPMLXm or P.CPMLXp) and pw > 0:
        if 2 * pw > L:
            return None
        corr = np.zeros(L, dtype=bool)          # cells whose field correction is applied
        if P.CPMLXm:
            corr[1:pw] = True
        if P.CPMLXp:
            corr[L - pw + 1:L] = True
        if not np.array_equal(arrs["Cb"][corr], arrs["UpExMat"][corr]):
            return None
        corrH = corr.copy()
        corrH[L - 1] = False
        if not np.all(arrs["C2"][corrH] == c2):
            return None
        if P.CPMLXp and not (arrs["Cb"][L - pw] == 0.0 and arrs["C2"][L - pw] == 0.0):
            return None
    return (*out, c2)


def probes_ok_for_tiles(probe_idx, cells_per_thread=8):
    """The tile kernel lets one thread own at most two probe cells (a thread owns <= 8 consecutive cells)."""
    p = np.sort(np.asarray(probe_idx, dtype=np.int64))
    return len(p) < 3 or bool(np.all(p[2:] - p[:-2] >= cells_per_thread))


_PINNED = {}


def pinned_buffer(n_doubles, tag="stage"):
    """Grow-only cache of pinned host staging buffers (allocating pinned memory costs milliseconds and a
    single run would otherwise do it every pass).  Callers must finish their copies before the next use."""
    torch = nat.require_cuda()
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < n_doubles:
        buf = torch.empty(int(n_doubles * 1.25) + 1024, dtype=torch.float64).pin_memory()
        _PINNED[tag] = buf
    return buf[:n_doubles]


class DeviceGrid:
    """Device-resident copy of one grid + its PfGrid descriptor."""

    def __init__(self, *, L, T, arrays, scalars, srcE, srcH, probe_idx, flags, Jx=None, z0=0, Lg=None,
                 need=None, device=None, stage_tag="stage"):
        torch = nat.require_cuda()
        self.torch = torch
        self.L, self.T = int(L), int(T)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        names = list(STATE) + list(COEF) if need is None else list(need)
        if Jx is not None:
            names.append("Jx")
            arrays = dict(arrays, Jx=Jx)
        Lp = (self.L + 31) // 32 * 32                       # 256-byte aligned rows
        Tp = (self.T + 31) // 32 * 32
        n_probe = len(probe_idx)
        self.names = names
        self.off = {}
        cur = 0
        for n in names:
            self.off[n] = cur
            cur += Lp
        for n in ("srcE", "srcH"):
            self.off[n] = cur
            cur += Tp
        self.off["probe_out"] = cur
        cur += max(n_probe, 1) * Tp
        self.n_doubles = cur
        self.Tp = Tp
        self.host = pinned_buffer(cur, tag=stage_tag)   # grids alive at the same time need their own staging buffers
        hv = self.host.numpy()
        hv[:] = 0.0
        for n in names:
            a = arrays.get(n)
            if a is not None:
                hv[self.off[n]: self.off[n] + self.L] = a
        hv[self.off["srcE"]: self.off["srcE"] + len(srcE)] = srcE
        hv[self.off["srcH"]: self.off["srcH"] + len(srcH)] = srcH
        self.pool = torch.empty(cur, dtype=torch.float64, device=self.device)
        self.pool.copy_(self.host, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the staging buffer is shared: the copy must have left it
        self.h2d_bytes = cur * 8
        self.probe_idx_t = torch.tensor(list(probe_idx) or [0], dtype=torch.int32, device=self.device)
        self.h2d_bytes += 4 * max(n_probe, 1)

        g = nat.PfGrid()
        g.L, g.pw, g.mf, g.mr, g.nzsrc = self.L, scalars["pw"], scalars["mf"], scalars["mr"], scalars["nzsrc"]
        g.flags = flags
        g.n_probes, g.probe_stride = n_probe, Tp
        g.n_src = min(len(srcE), len(srcH))
        g.z0, g.Lg = z0, self.L if Lg is None else Lg
        for k in ("dt_over_dz", "eps0", "polA", "polB", "polC", "cub_a", "cub_b", "cub_c", "nl_den0", "nl_den1",
                  "cE0", "cE1", "cH0", "cH1", "c2_pml"):
            setattr(g, k, float(scalars.get(k, 0.0)))
        base = self.pool.data_ptr()
        for n in list(STATE) + list(COEF) + ["Jx"]:
            setattr(g, n, base + 8 * self.off[n] if n in self.off else None)
        g.srcE = base + 8 * self.off["srcE"]
        g.srcH = base + 8 * self.off["srcH"]
        g.probe_idx = self.probe_idx_t.data_ptr()
        g.probe_out = base + 8 * self.off["probe_out"]
        self.g = g
        self.d2h_bytes = 0

    def ref(self):
        return ctypes.byref(self.g)

    def fetch(self, names, probes=True):
        """One D2H of the whole pool, then slice.  Returns dict name -> fresh numpy array."""
        torch = self.torch
        self.host.copy_(self.pool, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes += self.n_doubles * 8
        hv = self.host.numpy()
        out = {n: hv[self.off[n]: self.off[n] + self.L].copy() for n in names if n in self.off}
        if probes:
            n_p = self.g.n_probes
            po = hv[self.off["probe_out"]: self.off["probe_out"] + max(n_p, 1) * self.Tp]
            out["probe_out"] = po.reshape(max(n_p, 1), self.Tp)[:n_p, : self.T].copy()
        return out

    def fetch_probes(self):
        """D2H of the probe traces only -> [n_probes, T]."""
        torch = self.torch
        n_p = max(self.g.n_probes, 1)
        lo = self.off["probe_out"]
        self.host[lo: lo + n_p * self.Tp].copy_(self.pool[lo: lo + n_p * self.Tp], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes += n_p * self.Tp * 8
        po = self.host.numpy()[lo: lo + n_p * self.Tp]
        return po.reshape(n_p, self.Tp)[: self.g.n_probes, : self.T].copy()

    def tensor_view(self, name):
        return self.pool[self.off[name]: self.off[name] + self.L]
