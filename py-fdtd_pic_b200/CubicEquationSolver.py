"""Closed-form cubic roots (Cardano / trigonometric form).

Mirror of the reference's vendored ``CubicEquationSolver`` (CubicEquationSolver.py:29-105, the
1728.org algorithm): ``solve(a, b, c, d)`` returns the roots of ``a x^3 + b x^2 + c x + d`` in the
reference's order -- root[0] is the one the nonlinear path consumes (BaseFDTD11.py:836-838).  The
arithmetic runs in the ``k_cubic_solve`` CUDA kernel through ``pf_cubic_solve``; ``solve_many`` is the
batched form the hot path wants (one polynomial per slab cell per step, evidenced as the reference's
known bottleneck by benchmarkTests.py).  ``CubicSolver`` is the 4-tuple packer that
BaseFDTD11.Nonlin_Cubic_Solver calls but the reference never shipped.
"""
from __future__ import annotations

import numpy as np

from . import _native as nat


def CubicSolver(a, b, c, d):
    return tuple(float(np.real(x)) for x in (a, b, c, d))


def solve_many(coeffs):
    """coeffs: array [n, 4] -> (roots complex128 [n, 3] NaN-padded, nroots int32 [n])."""
    torch = nat.require_cuda()
    co = torch.as_tensor(np.ascontiguousarray(coeffs, dtype=np.float64), device="cuda").reshape(-1, 4)
    n = co.shape[0]
    roots = torch.zeros((n, 3, 2), dtype=torch.float64, device="cuda")
    nroots = torch.zeros(n, dtype=torch.int32, device="cuda")
    nat.check(nat.lib().pf_cubic_solve(co.data_ptr(), roots.data_ptr(), nroots.data_ptr(), n,
                                       nat.current_stream_ptr()), "pf_cubic_solve")
    r = roots.cpu().numpy()
    nr = nroots.cpu().numpy()
    out = r[..., 0] + 1j * r[..., 1]
    out[np.arange(3)[None, :] >= nr[:, None]] = np.nan
    return out, nr


def root0_many(coeffs, newton=False):
    """root[0] of every polynomial of coeffs [n, 4] (what the nonlinear path consumes).  newton=True: the
    PF_F_NEWTON variant (Newton iteration where a, b >= 0, c > 0, d < 0; the closed form elsewhere)."""
    torch = nat.require_cuda()
    co = torch.as_tensor(np.ascontiguousarray(coeffs, dtype=np.float64), device="cuda").reshape(-1, 4)
    out = torch.zeros(co.shape[0], dtype=torch.float64, device="cuda")
    fn = nat.lib().pf_cubic_root0_newton if newton else nat.lib().pf_cubic_root0
    nat.check(fn(co.data_ptr(), out.data_ptr(), co.shape[0], nat.current_stream_ptr()), "pf_cubic_root0")
    return out.cpu().numpy()


def solve(a, b=None, c=None, d=None):
    """``solve(a, b, c, d)`` like the reference; ``solve(CS)`` with CS = CubicSolver(...) also works."""
    if b is None:
        a, b, c, d = a
    roots, nr = solve_many([[a, b, c, d]])
    r = roots[0, : nr[0]]
    return np.real(r) if np.all(np.imag(r) == 0) else r


def findF(a, b, c):
    return ((3.0 * c / a) - ((b ** 2.0) / (a ** 2.0))) / 3.0


def findG(a, b, c, d):
    return (((2.0 * (b ** 3.0)) / (a ** 3.0)) - ((9.0 * b * c) / (a ** 2.0)) + (27.0 * d / a)) / 27.0


def findH(g, f):
    return ((g ** 2.0) / 4.0 + (f ** 3.0) / 27.0)
