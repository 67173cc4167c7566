#!/usr/bin/env python
"""Regression pins for the BUILDER-DEFINED models -- the Kerr-Lorentz composition (mode "lorentz_nl"), the Drude limit
of the Lorentz ADE and the PIC step -- which the reference cannot pin because it does not contain them.
The vectors are outputs of THIS oracle (oracle/fdtd_oracle.*, oracle/pic_oracle.py) at the commit that introduced them:
they detect drift of the definitions, they are not reference parity.  TEST INFRASTRUCTURE.
Usage: python oracle/make_builder_golden.py   (writes tests/golden/builder_*.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fdtd_oracle as fo  # noqa: E402
import pic_oracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def kerr_lorentz():
    c = fo.make_case("lorentz_nl", 9e9, 0.15, 300, 320, source="gauss", amplitude=4.0)
    out = fo.run_case(c)
    np.savez_compressed(os.path.join(OUT, "builder_kerr_lorentz.npz"), Ex=out["Ex"], Hy=out["Hy"], Dx=out["Dx"], P=out["P"],
                        Acubic=out["Acubic"], x1ColBe=out["x1ColBe"], x1ColAf=out["x1ColAf"])


def drude():
    c = fo.make_case("lorentz", 9e9, 0.15, 300, 320, source="sine", periods=1000.0)
    c.medium = dict(c.medium, w0=0.0, wp=2 * np.pi * 12e9, gam=2 * np.pi * 0.2e9)
    out = fo.run_case(c)
    np.savez_compressed(os.path.join(OUT, "builder_drude.npz"), Ex=out["Ex"], Hy=out["Hy"], P=out["P"], x1ColAf=out["x1ColAf"],
                        plasmaFreqE=np.float64(out["plasmaFreqE"]))


def pic():
    L, dz, dt, n = 257, 8.3e-5, 2.6e-13, 20_000
    z, ux, uz, w, cell = po.make_beam(n, L, dz, seed=3, thermal=0.3)
    rng = np.random.default_rng(0)
    Ex, Hy = 2e5 * rng.standard_normal(L), 5e2 * rng.standard_normal(L)
    kw = dict(dz=dz, dt=dt, q_over_m=-1.75882001076e11, c=299792458.0, mu0=1.25663706127e-06)
    zo, uxo, uzo, wo, co = po.sort_by_cell(z, ux, uz, w, cell)
    for _ in range(3):
        zp, uxp, uzp, cp = po.push(zo, uxo, uzo, Ex, Hy, **kw)
        Jf = po.deposit_fused(zp, uxp, uzp, wo, co, cp, L, po.sub_warps(n, L), dz=dz, c=299792458.0, jx_scale=-1.602176634e-19)
        zo, uxo, uzo, wo, co = po.sort_by_cell(zp, uxp, uzp, wo, cp)
    J = po.deposit(zo, uxo, uzo, wo, co, L, dz=dz, c=299792458.0, jx_scale=-1.602176634e-19)
    np.savez_compressed(os.path.join(OUT, "builder_pic.npz"), z=zo, ux=uxo, uz=uzo, cell=co, J=J, J_fused=Jf)


if __name__ == "__main__":
    kerr_lorentz()
    drude()
    pic()
    print("written:", [f for f in sorted(os.listdir(OUT)) if f.startswith("builder_")])
