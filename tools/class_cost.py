#!/usr/bin/env python
"""What a cell of each warp class costs in the headline kernel: the bench's Lorentz sweep batch with its geometry edited
(slab removed / CPML removed / slab everywhere), same arithmetic, same kernel.  usage: python tools/class_cost.py [members] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from pyfdtd_b200 import sweep, sweep_setup, _native as nat  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
f, a = bench.member_specs(M, 64)


def run(name, edit):
    t = sweep_setup.lorentz_sweep_tables(f, a, bench.DOM, *bench.WIN, periods=1000, nsteps=S)[1]
    edit(t)
    b = sweep.MemberBatch.from_table(t, "lorentz", T_alloc=S)
    b.upload()
    b.randomize_state(seed=1234)

    def step():
        b.reset_state(template=True)
        b.run(do_pol=True)
    sec = bench.time_cuda(torch, step, 2)
    L = t.L.astype(np.int64)
    cells = int(L.sum())
    slab = int(np.maximum(0, np.minimum(t.mr, L) - t.mf).sum())
    print(f"{name:34s} {b.cell_steps / sec / 1e9:8.1f} Gcell-updates/s   {sec / S * 1e6:8.2f} us/step   slab {slab / cells:5.1%}  pw {int(t.pw[0])} of {int(L[0])}")
    return sec / S


def no_slab(t):
    t.mf[:] = t.mr            # empty slab: vacuum + CPML both sides


def no_slab_no_cpml(t):
    t.mf[:] = t.mr
    t.flags[:] &= ~(nat.PF_F_CPML_M | nat.PF_F_CPML_P)


def no_cpml(t):
    t.flags[:] &= ~(nat.PF_F_CPML_M | nat.PF_F_CPML_P)


def all_slab(t):
    t.mf[:] = t.nzsrc + 2
    t.flags[:] &= ~(nat.PF_F_CPML_M | nat.PF_F_CPML_P)


run("standard geometry", lambda t: None)
run("no slab (vacuum + 2 CPML)", no_slab)
run("no slab, no CPML (all vacuum)", no_slab_no_cpml)
run("slab as is, no CPML", no_cpml)
run("slab from the source on, no CPML", all_slab)
