#!/usr/bin/env python
"""PF_F_FP32 (optional single-precision mode of the tile engine): accuracy against the reference goldens
(max abs error / peak of the reference array; stated tolerance 1e-5) and speed on the bench workloads
next to the fp64 modes.  Run on the GPU box.  Usage: python tools/fp32_report.py [--no-speed] [--full]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pyfdtd_b200  # noqa: F401,E402
from pyfdtd_b200 import MasterController as MC, Solver_Engine as SE  # noqa: E402
from conftest import load_golden  # noqa: E402
from test_host_layer import build_objects  # noqa: E402
import bench  # noqa: E402


def peak_err(got, want):
    s = float(np.max(np.abs(want)))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want))) / s) if s else float(np.max(np.abs(got)))


def accuracy(names):
    out = {}
    for name in names:
        g = load_golden(name)
        SE.USE_FP32 = True
        try:
            V, P, C_V, C_P = build_objects(g["spec"])
            V, P, C_V, C_P, Exs, Hys = MC.Controller(V, P, C_V, C_P)
        finally:
            SE.USE_FP32 = False
        row = {"steps": int(P.timeSteps), "cells": int(P.Nz) + 1}
        pairs = [("Ex", V.Ex), ("Hy", V.Hy)]
        if g["spec"]["mode"] == "nl":
            pairs += [("Port1", V.Port1), ("Port2", V.Port2), ("Acubic", V.Acubic)]
        else:
            pairs += [("x1ColBe", V.x1ColBe), ("x1ColAf", V.x1ColAf)]
        for nm, got in pairs:
            row[nm] = peak_err(got, g[nm])
        out[name] = row
        print(name, json.dumps(row), flush=True)
    return out


def speed(only=None):
    out = {}
    for label, fp32, fma, cubic in (("fp64_exact", False, False, "closed"), ("fp64_fma", False, True, "closed"),
                                    ("fp64_newton", False, False, "newton"), ("fp32", True, False, "closed")):
        if only and label not in only:
            continue
        try:
            b, table = bench.lorentz_sweep_batch(1024, 512, 64, fma=fma, fp32=fp32)
            b.upload()
            b.randomize_state(seed=1234)

            def step():
                b.reset_state(template=True)
                b.run(do_pol=True)
            sec = bench._time_cuda(torch, step, 3)
            out["lorentz_sweep_" + label] = b.cell_steps / sec / 1e9
            del b
            torch.cuda.empty_cache()
            nb, table = bench.nl_sweep_batch(256, 128, fp32=fp32, newton=(cubic == "newton"))
            nb.upload()
            nb.randomize_state()

            def nstep():
                nb.reset_state(template=True)
                nb.run(do_pol=False)
            sec = bench._time_cuda(torch, nstep, 2)
            out["nl_sweep_" + label] = nb.cell_steps / sec / 1e9
            del nb
            torch.cuda.empty_cache()
        finally:
            pass
        print(label, {k: round(v, 1) for k, v in out.items() if k.endswith(label)}, flush=True)
    return out


if __name__ == "__main__":
    names = ["free_sine_eps4", "free_gauss_eps4", "free_gauss_notfsf", "lorentz_sine", "lorentz_gauss", "lorentz_sine_6g",
             "nl_sine", "nl_sine_amp"]
    if "--full" in sys.argv:
        names += ["free_default_full", "lorentz_default_full"]
    res = {}
    if "--no-acc" not in sys.argv:
        res["accuracy_vs_reference_golden_peak_relative"] = accuracy(names)
    if "--no-speed" not in sys.argv:
        only = [a.split("=")[1].split(",") for a in sys.argv if a.startswith("--only=")]
        res["Gcell_updates_per_s"] = speed(only[0] if only else None)
    print(json.dumps(res))
