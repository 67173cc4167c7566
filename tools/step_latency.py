#!/usr/bin/env python
"""Per-step latency of the tile kernel on ONE default-geometry grid (15 CTAs: the latency-bound regime of a
single Controller() run).  Fits launch time = a + b * steps.  Run on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import pyfdtd_b200  # noqa: F401,E402
from pyfdtd_b200 import BaseFDTD11, Solver_Engine as SE, _device as dev, _native as nat  # noqa: E402
from test_host_layer import build_objects  # noqa: E402

lib = nat.lib()
for mode in ("free", "lorentz"):
    spec = dict(mode=mode, freq=9e9, dom=0.7, win=[7000, 8000], source="sine", periods=1000, epsRe=4.0 if mode == "free" else 1.0)
    V, P, C_V, C_P = build_objects(spec)
    C_V, Exs, Hys = SE.prepare_pass(V, P, C_V, C_P, lorentz=(mode == "lorentz"))
    arrs = BaseFDTD11._host_arrays(V, C_V, V.tempVarPol)
    rng = np.random.default_rng(0)
    for random_state in (False, True):
        if random_state:
            for k, sc in (("Ex", 1.0), ("Hy", 1 / 377.0), ("Dx", 8.85e-12), ("P", 8.85e-12), ("Pprev", 8.85e-12)):
                arrs[k] = rng.uniform(-1, 1, len(arrs[k])) * sc
        canon = dev.canonical_form(P, arrs)
        scal = BaseFDTD11.grid_scalars(V, P)
        scal.update(cE0=canon[0], cE1=canon[1], cH0=canon[2], cH1=canon[3], c2_pml=canon[4])
        flags = BaseFDTD11.grid_flags(P, False) | nat.PF_F_CANONICAL
        g = dev.DeviceGrid(L=len(V.Ex), T=int(P.timeSteps), arrays=arrs, scalars=scal, srcE=np.asarray(Exs) / P.courantNo,
                           srcH=np.asarray(Hys) / P.courantNo, probe_idx=[int(P.x1Loc)], flags=flags)
        sbytes = lib.pf_run_scratch_bytes(g.ref(), 1, nat.PF_ENGINE_TILE)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device="cuda")
        stream = nat.current_stream_ptr()
        res = []
        for ns in (8, 16, 32, 64):
            def go(reps):
                for r in range(reps):
                    nat.check(lib.pf_run_pass(g.ref(), dev.MODE_ID[mode], 1, 0, ns, nat.PF_ENGINE_TILE, None, 0, 0,
                                              scratch.data_ptr(), sbytes, stream), "run")
            go(10)
            torch.cuda.synchronize()
            lib.pf_profile_enable(1)
            go(50)
            ms, n = nat.c_double(), nat.c_int()
            lib.pf_profile_collect(ms, n)
            lib.pf_profile_enable(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(50); e1.record(); torch.cuda.synchronize()
            res.append((ns, 1e3 * ms.value / n.value, 1e3 * e0.elapsed_time(e1) / 50, n.value))
        (b, a) = np.polyfit([r[0] for r in res], [r[1] for r in res], 1)
        print(mode, "random" if random_state else "zero  ", " ".join("k=%d: %.2f us kernel / %.2f us call (%d launches)" % r for r in res))
        print("      fit: %.2f us per launch + %.3f us per step (%.0f cycles at 1.9 GHz)" % (a, b, b * 1900))
