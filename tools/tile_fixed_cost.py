#!/usr/bin/env python
"""Fixed cost of a k_tile launch on the headline workload: kernel time (CUDA events around the launch) against the number
of steps of the launch at a FIXED tile geometry (halo 64): time = P + s * steps.  Run on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pyfdtd_b200 import _native as nat  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = nat.lib()
pts = []
for S in (1, 4, 8, 16, 32, 48, 64):
    b, table = bench.lorentz_sweep_batch(M, S, 64)
    b.upload()
    b.randomize_state(seed=1234)

    def step():
        b.reset_state(template=True)
        b.run(do_pol=True, k_block=64)
    sec, kern = bench.timed_with_kernels(torch, nat, step, 3, warm=2)
    name, (n, ms) = max(kern.items(), key=lambda kv: kv[1][1])
    pts.append((S, ms / n))
    print(f"steps {S:3d}: {name} {ms / n * 1e3:9.1f} us per launch ({n} launches)", flush=True)
    del b
    torch.cuda.empty_cache()
x, y = np.array([p[0] for p in pts], float), np.array([p[1] for p in pts]) * 1e3
s, P = np.polyfit(x[2:], y[2:], 1)
print(f"fit over steps >= 8: {P:.1f} us fixed + {s:.2f} us per step  ->  fixed cost = {P / s:.1f} steps' worth; at k = 64 the launch is "
      f"{P / (P + 64 * s):.1%} fixed cost")
