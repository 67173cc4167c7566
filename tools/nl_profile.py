#!/usr/bin/env python
"""A few launches of the nonlinear (cubic) sweep batch (config 3), for ncu captures of k_tile<PF_NL,...>.
Usage: python tools/nl_profile.py [members] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
variant = sys.argv[3] if len(sys.argv) > 3 else "closed"      # closed | newton | fp32
batch, table = bench.nl_sweep_batch(M, S, newton=variant == "newton", fp32=variant == "fp32")
batch.upload()
batch.randomize_state()


def step():
    batch.reset_state(template=True)
    batch.run(do_pol=False)


sec = bench.time_cuda(torch, step, 2)
slab = int((table.mr - table.mf).sum())
print({"variant": variant, "members": M, "steps": S, "Gcell_updates_per_s": batch.cell_steps / sec / 1e9, "cubic_solves_per_s": slab * S / sec})
