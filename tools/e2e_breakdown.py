#!/usr/bin/env python
"""Where the end-to-end step of the batched sweep spends its time (run on the GPU box)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

wl = bench.ProductWorkload(1024, 512, 64)
b = wl.batch
b.randomize_state(seed=1)
b.upload()


def timed(fn, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for _ in range(3):
    b.reset_state(template=True); b.run(do_pol=True)
print("upload          %.3f ms  (%.1f MB)" % (timed(b.upload), b.h2d_bytes / 1e6))
print("reset_state     %.3f ms" % timed(lambda: b.reset_state(template=True)))
print("run             %.3f ms" % timed(lambda: b.run(do_pol=True)))
print("download_probes %.3f ms  (%.1f MB)" % (timed(b.download_probes), b.d2h_bytes / 1e6))
