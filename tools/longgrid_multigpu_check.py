#!/usr/bin/env python
"""Run under torchrun on N GPUs: the long Lorentz grid decomposed over the ranks (NCCL point-to-point
ghost exchange every k steps) must equal the single-GPU undecomposed run bit for bit.  Also reports the
weak/strong-scaling throughput.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/longgrid_multigpu_check.py [--cells 200000 --steps 256]"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyfdtd_b200  # noqa: E402,F401
from pyfdtd_b200 import longgrid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=400_000)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--mode", default="lorentz", choices=["free", "lorentz", "nl", "lorentz_nl"])
ap.add_argument("--no-check", action="store_true", help="throughput only (skip the gather + single-GPU comparison)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                            pg_options=longgrid.nccl_options_for_overlap() if os.environ.get("PF_NCCL_HIGH_PRIO", "0") != "0" else None)
POL = a.mode in ("lorentz", "lorentz_nl")
grid, info = longgrid.lorentz_long_grid(a.cells, T=a.steps, k=a.k, rank=rank, world_size=world, mode=a.mode)
grid.overlap = os.environ.get("PF_LONGGRID_OVERLAP", "0") != "0"   # A/B: ghost exchange serial with / overlapped by the inner tiles
if a.no_check:
    # throughput runs start from a synthetic non-zero state (as bench.py's extras do): from zero most cells are
    # quiescent, which costs the Lorentz arithmetic the same but lets the cubic law skip its root (|d| <= 1e-8)
    from pyfdtd_b200 import sweep
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    for which in (0, 1):
        for arrs in grid.bufs[which]:
            for n, t in arrs.items():
                if t is not None and which == 0:
                    t.copy_((torch.rand(t.shape, dtype=t.dtype, device=t.device, generator=gen) * 2 - 1) * sweep.MemberBatch.STATE_SCALE[n])
    for arrs0, arrs1 in zip(*grid.bufs):
        for n in arrs0:
            if arrs0[n] is not None:
                arrs1[n].copy_(arrs0[n])
grid.run(a.k, do_pol=POL)           # warm-up block
if os.environ.get("PF_LONGGRID_BREAKDOWN"):
    import ctypes
    from pyfdtd_b200 import _native as nat
    for blk in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if len(grid.pieces) > 1:
            grid.exchange()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        nat.check(nat.lib().pf_run_block(grid.grids[grid.cur], grid.grids[grid.cur ^ 1], len(grid.mine), grid.mode_id, 1,
                                         grid.n_done, a.k, a.k, 0, grid.scratch.data_ptr(), grid.scratch_bytes,
                                         nat.current_stream_ptr()), "pf_run_block")
        t2 = time.perf_counter()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        grid.cur ^= 1; grid.n_done += a.k
        print(f"rank {rank} block {blk}: exchange {1e3*(t1-t0):.2f} ms, launch call {1e3*(t2-t1):.2f} ms, kernel {1e3*(t3-t2):.2f} ms", flush=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
grid.run(a.steps - a.k, do_pol=POL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
ok = True
names = ("Ex", "Hy", "P") if POL else ("Ex", "Hy")
mine = {} if a.no_check else {n: grid.gather_owned(n) for n in names}
if world > 1 and not a.no_check:
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        full = {n: np.concatenate([g[n] for g in gathered]) for n in mine}
        ref, _ = longgrid.lorentz_long_grid(a.cells, T=a.steps, k=a.k, rank=0, world_size=1, mode=a.mode)
        ref.run(a.steps, do_pol=POL)
        for n in full:
            same = np.array_equal(full[n], ref.gather_owned(n))
            ok &= same
            print(f"{n}: decomposed over {world} GPUs == single GPU: {same}")
if rank == 0:
    print(f"mode={a.mode} cells={a.cells} steps={a.steps - a.k} world={world} time={dt*1e3:.2f} ms "
          f"rate={a.cells*(a.steps - a.k)/dt/1e9:.1f} Gcell-updates/s ok={ok}")
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
