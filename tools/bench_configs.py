#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4, 5 at N GPUs (config 2 is bench.py's own `value`).  Run directly (N = 1) or under
torchrun; weak scaling: per-GPU work is fixed.  Timing: barrier + synchronize on both sides, CUDA events, max over
ranks; rank 0 prints one JSON line per config.  Sweeps (3) shard members with no collective; long grids (1, 5) are
decomposed along z with a ghost exchange every k steps; PIC (4) runs one independent beam per GPU (replicas only).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import pyfdtd_b200  # noqa: F401,E402
from pyfdtd_b200 import Solver_Engine as SE, longgrid, pic, sweep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,3,4,5")
ap.add_argument("--cells-per-gpu", type=int, default=100_000_000)
ap.add_argument("--particles-per-gpu", type=int, default=20_000_000)
ap.add_argument("--members-per-gpu", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps):
    """seconds per call, max over ranks"""
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(**kw):
    if rank == 0:
        print(json.dumps(dict(kw, n_gpus=world, scaling="weak", dtype="f64", data="synthetic")), flush=True)


def long_grid(config, mode, alg_bytes, label):
    cells, steps = a.cells_per_gpu * world, 128
    grid, info = longgrid.lorentz_long_grid(cells, T=steps * (a.reps + 2), k=64, mode=mode, rank=rank, world_size=world)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    for arrs0, arrs1 in zip(*grid.bufs):
        for n, t in arrs0.items():
            if t is not None:
                t.copy_((torch.rand(t.shape, dtype=t.dtype, device=t.device, generator=gen) * 2 - 1) * sweep.MemberBatch.STATE_SCALE[n])
                arrs1[n].copy_(t)
    sec = timed(lambda: grid.run(steps, do_pol=(mode != "free")), a.reps)
    rate = cells * steps / sec / 1e9
    emit(config=config, workload=label, metric="Gcell-updates/s", value=rate, cells=cells, steps=steps, temporal_block_k=64,
         ranks_exchange="ghost cells every 64 steps (NCCL p2p)" if world > 1 else "none",
         algorithmic_GBps_k1=rate * alg_bytes, frac_of_hbm_peak_k1=rate * alg_bytes / (peak * world))
    del grid
    torch.cuda.empty_cache()


for cfg in [int(c) for c in a.configs.split(",")]:
    if cfg == 1:
        long_grid(1, "free", 32.0, "1D vacuum/dielectric Yee grid + CPML, fp64, streaming with 64-step temporal blocking")
    elif cfg == 5:
        long_grid(5, "lorentz", 32.0 + 0.7 * 40.0, "long Lorentz-dispersive grid (slab = right 70 %), z-decomposed")
        long_grid(5, "lorentz_nl", 32.0 + 0.7 * 40.0, "long dispersive + nonlinear grid (PF_LORENTZ_NL: Lorentz ADE + cubic Kerr law), z-decomposed")
    elif cfg == 3:
        for label, cubic in (("closed-form root (reference algorithm)", "closed"), ("Newton root (PF_F_NEWTON)", "newton")):
            SE.CUBIC = cubic
            try:
                batch, members = bench.build_nl_batch(a.members_per_gpu, 128)
                batch.randomize_state(seed=99 + rank)

                def step():
                    batch.reset_state(template=True)
                    batch.run(do_pol=False)
                sec = timed(step, a.reps)
                slab = sum(m.scalars["mr"] - m.scalars["mf"] for m in members)
                emit(config=3, workload="nonlinear cubic sweep, %d members/GPU x 128 steps, %s" % (a.members_per_gpu, label),
                     metric="Gcell-updates/s", value=batch.cell_steps * world / sec / 1e9,
                     cubic_solves_per_s=slab * 128 * world / sec, members=a.members_per_gpu * world)
                del batch, members
            finally:
                SE.CUBIC = "closed"
            torch.cuda.empty_cache()
    elif cfg == 4:
        L, dz, dt = 13194, 8.3276e-5, 2.6389e-13
        n = a.particles_per_gpu
        z, ux, uz, w = pic.make_beam(n, L, dz, seed=1 + rank)
        ps = pic.ParticleSet(z, ux, uz, w, L, dz, dt)
        Ex = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 1e5
        Hy = (torch.rand(L, dtype=torch.float64, device="cuda") * 2 - 1) * 3e2
        ps.sort()
        sec = timed(lambda: ps.step_sorted(Ex, Hy), 5)
        emit(config=4, workload="electron-beam PIC, %d macro-particles/GPU on a 13194-cell grid: push + cell re-sort + deposit (pf_pic_step_sorted), one beam per GPU (replicas)" % n,
             metric="particle-steps/s", value=n * world / sec, algorithmic_GBps=60.0 * n * world / sec / 1e9,
             frac_of_hbm_peak=60.0 * n / sec / 1e9 / peak)
        del ps
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
