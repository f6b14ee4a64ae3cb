"""World-size-2 test of the multi-GPU host logic on CPU (gloo): each rank renders its own
interleaved row bands and writes them into one shared host frame; rank 0 must end up with
the whole frame.  The per-rank renderer here is the CPU oracle standing in for a GPU (this
is test infrastructure — the product path has no CPU renderer); on the GPU box the same
logic runs with ert_render() filling the shared frame (bench.py --gpus N, tests -m gpu)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, w, h, depth, band_rows, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eraytracer_b200 import multigpu
    from eraytracer_b200 import scene as sc
    from helpers import oracle_frame
    flat = sc.flatten(sc.demo_scene())
    nbytes = w * h * 3 * 8
    shared = None
    if rank == 0:
        shared = multigpu.SharedFrame(name, nbytes, create=True)
    dist.barrier()
    if rank != 0:
        shared = multigpu.SharedFrame(name, nbytes, create=False)
    frame = shared.array(np.float64, (h, w, 3))
    rows = multigpu.part_rows(h, band_rows, world, rank)
    ys, xs = np.meshgrid(rows, np.arange(w), indexing="ij")
    rgb, rays, _ = oracle_frame(flat, w, h, depth, pixels=(xs.reshape(-1), ys.reshape(-1)), nthreads=2)
    frame[rows] = rgb.reshape(len(rows), w, 3)
    # max-over-ranks of a per-rank "time" and sum of rays, as bench.py reduces them
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    r = torch.tensor([float(rays)], dtype=torch.float64)
    dist.all_reduce(r, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        full, full_rays, _ = oracle_frame(flat, w, h, depth, nthreads=2)
        q.put((bool(np.array_equal(full, frame)), float(t.item()), int(r.item()) == full_rays))
    dist.barrier()
    del frame
    shared.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,band_rows", [(40, 30, 4), (33, 17, 8)])
def test_two_ranks_assemble_one_frame(w, h, band_rows):
    world = 2
    port = 29600 + (os.getpid() % 300) + band_rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "ert_test_frame_%d_%d" % (os.getpid(), band_rows)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, w, h, 3, band_rows, q))
             for r in range(world)]
    for p in procs:
        p.start()
    ok, tmax, rays_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and tmax == 2.0 and rays_ok
    assert not os.path.exists(os.path.join("/dev/shm", name))
