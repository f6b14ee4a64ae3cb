"""Concurrent device -> host copy ceiling of the box: every rank copies its share of an 8K RGB8 frame (or a whole
frame) to pinned host memory at the same time.  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/d2h_probe.py

Modes: private = cudaMallocHost buffer per rank; shared = one /dev/shm frame page-locked by every rank (what
bench.py does), each rank writing its interleaved row bands (cudaMemcpy2DAsync) or one contiguous slice."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from eraytracer_b200 import multigpu


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lr))
    w, h = 7680, 4320
    frame_bytes = w * h * 3
    reps = 30

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def report(tag, nbytes, secs):
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            tot = nbytes * world * reps / t.item() / 1e9
            print("%-58s %7.1f GB/s aggregate, %6.1f GB/s per GPU (%d ranks)" % (tag, tot, tot / world, world), flush=True)

    dev = torch.empty(frame_bytes, dtype=torch.uint8, device="cuda")
    # (a) private pinned buffers: a whole frame per rank, and a 1/world share per rank
    for share in (1, world):
        n = frame_bytes // share
        host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            host.copy_(dev[:n], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        report("private pinned, %5.1f MB per rank per copy" % (n / 1e6), n, dt)
        del host
    # (b) one shared /dev/shm frame registered by every rank; contiguous slices, then interleaved 8-row bands
    name = "ert_d2h_probe_%s" % os.environ.get("MASTER_PORT", "0")
    sh = multigpu.SharedFrame(name, frame_bytes, create=True) if rank == 0 else None
    barrier()
    if rank != 0:
        sh = multigpu.SharedFrame(name, frame_bytes, create=False)
    sh.register()
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12") if False else None
    n = frame_bytes // world
    host_t = torch.frombuffer(sh.map, dtype=torch.uint8)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        host_t[rank * n:(rank + 1) * n].copy_(dev[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    report("shared /dev/shm frame, contiguous slice per rank", n, dt)
    rows = multigpu.part_rows(h, 8, world, rank) if world > 1 else np.arange(h)
    bands = len(rows) // 8
    src = dev[:bands * 8 * w * 3].view(bands, 8 * w * 3)
    dst = host_t.view(h // 8, 8 * w * 3)[rank::world][:bands]
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    report("shared /dev/shm frame, interleaved 8-row bands (2-D copy)", bands * 8 * w * 3, dt)
    barrier()
    del host_t, dst
    sh.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
