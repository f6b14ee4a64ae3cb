"""Row-band partition of one frame over several GPUs / processes.

The reference's distributed driver cuts the pixel list into ~64 chunks and hands them to
the least-loaded node (distribute_work/7, raytracer.erl:139-149); results come back keyed by
X+Y*Width and are sorted (master/3, erl:151-161).  Here a frame is cut into bands of
`band_rows` rows dealt round-robin to the parts (band b -> part b % n_parts), which keeps
sky-heavy and floor-heavy rows spread over all GPUs.  Each part's rows are copied by its
GPU straight to their place in ONE host frame, so the keyed gather + keysort disappears:
the "gather" is address arithmetic.  No collective is on the data path.

With one process per GPU (torchrun) the host frame is a shared mapping in /dev/shm that
every rank page-locks (ert_host_register) and fills; rank 0 then owns the whole frame.
"""
import ctypes
import mmap
import os

import numpy as np


def part_rows(height, band_rows, n_parts, part):
    """Row indices (ascending) that belong to `part` — mirrors local_rows_of() in ert_api.cu."""
    if n_parts <= 1 or band_rows <= 0:
        return np.arange(height)
    rows = []
    n_bands = (height + band_rows - 1) // band_rows
    for b in range(part, n_bands, n_parts):
        rows.extend(range(b * band_rows, min((b + 1) * band_rows, height)))
    return np.asarray(rows, dtype=np.int64)


def default_band_rows(height, n_parts):
    """Bands of 8 rows (one thread-block tile) unless the image is too small to give every
    part a band."""
    if n_parts <= 1:
        return 0
    rows = 8
    while rows > 1 and (height + rows - 1) // rows < 4 * n_parts:
        rows //= 2
    return rows


class SharedFrame:
    """One frame in /dev/shm mapped by every rank; optionally page-locked for async D2H."""

    def __init__(self, name, nbytes, create):
        self.path = os.path.join("/dev/shm", name)
        self.nbytes = int(nbytes)
        self.created = create
        flags = os.O_RDWR | (os.O_CREAT if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self.map = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self._c = (ctypes.c_uint8 * self.nbytes).from_buffer(self.map)
        self.ptr = ctypes.addressof(self._c)
        self.registered = False

    def register(self):
        """Page-locks the mapping with the CUDA driver (needs a GPU)."""
        from . import _lib
        _lib.check(_lib.load().ert_host_register(self.ptr, self.nbytes))
        self.registered = True

    def array(self, dtype, shape):
        return np.frombuffer(self.map, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self, unlink=None):
        if self.registered:
            from . import _lib
            _lib.load().ert_host_unregister(self.ptr)
            self.registered = False
        self._c = None
        try:
            self.map.close()
        except BufferError:
            pass
        if unlink if unlink is not None else self.created:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass


def assemble_parts(height, width, band_rows, n_parts, part_frames, dtype):
    """Host-side reference of the placement rule: part frames (compact, own rows only, in
    ascending row order) -> full frame.  Used by the CPU tests of the partition logic."""
    out = np.zeros((height, width, 3), dtype=dtype)
    for part, fr in enumerate(part_frames):
        rows = part_rows(height, band_rows, n_parts, part)
        out[rows] = np.asarray(fr).reshape(len(rows), width, 3)
    return out
