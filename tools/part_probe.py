"""Times one row-band part of a frame on one GPU (what each GPU of an N-GPU run does).

usage: python tools/part_probe.py [workload] [n_parts] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from eraytracer_b200 import _lib, multigpu, scene as sc  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    n_parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    w, h, depth = 3840, 2160, 5
    flat = sc.synthetic_scene(wl)
    dev = flat.upload(0)
    band = multigpu.default_band_rows(h, n_parts)
    for flags, tag in ((0, "plain"), (_lib.FLAG_TIME_KERNELS, "timed")):
        ms, split = [], None
        for i in range(reps + 2):
            _lib.l2_flush(0)
            dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=band, n_parts=n_parts, part=0,
                             flags=flags)
            dev.wait(0)
            st = dev.stats(0)
            if i >= 2:
                ms.append(st["kernel_ms"])
                split = (st["path_ms"], st["shadow_ms"], st["other_ms"], st["gpu_launches"])
        print("%s part 0 of %d (%s): kernel_ms median %.3f min %.3f; path/shadow/other %s rays %d"
              % (wl, n_parts, tag, float(np.median(ms)), min(ms), split, st["rays"]))
    dev.close()


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("split", "flush")):
    main()


def split_probe():
    """Two sub-parts of one part on two slots (streams) at once, against the part on one slot: host wall time."""
    import time
    wl = sys.argv[2] if len(sys.argv) > 2 else "c4"
    n_parts = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    reps = 12
    w, h, depth = 3840, 2160, 5
    flat = sc.synthetic_scene(wl)
    dev = flat.upload(0)
    band = multigpu.default_band_rows(h, n_parts)

    def one():
        dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=band, n_parts=n_parts, part=0)
        dev.wait(0)

    def two():
        dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=band, n_parts=2 * n_parts, part=0)
        dev.render_async(w, h, depth, slot=1, fmt="rgb8", accel="auto", band_rows=band, n_parts=2 * n_parts, part=n_parts)
        dev.wait(0)
        dev.wait(1)

    for name, fn in (("one slot", one), ("two slots", two), ("one slot", one), ("two slots", two)):
        ts = []
        for i in range(reps + 3):
            _lib.l2_flush(0)
            t0 = time.perf_counter()
            fn()
            if i >= 3:
                ts.append(1e3 * (time.perf_counter() - t0))
        print("%s part 0 of %d, %s: wall ms median %.3f min %.3f" % (wl, n_parts, name, float(np.median(ts)), min(ts)))
    dev.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "split":
    split_probe()


def flush_probe():
    """One part of N with and without the L2 flush between frames."""
    wl = sys.argv[2] if len(sys.argv) > 2 else "c4"
    w, h, depth = 3840, 2160, 5
    flat = sc.synthetic_scene(wl)
    dev = flat.upload(0)
    for n_parts in (1, 8):
        band = multigpu.default_band_rows(h, n_parts)
        for flush in (True, False):
            ms = []
            for i in range(12):
                if flush:
                    _lib.l2_flush(0)
                dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=band, n_parts=n_parts, part=0)
                dev.wait(0)
                if i >= 2:
                    ms.append(dev.stats(0)["kernel_ms"])
            print("%s part 0 of %d, L2 %s: kernel_ms median %.3f min %.3f" % (wl, n_parts, "flushed" if flush else "as the last frame left it", float(np.median(ms)), min(ms)))
    dev.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "flush":
    flush_probe()
