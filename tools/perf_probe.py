"""Development probe: kernel times and test counters of the render kernels on one GPU."""
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import _lib, scene as sc


def run(dev, name, w, h, depth, accel, reps=3, **kw):
    best = None
    for _ in range(reps):
        dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel=accel, **kw)
        dev.wait(0)
        st = dev.stats(0)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel=accel, flags=_lib.FLAG_COUNT_TESTS, **kw)
    dev.wait(0)
    c = dev.stats(0)
    px = best["pixels"]
    print("%-28s %-6s %5dx%-5d d%d  kernel %9.3f ms  %8.1f Mrays/s  rays/px %.2f  filt/ray %.1f box/ray %.1f exact/ray %.2f"
          % (name, accel, w, h, depth, best["kernel_ms"], best["rays"] / best["kernel_ms"] / 1e3,
             best["rays"] / max(px, 1), c["sphere_filter_tests"] / max(c["rays"], 1),
             c["box_tests"] / max(c["rays"], 1), c["exact_sphere_tests"] / max(c["rays"], 1)), flush=True)
    return best, c


def main():
    which = sys.argv[1:] or ["demo", "c3", "c4"]
    peak = _lib.fp32_peak(0)
    print("fp32 peak %.3e lane-instr/s" % peak)
    if "demo" in which:
        dev = sc.flatten(sc.demo_scene()).upload(0)
        for accel in ("exact", "linear", "bvh"):
            run(dev, "demo", 1920, 1080, 1, accel)
            run(dev, "demo", 1920, 1080, 5, accel)
        run(dev, "demo 8K", 7680, 4320, 1, "exact")
        dev.close()
    if "c3" in which:
        t0 = time.time()
        dev = sc.synthetic_scene("c3").upload(0)
        print("c3 upload %.2fs" % (time.time() - t0))
        run(dev, "c3", 3840, 2160, 5, "bvh")
        run(dev, "c3", 3840, 2160, 1, "bvh")
        b, c = run(dev, "c3 1/16 rows", 3840, 2160, 5, "linear", reps=2, band_rows=1, n_parts=16, part=3)
        print("   linear roofline: %.3e lane-instr/s = %.1f%% of measured FFMA peak"
              % (c["sphere_filter_tests"] * 10 / (b["kernel_ms"] * 1e-3),
                 100 * c["sphere_filter_tests"] * 10 / (b["kernel_ms"] * 1e-3) / peak))
        dev.close()
    if "c4" in which:
        t0 = time.time()
        dev = sc.synthetic_scene("c4").upload(0)
        print("c4 upload %.2fs" % (time.time() - t0))
        run(dev, "c4", 3840, 2160, 5, "bvh")
        run(dev, "c4", 3840, 2160, 1, "bvh")
        b, c = run(dev, "c4 1/540 rows", 3840, 2160, 5, "linear", reps=1, band_rows=1, n_parts=540, part=300)
        print("   linear roofline: %.3e lane-instr/s = %.1f%% of measured FFMA peak"
              % (c["sphere_filter_tests"] * 10 / (b["kernel_ms"] * 1e-3),
                 100 * c["sphere_filter_tests"] * 10 / (b["kernel_ms"] * 1e-3) / peak))
        dev.close()


if __name__ == "__main__":
    main()
