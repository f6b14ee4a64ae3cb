"""Fixed cost of a frame: kernel time and wall time of very small parts of the C4 frame (launch overhead, tails)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import _lib, scene as sc
dev = sc.synthetic_scene("c4").upload(0)
w, h, depth = 3840, 2160, 5
for n_parts in (2160, 540, 135, 34, 8):
    ms, wall = [], []
    for i in range(12):
        t0 = time.perf_counter()
        dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=1, n_parts=n_parts, part=n_parts // 2)
        dev.wait(0)
        t1 = time.perf_counter()
        if i >= 2:
            ms.append(dev.stats(0)["kernel_ms"]); wall.append(1e3 * (t1 - t0))
    st = dev.stats(0)
    print("1 row of %4d (%7d rays): kernel %.3f ms (min %.3f), wall %.3f ms, launches %d" % (n_parts, st["rays"], np.median(ms), min(ms), np.median(wall), st["gpu_launches"]), flush=True)
dev.close()
