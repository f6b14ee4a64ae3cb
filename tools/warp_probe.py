"""Ray batches through the thread-per-ray scans and the warp-per-ray form (ERT_ACCEL_WARP): device time of the kernel.

usage: python tools/warp_probe.py [n_rays]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import scene as sc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(3)
for n_sph in (17, 64, 192, 1000):
    flat = sc.synthetic_scene("c3", n_spheres=n_sph)
    dev = flat.upload(0)
    # coherent rays (a 2000 x 1000 camera fan) and scattered rays (random origins and directions)
    xs, ys = np.meshgrid(np.linspace(-2, 2, 2000), np.linspace(-1.1, 1.1, n // 2000))
    d = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, 2.0)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    fan = np.concatenate([np.tile([0.0, 0.0, -2.0], (len(d), 1)), d], axis=1)
    o = np.stack([rng.uniform(-45, 45, n), rng.uniform(-35, 5, n), rng.uniform(-5, 90, n)], axis=1)
    d2 = rng.normal(size=(n, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    scat = np.concatenate([o, d2], axis=1)
    for name, rays in (("camera fan", fan), ("scattered", scat)):
        out = {}
        for accel in ("exact", "linear", "warp"):
            best = 1e9
            for _ in range(3):
                order, t = dev.trace_rays(rays, accel=accel)
                best = min(best, dev.stats(0)["kernel_ms"])
            out[accel] = (best, order, t)
        same = np.array_equal(out["exact"][1], out["warp"][1]) and np.array_equal(out["exact"][2], out["warp"][2])
        print("%5d spheres, %-10s %8d rays: exact %8.3f ms  linear %8.3f ms  warp %8.3f ms  (warp == exact: %s, hit rate %.2f)"
              % (n_sph, name, len(rays), out["exact"][0], out["linear"][0], out["warp"][0], same, (out["warp"][1] >= 0).mean()), flush=True)
    dev.close()
