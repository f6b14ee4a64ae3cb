"""Small frames through every wavefront kernel (cell grid, fused shadow + fold, BVH path, binned hits, scan) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import _lib, scene as sc
flat = sc.synthetic_scene("c3", n_spheres=3000)
dev = flat.upload(0)
ref = None
for accel, flags in (("grid", 0), ("bvh", 0), ("bvh", _lib.FLAG_NO_LIGHT_GRID), ("bvh", _lib.FLAG_NO_LIGHT_GRID | _lib.FLAG_WF_UNSORTED), ("linear", 0), ("bvh_mega", 0)):
    frame, st = dev.render(128, 72, 4, fmt="f64", accel=accel, flags=flags)
    if ref is None:
        ref = frame
    print(accel, flags, st["rays"], st["gpu_launches"], bool(np.array_equal(frame, ref)), flush=True)
rng = np.random.default_rng(1)
o = np.stack([rng.uniform(-45, 45, 4000), rng.uniform(-35, 5, 4000), rng.uniform(-5, 90, 4000)], axis=1)
d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([o, d], axis=1)
a = dev.trace_rays(rays, accel="exact"); b = dev.trace_rays(rays, accel="warp"); c = dev.trace_rays(rays, accel="grid")
print("rays", bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])))
dev.close()
