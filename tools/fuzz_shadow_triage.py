"""Randomised check of the FP32 shadow triage (shadow_blocked, ert_wavefront.cuh) against the literal shadow path.

Every scene is rendered twice: by the default path (direction grids, triage, packed literal path, fused fold) and
with ERT_FLAG_NO_LIGHT_GRID (every shadow ray on the literal path through the BVH).  A shadow ray the triage wrongly
calls blocked changes a pixel.  usage: python tools/fuzz_shadow_triage.py [n_scenes] [seed]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import _lib, scene as sc

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = 0
shadow_rays = 0
t0 = time.time()
for k in range(n_scenes):
    n = int(rng.integers(300, 6000))
    scale = float(10.0 ** rng.uniform(-1, 2))                      # overall size of the scene: 0.1 ... 100
    box = np.array([rng.uniform(10, 40), rng.uniform(3, 20), rng.uniform(10, 40)]) * scale
    offset = (rng.uniform(-1, 1, 3) * (10.0 ** rng.uniform(0, 4)) if rng.random() < 0.3 else np.array([0.0, -box[1], box[2] + 5 * scale]))
    c = rng.uniform(-1, 1, (n, 3)) * box + offset
    r = (10.0 ** rng.uniform(-1.7, 0.5, n)) * scale * 0.5
    if rng.random() < 0.6:
        c = c.astype(np.float32).astype(np.float64)
        r = r.astype(np.float32).astype(np.float64)
    flat = sc.synthetic_scene("c3", n_spheres=n, seed=seed * 100003 + k)
    flat.spheres['center'] = c
    flat.spheres['radius'] = r
    n_l = int(rng.integers(1, 5))
    lights = np.zeros(n_l, dtype=_lib.LIGHT_DT)
    for l in range(n_l):
        kind = rng.random()
        if kind < 0.35:                                            # a hair's breadth to a few radii from a sphere
            s = int(rng.integers(0, n))
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            loc = c[s] + u * r[s] * (1.0 + 10.0 ** rng.uniform(-4, 0.5))
        elif kind < 0.7:                                           # somewhere in the cloud
            loc = offset + rng.uniform(-1, 1, 3) * box
        else:                                                      # far outside
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            loc = offset + u * np.linalg.norm(box) * 10.0 ** rng.uniform(0.3, 2.5)
        lights[l]['location'] = loc
        lights[l]['diffuse_colour'] = rng.uniform(0.2, 1, 3)
        lights[l]['specular_colour'] = (1, 1, 1)
    flat.lights = lights
    if rng.random() < 0.5:                                         # the floor plane under the cloud, sometimes skew
        flat.planes['normal'] = [(rng.uniform(-0.3, 0.3), -1.0 * rng.uniform(0.5, 2.0), rng.uniform(-0.3, 0.3))]
        flat.planes['distance'] = float(offset[1] + box[1] + scale)
    i = 0
    for tab in (flat.lights, flat.spheres, flat.triangles, flat.planes):
        tab['order'] = np.arange(i, i + len(tab), dtype=np.int32); i += len(tab)
    cam = sc.pose_camera(0)
    cam.location[:] = offset + np.array([0.0, 0.0, -1.6 * box[2]])
    cam.screen_width, cam.screen_height = 4.0, 2.25
    dev = flat.upload(0)
    a, sa = dev.render(192, 108, 3, fmt="f64", accel="auto", camera=cam)
    b, sb = dev.render(192, 108, 3, fmt="f64", accel="auto", camera=cam, flags=_lib.FLAG_NO_LIGHT_GRID)
    same = np.array_equal(a, b) and sa["rays"] == sb["rays"]
    shadow_rays += sum(sa["bounce_hits"]) * n_l
    if not same:
        bad += 1
        print("MISMATCH scene %d: n %d scale %.3g lights %d, %d pixels differ, max |d| %.3g" % (k, n, scale, n_l, int((a != b).any(axis=2).sum()), float(np.abs(a - b).max())), flush=True)
    dev.close()
print("%d scenes, %.1f M shadow rays, %d mismatches, %.0f s" % (n_scenes, shadow_rays / 1e6, bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
