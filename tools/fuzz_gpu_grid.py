"""Randomised GPU check of the cell-grid walk (128-byte blocks, overflow groups, survivor masks) against the linear scan.

Random scenes (sizes 0.1-100, offsets up to 1e4, radii over two decades, knots of small spheres that make long cell
lists, a few big spheres) and ray batches (inside, outside, aimed at knots): nearest hit by accel=grid must equal
accel=linear on (list position, Distance bits).  usage: python tools/fuzz_gpu_grid.py [n_scenes] [seed]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import scene as sc

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = grids = 0
rays_total = 0
t0 = time.time()
for k in range(n_scenes):
    n = int(rng.integers(300, 20000))
    scale = float(10.0 ** rng.uniform(-1, 2))
    box = np.array([rng.uniform(5, 40), rng.uniform(2, 20), rng.uniform(5, 40)]) * scale
    offset = rng.uniform(-1, 1, 3) * (10.0 ** rng.uniform(0, 4)) if rng.random() < 0.4 else np.zeros(3)
    c = rng.uniform(-1, 1, (n, 3)) * box + offset
    r = (10.0 ** rng.uniform(-1.5, 0.3, n)) * scale * 0.5
    n_knots = int(rng.integers(0, 4))
    knots = []
    for _ in range(n_knots):                                       # 20-90 small spheres in a region of about one cell
        m = int(rng.integers(20, 90))
        at = offset + rng.uniform(-0.8, 0.8, 3) * box
        idx = rng.choice(n, m, replace=False)
        c[idx] = at + rng.uniform(-0.4, 0.4, (m, 3)) * scale
        r[idx] = rng.uniform(0.02, 0.08, m) * scale
        knots.append(at)
    if rng.random() < 0.3:                                         # a few spheres too large for the cells
        idx = rng.choice(n, int(rng.integers(1, 6)), replace=False)
        r[idx] = rng.uniform(8, 30, len(idx)) * scale
    if rng.random() < 0.6:
        c = c.astype(np.float32).astype(np.float64)
        r = r.astype(np.float32).astype(np.float64)
    flat = sc.synthetic_scene("c3", n_spheres=n, seed=seed * 7919 + k)
    flat.spheres['center'] = c
    flat.spheres['radius'] = r
    dev = flat.upload(0)
    m = 30000
    o = offset + rng.uniform(-1.3, 1.3, (m, 3)) * box
    d = rng.normal(size=(m, 3))
    for at in knots:                                               # a third of the rays aim at a knot
        sel = rng.random(m) < 0.33 / max(len(knots), 1)
        d[sel] = at + rng.normal(size=(int(sel.sum()), 3)) * 0.2 * scale - o[sel]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1)
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    frame, st = dev.render(32, 18, 1, fmt="rgb8", accel="auto")
    grids += int(st["has_cell_grid"])
    rays_total += m
    if not (np.array_equal(og, ol) and np.array_equal(tg, tl)):
        bad += 1
        w = np.flatnonzero((og != ol) | (tg != tl))
        print("MISMATCH scene %d: n %d scale %.3g knots %d grid %d: %d rays differ, first %s" % (k, n, scale, n_knots, st["has_cell_grid"], len(w), rays[w[0]].tolist()), flush=True)
    dev.close()
print("%d scenes (%d with a cell grid), %.1f M rays, %d mismatches, %.0f s" % (n_scenes, grids, rays_total / 1e6, bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
