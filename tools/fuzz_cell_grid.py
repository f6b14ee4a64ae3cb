"""Offline fuzz of the cell-grid walk (CPU only): random scenes and ray families through the host
restatement of the device walk (tests/cellgrid/shim.cpp); every sphere the exact ray touches must be
listed in a cell entered no later than the contact.  Not part of the test suite (run it after
touching cell_grid.cpp or grid_start/grid_advance):

    python tools/fuzz_cell_grid.py [first_seed] [n_seeds]
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import test_cell_grid as T  # noqa: E402


def load():
    # the pytest fixture's body, without pytest
    import subprocess
    os.makedirs(T.BUILD, exist_ok=True)
    src = [os.path.join(T.HERE, "cellgrid", "shim.cpp"), os.path.join(ROOT, "eraytracer_b200", "csrc", "cell_grid.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", T.SO] + src)
    L = ctypes.CDLL(T.SO)
    vp = ctypes.c_void_p
    L.cg_build.restype = vp
    L.cg_build.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_double]
    L.cg_free.argtypes = [vp]
    L.cg_enabled.argtypes = [vp]
    L.cg_geometry.argtypes = [vp, vp, vp, vp, vp]
    for name in ("cg_n_cells", "cg_n_refs", "cg_n_big"):
        getattr(L, name).restype = ctypes.c_longlong
        getattr(L, name).argtypes = [vp]
    for name in ("cg_cells", "cg_ref_sph", "cg_big"):
        getattr(L, name).restype = vp
        getattr(L, name).argtypes = [vp]
    L.cg_walk.restype = ctypes.c_longlong
    L.cg_walk.argtypes = [vp, vp, vp, ctypes.c_float, vp, vp, ctypes.c_longlong]
    return L


def one_scene(L, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(300, 6000))
    box = rng.uniform(5, 80, 3)
    box[rng.integers(0, 3)] *= rng.choice([1, 0.05, 3])
    scale = 10.0 ** rng.uniform(-1, 2.5)
    offset = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(0, 4.5)
    rlo = rng.uniform(0.05, 0.5)
    rhi = rlo * rng.uniform(1.2, 6)
    c, r = T.rand_scene(rng, n, box * scale, rlo * scale, rhi * scale, offset=offset, f32=bool(rng.integers(0, 2)))
    g = T.Grid(L, c, r, float(rng.choice([0.2, 0.35, 0.5, 1.0, 3.0])))
    if not g.enabled:
        return 0
    lo = c.min(axis=0) - 2 * rhi * scale
    hi = c.max(axis=0) + 2 * rhi * scale
    rays = [T.rand_rays(rng, 150, lo, hi), T.rand_rays(rng, 60, lo - 3 * (hi - lo), hi + 3 * (hi - lo))]
    a = T.rand_rays(rng, 60, lo, hi)
    k = rng.integers(0, 3, len(a))
    tiny = rng.choice([0.0, 1e-25, -1e-12, 1e-7, -1e-5], len(a))
    a[:, 3:] = tiny[:, None]
    a[np.arange(len(a)), 3 + k] = rng.choice([-1.0, 1.0], len(a))
    rays.append(a)
    b = T.rand_rays(rng, 60, lo, hi)
    for j in range(len(b)):
        for ax in rng.choice(3, int(rng.integers(1, 4)), replace=False):
            kp = np.floor((b[j, ax] - g.lo[ax]) / g.cs)
            b[j, ax] = float(np.float32(g.lo[ax]) + np.float32(kp) * np.float32(g.cs))
    rays.append(b)
    grazing = []
    for s in rng.integers(0, n, 60).tolist():
        o = c[s] + rng.normal(size=3) * (hi - lo).max() * 0.2
        to_c = c[s] - o
        if np.linalg.norm(to_c) <= r[s] * 1.01:
            continue
        perp = np.cross(to_c, rng.normal(size=3))
        perp /= np.linalg.norm(perp)
        d = c[s] + perp * r[s] * (1 - 10.0 ** rng.uniform(-12, -1)) - o
        grazing.append(np.concatenate([o, d / np.linalg.norm(d)]))
    rays.append(np.array(grazing))
    touched = T.check_complete(g, np.concatenate(rays), min_touch=0)
    g.close()
    return touched


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    L = load()
    t0 = time.time()
    scenes = pairs = 0
    for seed in range(first, first + count):
        t = one_scene(L, seed)
        scenes += 1 if t else 0
        pairs += t
    print("seeds %d..%d: %d scenes with a grid, %d (ray, touched sphere) pairs checked, %.0f s"
          % (first, first + count - 1, scenes, pairs, time.time() - t0))


if __name__ == "__main__":
    main()
