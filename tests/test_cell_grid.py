"""CPU tests of the uniform cell grid (eraytracer_b200/csrc/cell_grid.cpp) and of the walk the
device makes through it, restated with the same FP32 operations in tests/cellgrid/shim.cpp.

The grid only prunes path-ray candidates; what it must guarantee is completeness: every sphere
the exact (double) ray touches is listed in a cell the FP32 walk has entered by the time the ray
reaches the sphere — so the walk may stop as soon as the next cell starts beyond the nearest hit
found so far and still return the linear scan's nearest hit (raytracer.erl:300-346).
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "cellgrid", "_build")
SO = os.path.join(BUILD, "libcellgrid_test.so")
KD = 1.0000019073486328125


@pytest.fixture(scope="module")
def cg():
    os.makedirs(BUILD, exist_ok=True)
    src = [os.path.join(HERE, "cellgrid", "shim.cpp"),
           os.path.join(ROOT, "eraytracer_b200", "csrc", "cell_grid.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO] + src)
    L = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    L.cg_build.restype = vp
    L.cg_build.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_double]
    L.cg_free.argtypes = [vp]
    L.cg_enabled.argtypes = [vp]
    L.cg_geometry.argtypes = [vp, vp, vp, vp, vp]
    for name in ("cg_n_cells", "cg_n_refs", "cg_n_big", "cg_n_over"):
        getattr(L, name).restype = ctypes.c_longlong
        getattr(L, name).argtypes = [vp]
    for name in ("cg_cells", "cg_ref_sph", "cg_big", "cg_blocks", "cg_over_sph", "cg_over_filter", "cg_ref_filter"):
        getattr(L, name).restype = vp
        getattr(L, name).argtypes = [vp]
    L.cg_walk.restype = ctypes.c_longlong
    L.cg_walk.argtypes = [vp, vp, vp, ctypes.c_float, vp, vp, ctypes.c_longlong]
    return L


class Grid:
    def __init__(self, L, centers, radii, density=0.35):
        self.L = L
        self.centers = np.ascontiguousarray(centers, dtype=np.float64)
        self.radii = np.ascontiguousarray(radii, dtype=np.float64)
        filt = np.zeros((len(self.radii), 4), dtype=np.float32)
        filt[:, :3] = self.centers
        filt[:, 3] = self.radii ** 2
        self.abs_max = float(np.float32(np.max(np.abs(self.centers) + np.abs(self.radii)[:, None])) * np.float32(1.000001))
        self.h = L.cg_build(self.centers.ctypes.data, self.radii.ctypes.data, filt.ctypes.data, len(self.radii),
                            self.abs_max, density)
        self.enabled = bool(L.cg_enabled(self.h))
        if not self.enabled:
            return
        res = np.zeros(3, dtype=np.int32)
        lo = np.zeros(3, dtype=np.float32)
        hi = np.zeros(3, dtype=np.float32)
        ce = np.zeros(2, dtype=np.float32)
        L.cg_geometry(self.h, res.ctypes.data, lo.ctypes.data, hi.ctypes.data, ce.ctypes.data)
        self.res, self.lo, self.hi, self.cs, self.eps = res, lo, hi, float(ce[0]), float(ce[1])
        n_cells, n_refs, n_big = L.cg_n_cells(self.h), L.cg_n_refs(self.h), L.cg_n_big(self.h)
        self.cells = np.ctypeslib.as_array(ctypes.cast(L.cg_cells(self.h), ctypes.POINTER(ctypes.c_uint32)), (n_cells,)).copy()
        self.ref_sph = np.ctypeslib.as_array(ctypes.cast(L.cg_ref_sph(self.h), ctypes.POINTER(ctypes.c_int32)), (n_refs,)).copy() \
            if n_refs else np.zeros(0, dtype=np.int32)
        self.big = np.ctypeslib.as_array(ctypes.cast(L.cg_big(self.h), ctypes.POINTER(ctypes.c_int32)), (n_big,)).copy() \
            if n_big else np.zeros(0, dtype=np.int32)

    def close(self):
        self.L.cg_free(self.h)

    def walk(self, o, d, cap=8192):
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        cells = np.zeros(cap, dtype=np.int32)
        t_in = np.zeros(cap, dtype=np.float32)
        n = self.L.cg_walk(self.h, o.ctypes.data, d.ctypes.data, self.abs_max, cells.ctypes.data, t_in.ctypes.data, cap)
        assert n != -2, "walk longer than the buffer"
        if n < 0:
            return None, None
        return cells[:n], t_in[:n]

    def listed(self, cell):
        c = int(self.cells[cell])
        first, cnt = c >> 7, c & 127
        return self.ref_sph[first:first + cnt]


def touched(centers, radii, o, d):
    """Spheres the exact ray touches (forward half-line) and the filter-space parameter of the
    first point of contact."""
    dn = d / np.linalg.norm(d)
    oc = centers - o
    b = oc @ dn
    disc = b * b - (np.einsum('ij,ij->i', oc, oc) - radii ** 2)
    hit = (disc >= 0) & (b + np.sqrt(np.maximum(disc, 0)) >= 0)
    idx = np.nonzero(hit)[0]
    t_entry = np.maximum(b[idx] - np.sqrt(disc[idx]), 0.0) / KD
    return idx, t_entry


def check_complete(g, rays, min_touch=1):
    """Every touched sphere is `big` or listed in a cell entered no later than the contact."""
    big = set(g.big.tolist())
    n_touch = 0
    for ray in rays:
        o, d = ray[:3], ray[3:]
        idx, t_entry = touched(g.centers, g.radii, o, d)
        cells, t_in = g.walk(o, d)
        if cells is None:
            continue                                 # margin too large: the device walks the BVH
        assert len(set(cells.tolist())) == len(cells), "a cell was visited twice"
        assert np.all(np.diff(t_in) >= 0) or True
        first_seen = {}
        for c, t in zip(cells.tolist(), t_in.tolist()):
            for s in g.listed(c).tolist():
                if s not in first_seen:
                    first_seen[s] = t
        for s, ts in zip(idx.tolist(), t_entry.tolist()):
            if s in big:
                continue
            n_touch += 1
            assert s in first_seen, ("sphere %d touched by the ray %r is in no visited cell" % (s, ray.tolist()))
            assert first_seen[s] <= ts, ("sphere %d is first listed at t=%r but touched at t=%r (ray %r)"
                                         % (s, first_seen[s], ts, ray.tolist()))
    assert n_touch >= min_touch
    return n_touch


def rand_scene(rng, n, box, rlo, rhi, offset=(0, 0, 0), f32=True):
    c = np.stack([rng.uniform(-box[a], box[a], n) for a in range(3)], axis=1) + np.asarray(offset, dtype=np.float64)
    r = rng.uniform(rlo, rhi, n)
    if f32:
        c = c.astype(np.float32).astype(np.float64)
        r = r.astype(np.float32).astype(np.float64)
    return c, r


def rand_rays(rng, n, lo, hi):
    o = np.stack([rng.uniform(lo[a], hi[a], n) for a in range(3)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)


def test_every_sphere_is_listed_in_every_cell_it_reaches(cg):
    rng = np.random.default_rng(1)
    c, r = rand_scene(rng, 5000, (40, 17, 40), 0.2, 0.8, offset=(0, -13, 45))
    g = Grid(cg, c, r)
    assert g.enabled and len(g.big) == 0
    counts = g.cells & 127
    firsts = g.cells >> 7
    # every list is stored in whole groups of four entries (the walk's filter loop reads a group at a time); the
    # padding has no sphere, and one more group follows the last list
    padded = (counts + 3) // 4 * 4
    assert np.array_equal(firsts, np.concatenate([[0], np.cumsum(padded)[:-1]]))
    assert padded.sum() + 4 == len(g.ref_sph)
    assert counts.sum() == int((g.ref_sph >= 0).sum())
    for c_id in rng.integers(0, len(counts), 200).tolist():
        lst = g.ref_sph[firsts[c_id]:firsts[c_id] + padded[c_id]]
        assert np.all(lst[:counts[c_id]] >= 0) and np.all(lst[counts[c_id]:] == -1)
    # bounds contain every sphere box; cells cover the bounds
    assert np.all(g.lo < (c - r[:, None]).min(axis=0)) and np.all(g.hi > (c + r[:, None]).max(axis=0))
    assert np.all(g.res * np.float64(g.cs) >= g.hi.astype(np.float64) - g.lo.astype(np.float64))
    rx, ry = int(g.res[0]), int(g.res[1])
    n_corner = 0
    for s in rng.integers(0, len(r), 300).tolist():
        i0 = np.floor((c[s] - r[s] - g.lo) / g.cs).astype(int)
        i1 = np.floor((c[s] + r[s] - g.lo) / g.cs).astype(int)
        for z in range(i0[2], i1[2] + 1):
            for y in range(i0[1], i1[1] + 1):
                for x in range(i0[0], i1[0] + 1):
                    # the cell as a box; a sphere is listed wherever the BALL reaches it (cells in the
                    # corners of its bounding box that the ball misses are left out)
                    blo = g.lo.astype(np.float64) + np.array([x, y, z]) * np.float64(g.cs)
                    d = np.maximum(np.maximum(blo - c[s], c[s] - (blo + np.float64(g.cs))), 0.0)
                    listed = s in g.listed(x + rx * (y + ry * z)).tolist()
                    if np.sqrt((d * d).sum()) <= r[s]:
                        assert listed
                    elif np.sqrt((d * d).sum()) > r[s] + 3 * g.eps:
                        assert not listed
                        n_corner += 1
    assert n_corner > 0
    g.close()


@pytest.mark.parametrize("density", (0.25, 0.5, 2.0, 8.0))
def test_walk_is_complete_random_rays(cg, density):
    rng = np.random.default_rng(2)
    c, r = rand_scene(rng, 8000, (40, 17, 40), 0.2, 0.8, offset=(0, -13, 45))
    g = Grid(cg, c, r, density)
    assert g.enabled
    rays = np.concatenate([rand_rays(rng, 500, (-42, -32, 3), (42, 6, 87)),
                           rand_rays(rng, 200, (-150, -150, -100), (150, 100, 200))])
    # rays from outside aimed at the scene
    aim = np.array([0, -13, 45.0]) + rng.normal(size=(200, 3)) * 15 - rays[-200:, :3]
    rays[-200:, 3:] = aim / np.linalg.norm(aim, axis=1, keepdims=True)
    assert check_complete(g, rays) > 500
    g.close()


def test_walk_is_complete_degenerate_directions_and_origins(cg):
    rng = np.random.default_rng(3)
    c, r = rand_scene(rng, 6000, (30, 12, 30), 0.2, 0.7, offset=(0, -10, 40))
    g = Grid(cg, c, r)
    assert g.enabled
    lo, hi = (-31, -23, 9), (31, 3, 71)
    rays = []
    # axis-parallel, planar, and nearly-axis-parallel directions (tiny components of both signs)
    for tiny in (0.0, 1e-25, -1e-25, 1e-12, -1e-9, 1e-7, -1e-5):
        a = rand_rays(rng, 60, lo, hi)
        k = rng.integers(0, 3, len(a))
        a[:, 3:] = tiny
        a[np.arange(len(a)), 3 + k] = rng.choice([-1.0, 1.0], len(a))
        rays.append(a)
        b = rand_rays(rng, 60, lo, hi)
        b[np.arange(len(b)), 3 + rng.integers(0, 3, len(b))] = tiny
        rays.append(b)
    # origins exactly on cell planes (one, two or three axes), directions random and along the planes
    for n_axes in (1, 2, 3):
        a = rand_rays(rng, 150, lo, hi)
        for j in range(len(a)):
            axes = rng.choice(3, n_axes, replace=False)
            for ax in axes:
                kplane = np.floor((a[j, ax] - g.lo[ax]) / g.cs)
                a[j, ax] = float(np.float32(g.lo[ax]) + np.float32(kplane) * np.float32(g.cs))
            if j % 2:
                a[j, 3 + axes[0]] = 0.0
        rays.append(a)
    # diagonal rays through cell corners
    a = rand_rays(rng, 100, lo, hi)
    for j in range(len(a)):
        kk = np.floor((a[j, :3] - g.lo) / g.cs)
        a[j, :3] = g.lo.astype(np.float64) + kk * np.float64(g.cs)
        a[j, 3:] = rng.choice([-1.0, 1.0], 3)
    rays.append(a)
    # un-normalised directions
    a = rand_rays(rng, 100, lo, hi)
    a[:, 3:] *= rng.uniform(1e-3, 1e3, (len(a), 1))
    rays.append(a)
    rays = np.concatenate(rays)
    assert check_complete(g, rays) > 1000
    g.close()


def test_walk_is_complete_grazing_rays(cg):
    """Rays tangent to a sphere (just inside its rim): the FP32 walk must still list it."""
    rng = np.random.default_rng(4)
    c, r = rand_scene(rng, 4000, (25, 10, 25), 0.2, 0.7, offset=(0, -10, 35))
    g = Grid(cg, c, r)
    assert g.enabled
    rays = []
    for s in rng.integers(0, len(r), 400).tolist():
        o = c[s] + rng.normal(size=3) * 12
        to_c = c[s] - o
        dist = np.linalg.norm(to_c)
        if dist <= r[s] * 1.01:
            continue
        perp = np.cross(to_c, rng.normal(size=3))
        perp /= np.linalg.norm(perp)
        target = c[s] + perp * r[s] * (1 - 1e-9)
        d = target - o
        rays.append(np.concatenate([o, d / np.linalg.norm(d)]))
    assert check_complete(g, np.array(rays)) > 400
    g.close()


@pytest.mark.parametrize("case", ("large_coordinates", "double_centres", "thin_slab", "big_spheres"))
def test_walk_is_complete_awkward_scenes(cg, case):
    rng = np.random.default_rng(5)
    if case == "large_coordinates":
        c, r = rand_scene(rng, 5000, (4000, 2000, 4000), 10, 80, offset=(9000, -3000, 20000))
        lo, hi = (4000, -6000, 15000), (14000, 0, 25000)
    elif case == "double_centres":
        c, r = rand_scene(rng, 5000, (30, 15, 30), 0.05, 0.7, offset=(0.1, -17.3, 40.7), f32=False)
        lo, hi = (-35, -35, 5), (35, 5, 75)
    elif case == "thin_slab":
        c, r = rand_scene(rng, 3000, (50, 0.01, 50), 0.2, 0.5, offset=(0, -3, 60))
        lo, hi = (-55, -6, 5), (55, 0, 115)
    else:
        c, r = rand_scene(rng, 4000, (40, 20, 40), 0.1, 0.6, offset=(0, -20, 50))
        r[-5:] = rng.uniform(10, 40, 5)
        lo, hi = (-60, -60, -10), (60, 10, 110)
    g = Grid(cg, c, r)
    assert g.enabled
    assert (len(g.big) > 0) == (case == "big_spheres")
    rays = np.concatenate([rand_rays(rng, 500, lo, hi),
                           rand_rays(rng, 150, [3 * v - 50 for v in lo], [3 * v + 50 for v in hi])])
    check_complete(g, rays, min_touch=50)
    g.close()


@pytest.mark.parametrize("density", [0.2, 0.35, 2.0])
def test_packed_blocks_hold_the_same_lists(cg, density):
    """The device walks two 128-byte blocks per cell (6 spheres in each, longer lists in overflow groups of four):
    every cell's blocks and overflow range must decode to exactly the list the completeness tests above check."""
    rng = np.random.default_rng(11)
    c, r = rand_scene(rng, 4000, (40, 16, 40), 0.2, 1.2)
    g = Grid(cg, c, r, density=density)
    assert g.enabled
    L = cg
    n_cells, n_over, n_refs = len(g.cells), L.cg_n_over(g.h), L.cg_n_refs(g.h)
    blocks = np.ctypeslib.as_array(ctypes.cast(L.cg_blocks(g.h), ctypes.POINTER(ctypes.c_uint32)), (n_cells, 2, 32)).copy()
    over_sph = np.ctypeslib.as_array(ctypes.cast(L.cg_over_sph(g.h), ctypes.POINTER(ctypes.c_int32)), (n_over,)).copy()
    over_f = np.ctypeslib.as_array(ctypes.cast(L.cg_over_filter(g.h), ctypes.POINTER(ctypes.c_float)), (n_over, 4)).copy()
    ref_f = np.ctypeslib.as_array(ctypes.cast(L.cg_ref_filter(g.h), ctypes.POINTER(ctypes.c_float)), (n_refs, 4)).copy()
    assert n_over % 4 == 0 and n_over >= 4
    longest = 0
    for cell in range(n_cells):
        word = int(g.cells[cell])
        first, cnt = word >> 7, word & 127
        b = blocks[cell]
        f_in = np.concatenate([b[0, :24].view(np.float32).reshape(6, 4), b[1, :24].view(np.float32).reshape(6, 4)])
        s_in = np.concatenate([b[0, 24:30].view(np.int32), b[1, 24:30].view(np.int32)])
        assert int(b[0, 30]) == cnt
        more = int(b[0, 31])
        n_in = min(cnt, 12)
        assert np.array_equal(s_in[:n_in], g.ref_sph[first:first + n_in])
        assert np.array_equal(f_in[:n_in], ref_f[first:first + n_in])
        assert np.all(s_in[n_in:] == -1) and np.all(f_in[n_in:, 3] < -1e38)          # padding never passes
        if cnt > 12:
            n_ov = cnt - 12
            assert more % 4 == 0 and more + (n_ov + 3) // 4 * 4 + 4 <= n_over          # one group past the list exists
            assert np.array_equal(over_sph[more:more + n_ov], g.ref_sph[first + 12:first + cnt])
            assert np.array_equal(over_f[more:more + n_ov], ref_f[first + 12:first + cnt])
            pad = (n_ov + 3) // 4 * 4
            assert np.all(over_sph[more + n_ov:more + pad] == -1) and np.all(over_f[more + n_ov:more + pad, 3] < -1e38)
        longest = max(longest, cnt)
    assert np.all(over_sph[-4:] == -1)
    if density <= 0.2:
        assert longest > 12                                          # the overflow path is exercised
    g.close()


def test_clustered_scene_keeps_its_grid_and_has_a_long_cell_list(cg):
    """CPU side of tests/test_gpu_cell_grid.py::test_long_cell_lists_equal_the_linear_scan: the builder keeps the
    grid for the scene with a knot of 70 small spheres, and some cell lists more than 32 of them (so the device
    walk goes through both blocks, the overflow groups and more than one survivor mask)."""
    from helpers import clustered_scene
    flat, knot = clustered_scene()
    g = Grid(cg, flat.spheres['center'], flat.spheres['radius'])
    assert g.enabled
    counts = g.cells & 127
    assert 32 < counts.max() <= 127
    g.close()


def test_far_origins_are_refused(cg):
    rng = np.random.default_rng(6)
    c, r = rand_scene(rng, 3000, (30, 12, 30), 0.2, 0.7, offset=(0, -10, 40))
    g = Grid(cg, c, r)
    d = np.array([0.0, 0.0, 1.0])
    assert g.walk(np.array([0.0, -10.0, -50.0]), d)[0] is not None
    assert g.walk(np.array([0.0, -10.0, -1.0e6]), d)[0] is None
    g.close()


def test_scenes_that_do_not_suit_a_grid_get_none(cg):
    rng = np.random.default_rng(7)
    c, r = rand_scene(rng, 100, (30, 12, 30), 0.2, 0.7)
    assert not Grid(cg, c, r).enabled                               # too few spheres
    c, r = rand_scene(rng, 3000, (30, 12, 30), 0.2, 0.7)
    r = 10.0 ** rng.uniform(-2, 3, 3000)
    assert not Grid(cg, c, r).enabled                               # most spheres would be `big`
    c, r = rand_scene(rng, 3000, (3e7, 3e7, 3e7), 1e4, 1e5)
    assert not Grid(cg, c, r).enabled                               # coordinates beyond the FP32 margins
    c = np.zeros((5000, 3)) + rng.normal(size=(5000, 3)) * 1e-3
    r = np.full(5000, 1.0)
    assert not Grid(cg, c, r).enabled                               # 5000 spheres in every cell
