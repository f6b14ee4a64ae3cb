"""CPU tests of the per-light direction grid (eraytracer_b200/csrc/light_grid.cpp).

The grid only prunes shadow-ray candidates; what it must guarantee is completeness: every sphere
a ray from the light can touch is listed in the ray's cell (or in the `always` list), and the
entries of a cell come nearest first with a valid lower bound of their distance.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "lightgrid", "_build")
SO = os.path.join(BUILD, "liblightgrid_test.so")

ENTRY_DT = np.dtype([("sphere", "<i4"), ("dmin", "<f4")])


@pytest.fixture(scope="module")
def lg():
    os.makedirs(BUILD, exist_ok=True)
    src = [os.path.join(HERE, "lightgrid", "shim.cpp"),
           os.path.join(ROOT, "eraytracer_b200", "csrc", "light_grid.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO] + src + ["-lpthread"])
    L = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    L.lg_build.restype = vp
    L.lg_build.argtypes = [vp, vp, vp, ctypes.c_longlong, vp, ctypes.c_int]
    L.lg_free.argtypes = [vp]
    L.lg_cell.restype = ctypes.c_longlong
    L.lg_cell.argtypes = [vp, ctypes.c_int]
    L.lg_cell_f32.restype = ctypes.c_longlong
    L.lg_cell_f32.argtypes = [vp, ctypes.c_int]
    for name in ("lg_n_entries", "lg_n_always"):
        getattr(L, name).restype = ctypes.c_longlong
        getattr(L, name).argtypes = [vp]
    for name in ("lg_offsets", "lg_entries", "lg_always"):
        getattr(L, name).restype = vp
        getattr(L, name).argtypes = [vp]
    return L


def build(L, centers, radii, light, res):
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    radii = np.ascontiguousarray(radii, dtype=np.float64)
    filt = np.zeros((len(radii), 4), dtype=np.float32)
    filt[:, :3] = centers
    filt[:, 3] = radii ** 2
    light = np.ascontiguousarray(light, dtype=np.float64)
    h = L.lg_build(centers.ctypes.data, radii.ctypes.data, filt.ctypes.data, len(radii), light.ctypes.data, res)
    n_cells = 6 * res * res
    off = np.ctypeslib.as_array(ctypes.cast(L.lg_offsets(h), ctypes.POINTER(ctypes.c_uint32)), (n_cells + 1,)).copy()
    n_e = L.lg_n_entries(h)
    ent = np.frombuffer((ctypes.c_char * (n_e * 8)).from_address(L.lg_entries(h)), dtype=ENTRY_DT).copy() \
        if n_e else np.zeros(0, dtype=ENTRY_DT)
    n_a = L.lg_n_always(h)
    always = np.ctypeslib.as_array(ctypes.cast(L.lg_always(h), ctypes.POINTER(ctypes.c_int32)), (n_a,)).copy() \
        if n_a else np.zeros(0, dtype=np.int32)
    L.lg_free(h)
    return off, ent, always


def cells_of(L, dirs, res):
    dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    return np.array([L.lg_cell(dirs[i].ctypes.data, res) for i in range(len(dirs))], dtype=np.int64)


def cells_of_f32(L, dirs, res, scale=1.0000019073486328):
    """The cell the shadow kernel computes: the direction scaled like the filter ray (K_D), rounded to float32."""
    d32 = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64) * scale, dtype=np.float32)
    return np.array([L.lg_cell_f32(d32[i].ctypes.data, res) for i in range(len(d32))], dtype=np.int64)


def touched(light, d, centers, radii):
    """Spheres the ray light + s*d (s >= 0, |d| = 1) touches, geometrically, in double."""
    oc = centers - light
    b = oc @ d
    disc = b * b - (np.einsum("ij,ij->i", oc, oc) - radii ** 2)
    return np.nonzero((disc >= 0) & (b + np.sqrt(np.maximum(disc, 0)) >= 0))[0]


@pytest.mark.parametrize("res", [16, 128])
def test_every_touched_sphere_is_listed(lg, res):
    rng = np.random.default_rng(5)
    n = 4000
    centers = rng.uniform(-40, 40, (n, 3))
    radii = rng.uniform(0.2, 1.5, n)
    light = np.array([3.0, -7.0, 1.5])
    # a sphere that contains the light, one that just does not, one straddling a face plane
    centers[0], radii[0] = light + [0.1, 0.2, -0.1], 1.0
    centers[1], radii[1] = light + [2.0, 0.0, 0.0], 1.999
    centers[2], radii[2] = light + [5.0, 5.0, 0.0], 2.0
    off, ent, always = build(lg, centers, radii, light, res)
    assert 0 in always and 1 not in always
    # directions: random, towards sphere rims (tangent rays), and exactly on face edges / corners
    d = rng.normal(size=(3000, 3))
    to_c = centers[rng.integers(0, n, 3000)] - light
    perp = np.cross(to_c, rng.normal(size=(3000, 3)))
    perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    rim = to_c + perp * radii[rng.integers(0, n, 3000), None] * rng.uniform(0.9, 1.0, (3000, 1))
    special = np.array([[1, 1, 0], [1, -1, 0], [0, 1, 1], [1, 1, 1], [-1, 1, 1], [1, 0, 0], [0, 0, -1], [-1, -1, -1],
                        [1, 1, 1e-17], [1, 1 - 1e-16, 0]], dtype=np.float64)
    dirs = np.concatenate([d, rim, special])
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    cells = cells_of(lg, dirs, res)
    cells32 = cells_of_f32(lg, dirs, res)
    assert cells.min() >= 0 and cells.max() < 6 * res * res
    assert cells32.min() >= 0 and cells32.max() < 6 * res * res
    assert (cells32 != cells).mean() < 0.01            # they differ only for directions on a cell boundary
    n_touch = 0
    for k in range(len(dirs)):
        hit = touched(light, dirs[k], centers, radii)
        n_touch += len(hit)
        # the FP64 cell (host builder's own rule) and the FP32 cell the shadow kernel looks in
        for c in (cells[k], cells32[k]):
            listed = set(ent["sphere"][off[c]:off[c + 1]].tolist()) | set(always.tolist())
            assert set(hit.tolist()) <= listed, (k, dirs[k])
    assert n_touch > 3000


def test_cells_are_sorted_nearest_first_with_a_valid_lower_bound(lg):
    rng = np.random.default_rng(9)
    n = 3000
    centers = rng.uniform(-30, 30, (n, 3))
    radii = rng.uniform(0.2, 1.0, n)
    light = np.array([0.0, -50.0, 10.0])
    off, ent, always = build(lg, centers, radii, light, 32)
    assert off[0] == 0 and off[-1] == len(ent) and np.all(np.diff(off.astype(np.int64)) >= 0)
    true_min = np.linalg.norm(centers - light, axis=1) - radii
    assert np.all(ent["dmin"].astype(np.float64) <= true_min[ent["sphere"]])
    for c in range(len(off) - 1):
        seg = ent["dmin"][off[c]:off[c + 1]]
        assert np.all(np.diff(seg) >= 0)
    # each sphere appears at most once per cell
    cell_of_entry = np.repeat(np.arange(len(off) - 1), np.diff(off.astype(np.int64)))
    pairs = cell_of_entry.astype(np.int64) * n + ent["sphere"]
    assert len(np.unique(pairs)) == len(pairs)


def test_empty_and_degenerate_inputs(lg):
    off, ent, always = build(lg, np.zeros((0, 3)), np.zeros(0), np.zeros(3), 8)
    assert len(ent) == 0 and len(always) == 0 and off[-1] == 0
    # the light at the centre of the only sphere; a zero-radius sphere elsewhere
    off, ent, always = build(lg, np.array([[1.0, 2.0, 3.0], [9.0, 2.0, 3.0]]), np.array([2.0, 0.0]),
                             np.array([1.0, 2.0, 3.0]), 8)
    assert always.tolist() == [0]
    assert set(ent["sphere"].tolist()) == {1}
