"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from eraytracer_b200 import _lib, scene as sc
from helpers import (assert_double_parity, assert_image_parity, oracle_frame,
                     oracle_scene_from_flat, quantise)
from oracle import orc

pytestmark = pytest.mark.gpu

ACCELS = ("exact", "linear", "bvh", "bvh_mega")


@pytest.fixture(scope="module")
def demo(gpu):
    flat = sc.flatten(sc.demo_scene())
    dev = flat.upload(0)
    yield flat, dev
    dev.close()


@pytest.fixture(scope="module")
def c3(gpu):
    flat = sc.synthetic_scene("c3")
    dev = flat.upload(0)
    yield flat, dev
    dev.close()


# ---- the reference's own known-answer tests, through the CUDA path -----------------------
def _sphere(r, c, colour=(0.4, 0.4, 0.4)):
    return ('sphere', r, ('vector',) + tuple(c), ('material', ('colour',) + tuple(colour), 1, 0, 0))


CAM = ('camera', ('vector', 0, 0, 0), ('vector', 0, 0, 0), 90, ('screen', 1, 1))


@pytest.mark.parametrize("accel", ACCELS)
def test_ray_sphere_intersection_kat(gpu, accel):
    """ray_sphere_intersection_test, raytracer.erl:1013-1034."""
    dev = sc.flatten([CAM, _sphere(3, (0, 0, 10))]).upload(0)
    rays = np.array([[0, 0, 0, 0, 0, 1], [3, 0, 0, 0, 0, 1], [4, 0, 0, 0, 0, 1]], dtype=np.float64)
    order, t = dev.trace_rays(rays, accel=accel)
    assert order.tolist() == [0, -1, -1]
    assert t[0] == 7.0
    dev.close()


@pytest.mark.parametrize("accel", ACCELS)
def test_nearest_object_intersecting_ray_kat(gpu, accel):
    """nearest_object_intersecting_ray_test, raytracer.erl:1068-1097."""
    spheres = [_sphere(5, (0, 0, 10)), _sphere(5, (0, 0, 20)), _sphere(5, (0, 0, 30)),
               _sphere(5, (0, 0, -10))]
    dev = sc.flatten([CAM] + spheres).upload(0)
    order, t = dev.trace_rays(np.array([[0, 0, 0, 0, 0, 1]], dtype=np.float64), accel=accel)
    assert order[0] == 0 and t[0] == 5
    dev.close()


# ---- demo scene (configs C1, C2 and the reference's defaults) ----------------------------
@pytest.mark.parametrize("accel", ACCELS)
@pytest.mark.parametrize("size", [(4, 3, 5), (16, 12, 1), (32, 24, 1), (32, 24, 5), (33, 17, 3),
                                  (320, 240, 1), (320, 240, 5), (640, 480, 1)])
def test_demo_scene_matches_oracle(demo, accel, size):
    flat, dev = demo
    w, h, depth = size
    ref, ref_rays, _ = oracle_frame(flat, w, h, depth)
    frame, st = dev.render(w, h, depth, fmt="f64", accel=accel)
    assert st["accel_used"] == accel
    assert_double_parity(frame, ref)
    r = assert_image_parity(quantise(frame), quantise(ref))
    assert r["max"] == 0
    assert st["rays"] == ref_rays        # unique rays: one per nearest-object scan


def test_demo_scene_1080p_depth5_all_formats(demo):
    """Config C2 at the reference's default depth, every output format."""
    flat, dev = demo
    w, h, depth = 1920, 1080, 5
    ref, _, _ = oracle_frame(flat, w, h, depth)
    ref_q = quantise(ref)
    f64, _ = dev.render(w, h, depth, fmt="f64")
    assert_double_parity(f64, ref)
    rgb8, _ = dev.render(w, h, depth, fmt="rgb8")
    assert np.array_equal(rgb8.astype(np.int64), np.maximum(ref_q, 0))
    f32, _ = dev.render(w, h, depth, fmt="f32")
    assert np.array_equal(f32, f64.astype(np.float32))
    assert_image_parity(quantise(f32), ref_q)


def test_demo_scene_1080p_depth1(demo):
    """Config C2 with run-concurrent.sh's depth (1)."""
    flat, dev = demo
    ref, _, _ = oracle_frame(flat, 1920, 1080, 1)
    frame, _ = dev.render(1920, 1080, 1, fmt="f64", accel="bvh")
    assert_double_parity(frame, ref)
    assert np.array_equal(quantise(frame), quantise(ref))


# ---- synthetic 10k-sphere scene (config C3) at oracle-sized resolutions ------------------
@pytest.mark.parametrize("accel", ("linear", "bvh", "bvh_mega"))
def test_c3_matches_oracle(c3, accel):
    flat, dev = c3
    w, h, depth = 192, 108, 5
    ref, ref_rays, _ = oracle_frame(flat, w, h, depth)
    frame, st = dev.render(w, h, depth, fmt="f64", accel=accel)
    assert_double_parity(frame, ref)
    assert np.array_equal(quantise(frame), quantise(ref))
    assert st["rays"] == ref_rays


def test_c3_exact_equals_bvh_equals_linear_on_gpu(c3):
    flat, dev = c3
    w, h, depth = 480, 270, 5
    a, _ = dev.render(w, h, depth, fmt="f64", accel="exact")
    b, _ = dev.render(w, h, depth, fmt="f64", accel="bvh")
    c, _ = dev.render(w, h, depth, fmt="f64", accel="linear")
    d, _ = dev.render(w, h, depth, fmt="f64", accel="bvh_mega")
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)
    assert np.array_equal(a, d)


def test_c3_ray_batch_bvh_equals_linear_scan(c3):
    """The BVH must return the linear scan's nearest hit (distance bits and list position)."""
    flat, dev = c3
    rng = np.random.default_rng(7)
    n = 200_000
    o = np.stack([rng.uniform(-45, 45, n), rng.uniform(-35, 5, n), rng.uniform(-5, 90, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1)
    oe, te = dev.trace_rays(rays, accel="exact")
    ob, tb = dev.trace_rays(rays, accel="bvh")
    om, tm = dev.trace_rays(rays, accel="bvh_mega")
    ol, tl = dev.trace_rays(rays, accel="linear")
    assert np.array_equal(oe, ob) and np.array_equal(te, tb)
    assert np.array_equal(oe, om) and np.array_equal(te, tm)
    assert np.array_equal(oe, ol) and np.array_equal(te, tl)
    # and the oracle agrees on a subset
    cam, kind, f = oracle_scene_from_flat(flat)
    oi, ot = orc.nearest_batch(rays[:4000], kind, f)
    assert np.array_equal(oi, oe[:4000]) and np.array_equal(ot, te[:4000])
    assert (oe >= 0).mean() > 0.2


def test_warp_wide_nearest_hit_reduction_equals_the_scan(c3, gpu):
    """ERT_ACCEL_WARP (one warp per ray, lanes stride over the spheres, warp-wide lexicographic minimum of
    (Distance, list position)) returns the linear scan's nearest object: erl:300-346."""
    flat, dev = c3
    rng = np.random.default_rng(11)
    n = 60_000
    o = np.stack([rng.uniform(-45, 45, n), rng.uniform(-35, 5, n), rng.uniform(-5, 90, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1)
    oe, te = dev.trace_rays(rays, accel="exact")
    ow, tw = dev.trace_rays(rays, accel="warp")
    assert np.array_equal(oe, ow) and np.array_equal(te, tw)
    assert (ow >= 0).mean() > 0.2 and (ow < 0).any()
    # frames are rendered by the per-pixel kernels: the warp form serves ray batches only
    with pytest.raises(_lib.ErtError):
        dev.render(16, 16, 1, accel=6)


def test_warp_wide_reduction_ties_planes_triangles_and_small_scenes(gpu):
    """Exact Distance ties go to the smaller list position, a triangle hit behind the origin (erl:402-455 has no
    T >= 0 test) beats every sphere, and scenes with fewer spheres than lanes work."""
    CAM = ('camera', ('vector', 0, 0, -2), ('vector', 0, 0, 0), 90, ('screen', 4, 3))
    L1 = ('point_light', ('colour', 1, 1, 0.5), ('vector', 5, -2, 0), ('colour', 1, 1, 1))
    m = ('material', ('colour', 0.5, 0.5, 0.5), 4, 0.5, 0.3)
    scene = [CAM, L1,
             ('sphere', 1.0, ('vector', 0, 0, 10), m), ('sphere', 1.0, ('vector', 0, 0, 10), m),     # exact tie
             ('sphere', 0.5, ('vector', 3, 0, 8), m),
             ('plane', ('vector', 0, -1, 0), 3, m),
             ('triangle', ('vector', -1, -1, -4), ('vector', 1, -1, -4), ('vector', 0, 1, -4), m)]
    flat = sc.flatten(scene)
    dev = flat.upload(0)
    rays = np.array([[0, 0, 0, 0, 0, 1], [0, 0, 0, 3, 0, 8], [0, 0, 0, 0, 1, 0.2], [0, 0, 0, 0, -1, 0.1],
                     [0, 0, 0, 0.01, 0.02, 1], [5, 5, 5, 1, 0, 0]], dtype=np.float64)
    oe, te = dev.trace_rays(rays, accel="exact")
    ow, tw = dev.trace_rays(rays, accel="warp")
    assert np.array_equal(oe, ow) and np.array_equal(te, tw)
    dev.close()
