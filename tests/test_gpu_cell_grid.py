"""GPU parity tests of the cell-grid walk (ERT_ACCEL_GRID): the path rays of the wavefront step
through a uniform grid over the spheres instead of walking the BVH.  The grid only prunes, so
every result must equal the linear scan's (raytracer.erl:300-346): same distance bits, same list
position, bit-identical frames."""
import numpy as np
import pytest

import eraytracer_b200 as ert
from eraytracer_b200 import _lib, multigpu, scene as sc
from oracle import orc
from helpers import assert_double_parity, clustered_scene, oracle_frame, oracle_scene_from_flat, quantise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3(gpu):
    flat = sc.synthetic_scene("c3")
    dev = flat.upload(0)
    yield flat, dev
    dev.close()


def random_rays(rng, n, lo, hi):
    o = np.stack([rng.uniform(lo[a], hi[a], n) for a in range(3)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)


def test_auto_prefers_the_grid_and_counts_cells(c3):
    flat, dev = c3
    frame, st = dev.render(96, 54, 3, fmt="f64", accel="auto", flags=_lib.FLAG_COUNT_TESTS)
    assert st["accel_used"] == "grid" and st["has_cell_grid"] == 1
    assert st["cell_steps"] > 96 * 54 and st["path_box_tests"] == 0
    bvh, sb = dev.render(96, 54, 3, fmt="f64", accel="bvh", flags=_lib.FLAG_COUNT_TESTS)
    assert sb["accel_used"] == "bvh" and sb["cell_steps"] == 0 and sb["path_box_tests"] > 0
    assert np.array_equal(frame, bvh) and st["rays"] == sb["rays"]


def test_c3_grid_matches_oracle(c3):
    flat, dev = c3
    w, h, depth = 192, 108, 5
    ref, ref_rays, _ = oracle_frame(flat, w, h, depth)
    frame, st = dev.render(w, h, depth, fmt="f64", accel="grid")
    assert st["accel_used"] == "grid"
    assert_double_parity(frame, ref)
    assert np.array_equal(quantise(frame), quantise(ref))
    assert st["rays"] == ref_rays


def test_c3_grid_equals_exact_scan_on_gpu(c3):
    flat, dev = c3
    w, h, depth = 480, 270, 6
    a, sa = dev.render(w, h, depth, fmt="f64", accel="exact")
    b, sb = dev.render(w, h, depth, fmt="f64", accel="grid")
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]
    # row bands of the multi-GPU split
    out = np.zeros_like(b)
    for part in range(3):
        dev.render(w, h, depth, fmt="f64", accel="grid", band_rows=8, n_parts=3, part=part, out=out)
    assert np.array_equal(out, b)


def test_c3_ray_batch_grid_equals_linear_scan(c3):
    """Distance bits and list position, for rays from inside and outside the grid, rays along
    the axes, rays that start exactly on cell planes, and rays with zero components."""
    flat, dev = c3
    rng = np.random.default_rng(17)
    rays = [random_rays(rng, 150_000, (-45, -35, -5), (45, 5, 90)),
            random_rays(rng, 30_000, (-400, -400, -400), (400, 400, 400))]
    # axis-parallel and planar directions
    ax = random_rays(rng, 30_000, (-45, -35, -5), (45, 5, 90))
    k = rng.integers(0, 3, len(ax))
    ax[:, 3:] = 0.0
    ax[np.arange(len(ax)), 3 + k] = rng.choice([-1.0, 1.0], len(ax))
    pl = random_rays(rng, 30_000, (-45, -35, -5), (45, 5, 90))
    pl[np.arange(len(pl)), 3 + rng.integers(0, 3, len(pl))] = 0.0
    pl[:, 3:] /= np.maximum(np.linalg.norm(pl[:, 3:], axis=1, keepdims=True), 1e-30)
    # origins snapped to a lattice that contains the cell planes of any power-of-two-ish edge
    sn = random_rays(rng, 30_000, (-45, -35, -5), (45, 5, 90))
    sn[:, :3] = np.round(sn[:, :3] * 4) / 4
    # un-normalised directions (reflections off un-normalised plane normals, erl:461-480)
    un = random_rays(rng, 20_000, (-45, -35, -5), (45, 5, 90))
    un[:, 3:] *= rng.uniform(0.01, 50.0, (len(un), 1))
    rays += [ax, pl, sn, un]
    rays = np.concatenate(rays, axis=0)
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    ob, tb = dev.trace_rays(rays, accel="bvh")
    assert np.array_equal(og, ol) and np.array_equal(tg, tl)
    assert np.array_equal(og, ob) and np.array_equal(tg, tb)
    assert (og >= 0).mean() > 0.2
    cam, kind, f = oracle_scene_from_flat(flat)
    oi, ot = orc.nearest_batch(rays[:3000], kind, f)
    assert np.array_equal(oi, og[:3000]) and np.array_equal(ot, tg[:3000])


def stress_scene(n, seed, box, radii, big=0, scale=1.0, offset=(0.0, 0.0, 0.0), float_exact=True):
    """n small spheres in a box plus `big` large ones, demo floor plane, two lights."""
    rng = np.random.default_rng(seed)
    flat = sc.synthetic_scene("c3", n_spheres=n + big, seed=seed)
    c = np.stack([rng.uniform(-box[a], box[a], n + big) for a in range(3)], axis=1) * scale + np.asarray(offset)
    r = rng.uniform(radii[0], radii[1], n + big) * scale
    if big:
        r[-big:] = rng.uniform(10.0, 40.0, big) * scale
    if float_exact:
        c = c.astype(np.float32).astype(np.float64)
        r = r.astype(np.float32).astype(np.float64)
    flat.spheres['center'] = c
    flat.spheres['radius'] = r
    return flat


@pytest.mark.parametrize("case", ("big_spheres", "large_coordinates", "double_centres", "thin_slab", "dense"))
def test_grid_stress_scenes_equal_linear_scan(gpu, case):
    if case == "big_spheres":
        flat = stress_scene(4000, 5, (40, 20, 40), (0.1, 0.6), big=5, offset=(0, -20, 50))
        lo, hi = (-60, -60, -10), (60, 10, 110)
    elif case == "large_coordinates":
        flat = stress_scene(6000, 6, (40, 20, 40), (0.1, 0.8), scale=100.0, offset=(9000.0, -3000.0, 20000.0))
        lo, hi = (4000, -6000, 15000), (14000, 0, 25000)
    elif case == "double_centres":
        flat = stress_scene(5000, 7, (30, 15, 30), (0.05, 0.7), offset=(0.1, -17.3, 40.7), float_exact=False)
        lo, hi = (-35, -35, 5), (35, 5, 75)
    elif case == "thin_slab":
        flat = stress_scene(3000, 8, (50, 0.01, 50), (0.2, 0.5), offset=(0, -3, 60))
        lo, hi = (-55, -10, 5), (55, 4, 115)
    else:
        flat = stress_scene(20000, 9, (6, 6, 6), (0.2, 0.6), offset=(0, -8, 20))
        lo, hi = (-8, -16, 10), (8, 0, 30)
    dev = flat.upload(0)
    rng = np.random.default_rng(3)
    rays = np.concatenate([random_rays(rng, 60_000, lo, hi),
                           random_rays(rng, 20_000, [3 * v - 50 for v in lo], [3 * v + 50 for v in hi])], axis=0)
    # half of the outside rays aim at the scene
    mid = (np.asarray(lo) + np.asarray(hi)) / 2
    aim = mid + rng.normal(size=(10_000, 3)) * (np.asarray(hi) - np.asarray(lo)) / 4 - rays[-10_000:, :3]
    rays[-10_000:, 3:] = aim / np.linalg.norm(aim, axis=1, keepdims=True)
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    assert np.array_equal(og, ol) and np.array_equal(tg, tl)
    assert (og >= 0).mean() > 0.05
    w, h, depth = 160, 90, 4
    a, sa = dev.render(w, h, depth, fmt="f64", accel="grid")
    b, sb = dev.render(w, h, depth, fmt="f64", accel="bvh")
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]
    if case != "dense":
        assert sa["accel_used"] == "grid", "the builder should accept this scene"
    dev.close()


def test_far_origins_fall_back_to_the_bvh_walk(c3):
    """A ray whose FP32 margin exceeds the inflation of the cells (origin far outside the
    scene) must not use the grid; the result is still the linear scan's."""
    flat, dev = c3
    rng = np.random.default_rng(23)
    n = 20_000
    target = np.stack([rng.uniform(-40, 40, n), rng.uniform(-30, 4, n), rng.uniform(5, 85, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dist = 10.0 ** rng.uniform(2, 7, (n, 1))
    rays = np.concatenate([target - d * dist, d], axis=1)
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    assert np.array_equal(og, ol) and np.array_equal(tg, tl)
    assert (og >= 0).sum() > 100
    # a camera that far away renders through the same fallback
    cam = sc.pose_camera(0)
    cam.location[:] = (0.0, -10.0, -3.0e5)
    cam.fov = 0.02
    a, sa = dev.render(64, 36, 3, fmt="f64", accel="grid", camera=cam)
    b, sb = dev.render(64, 36, 3, fmt="f64", accel="exact", camera=cam)
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]


def test_scenes_without_a_grid_use_the_bvh(gpu):
    # too few spheres for a grid; ERT_ACCEL_GRID then means the BVH wavefront
    flat = sc.synthetic_scene("c3", n_spheres=200)
    dev = flat.upload(0)
    f, st = dev.render(64, 36, 3, fmt="f64", accel="grid")
    assert st["accel_used"] == "bvh" and st["has_cell_grid"] == 0
    e, _ = dev.render(64, 36, 3, fmt="f64", accel="exact")
    assert np.array_equal(f, e)
    dev.close()
    # radii spread over five decades: most spheres would be `big`
    rng = np.random.default_rng(1)
    flat = sc.synthetic_scene("c3", n_spheres=3000)
    flat.spheres['radius'] = (10.0 ** rng.uniform(-2, 3, 3000)).astype(np.float32)
    dev = flat.upload(0)
    f, st = dev.render(64, 36, 3, fmt="f64", accel="auto")
    assert st["accel_used"] == "bvh" and st["has_cell_grid"] == 0
    e, _ = dev.render(64, 36, 3, fmt="f64", accel="exact")
    assert np.array_equal(f, e)
    dev.close()


def test_c4_full_4k_grid_equals_bvh(gpu):
    """Config C4 at full size: the grid frame is the BVH frame, bit for bit (both doubles and RGB8)."""
    flat = sc.synthetic_scene("c4")
    dev = flat.upload(0)
    w, h, depth = 3840, 2160, 5
    a, sa = dev.render(w, h, depth, fmt="f64", accel="grid")
    assert sa["accel_used"] == "grid"
    b, sb = dev.render(w, h, depth, fmt="f64", accel="bvh")
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]
    del b
    sub = np.zeros_like(a)
    dev.render(w, h, depth, fmt="f64", accel="linear", band_rows=1, n_parts=1080, part=333, out=sub)
    rows = multigpu.part_rows(h, 1, 1080, 333)
    assert np.array_equal(sub[rows], a[rows])
    rng = np.random.default_rng(12)
    rays = random_rays(rng, 20000, (-210, -160, -5), (210, 5, 410))
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    assert np.array_equal(og, ol) and np.array_equal(tg, tl)
    dev.close()


def test_per_level_counts_and_the_reference_ray_count(c3):
    """ert_stats.bounce_*: rays = sum_b (path[b] + L*hits[b]); the reference's own count for the frame
    (reflection re-traced once per light, erl:216-224) = sum_b L^b * (path[b] + L*hits[b]) — checked
    against the oracle in its literal re-tracing mode."""
    flat, dev = c3
    w, h, depth = 48, 27, 4
    for accel in ("grid", "bvh"):
        frame, st = dev.render(w, h, depth, fmt="f64", accel=accel)
        L = len(flat.lights)
        assert st["bounces_recorded"] == depth and st["bounce_path_rays"][0] == w * h
        assert st["rays"] == sum(p + L * q for p, q in zip(st["bounce_path_rays"], st["bounce_hits"]))
        cam, kind, f = oracle_scene_from_flat(flat)
        ref, ref_rays, _ = orc.render(cam, kind, f, w, h, depth, retrace=True)
        assert ref_rays == sum((L ** b) * (p + L * q)
                               for b, (p, q) in enumerate(zip(st["bounce_path_rays"], st["bounce_hits"])))
        assert np.array_equal(quantise(frame), quantise(ref.reshape(h, w, 3)))
    # frames that do not go through the wavefront record no levels
    _, st = dev.render(w, h, depth, fmt="f64", accel="exact")
    assert st["bounces_recorded"] == 0 and st["bounce_hits"] == []


def _with_extras(flat, triangles=None, planes=None, lights=None):
    """Replaces the triangle / plane / light tables of a synthetic scene, list positions re-dealt:
    lights, spheres, triangles, planes."""
    if lights is not None:
        flat.lights = lights
    if triangles is not None:
        flat.triangles = triangles
    if planes is not None:
        flat.planes = planes
    k = 0
    for tab in (flat.lights, flat.spheres, flat.triangles, flat.planes):
        tab['order'] = np.arange(k, k + len(tab), dtype=np.int32)
        k += len(tab)
    return flat


@pytest.mark.parametrize("case", ("twelve_lights", "no_lights", "deep", "skew_plane_and_triangle", "camera_in_the_cloud"))
def test_grid_path_edge_scenes_match_the_oracle(gpu, case):
    """The reference's awkward corners, on scenes large enough for the cell grid."""
    flat = sc.synthetic_scene("c3", n_spheres=1200)
    w, h, depth, camera = 64, 36, 3, None
    if case == "twelve_lights":               # lights past the eighth send their shadow rays through the BVH
        lights = np.zeros(12, dtype=_lib.LIGHT_DT)
        for k in range(12):
            lights[k]['diffuse_colour'] = (0.05 * k, 0.3, 0.6 - 0.04 * k)
            lights[k]['location'] = (7.0 * k - 35, -25 + 3 * k, 2.0 * k)
            lights[k]['specular_colour'] = (1, 1, 1)
        flat = _with_extras(flat, lights=lights)
    elif case == "no_lights":                 # the fold over no lights is black (erl:211-252)
        flat = _with_extras(flat, lights=np.zeros(0, dtype=_lib.LIGHT_DT))
    elif case == "deep":                      # the queue runs dry long before depth 14
        depth = 14
        flat.spheres['material']['reflectivity'] = np.where(np.arange(len(flat.spheres)) % 3 == 0, 0.0, 0.9)
    elif case == "skew_plane_and_triangle":   # un-normalised normal => non-unit reflection rays; t < 0 triangle hits
        planes = np.zeros(2, dtype=_lib.PLANE_DT)
        planes['normal'] = [(0, -1, 0), (0.3, -2.5, 0.4)]
        planes['distance'] = [5, 40]
        planes['material']['colour'] = [(1, 1, 1), (0.4, 0.8, 0.9)]
        planes['material']['specular_power'] = [1, 4]
        planes['material']['shininess'] = [0, 0.5]
        planes['material']['reflectivity'] = [0.01, 0.6]
        tris = np.zeros(2, dtype=_lib.TRIANGLE_DT)
        tris['v1'] = [(-20, 4, 20), (-10, -20, -40)]
        tris['v2'] = [(25, 4, 60), (10, -20, -40)]
        tris['v3'] = [(25, -30, 60), (0, 5, -40)]
        tris['material']['colour'] = [(1, 0.5, 0), (0.2, 0.9, 0.3)]
        tris['material']['specular_power'] = 4
        tris['material']['shininess'] = 0.25
        tris['material']['reflectivity'] = 0.5
        flat = _with_extras(flat, triangles=tris, planes=planes)
    else:                                     # origins inside the grid, some inside spheres
        camera = sc.pose_camera(0)
        camera.location[:] = (1.5, -12.0, 40.0)
        camera.screen_width, camera.screen_height = 4.0, 2.25
    dev = flat.upload(0)
    if camera is not None:
        flat.camera = camera
    ref, ref_rays, _ = oracle_frame(flat, w, h, depth)
    for accel in ("grid", "bvh"):
        frame, st = dev.render(w, h, depth, fmt="f64", accel=accel, camera=camera)
        assert st["accel_used"] == accel
        assert_double_parity(frame, ref)
        assert np.array_equal(quantise(frame), quantise(ref)), (case, accel)
        # a hit of zero reflectivity ends the path on the GPU (0 * child adds nothing); the reference and
        # the oracle still trace that reflection, so only the "deep" case may count fewer rays
        assert st["rays"] == ref_rays or (case == "deep" and st["rays"] < ref_rays), (case, accel)
    dev.close()


def test_committed_synthetic_golden_fixture(gpu):
    """Every strategy against tests/golden/synthetic_images.json (no oracle build needed)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synthetic_images.json")))
    for key, img in g["images"].items():
        dev = sc.synthetic_scene("c3", n_spheres=img["n_spheres"]).upload(0)
        for accel in ("auto", "grid", "bvh", "linear", "bvh_mega", "exact"):
            rgb8, st = dev.render(img["width"], img["height"], img["depth"], fmt="rgb8", accel=accel)
            assert rgb8.reshape(-1).tolist() == img["rgb8"], (key, accel)
            assert st["rays"] == img["rays"], (key, accel)
        dev.close()


# ---- the FP32 triage of shadow rays (shadow_blocked in ert_wavefront.cuh) ---------------------------------------
def _lights(rows):
    lights = np.zeros(len(rows), dtype=_lib.LIGHT_DT)
    for k, (loc, col) in enumerate(rows):
        lights[k]['location'] = loc
        lights[k]['diffuse_colour'] = col
        lights[k]['specular_colour'] = (1, 1, 1)
    return lights


@pytest.mark.parametrize("case", ("c3", "lights_among_the_spheres", "far_and_unrounded_lights", "large_coordinates",
                                  "double_centres", "radii_over_four_decades", "touching_pairs", "skew_plane"))
def test_shadow_triage_equals_the_literal_shadow_path(gpu, case):
    """A shadow ray the triage declares blocked must be one the literal path (FP64 direction, literal tests, here
    with the shadow rays walking the BVH: ERT_FLAG_NO_LIGHT_GRID) finds blocked: frames and ray counts are equal,
    on scenes that push the triage's margins — lights a hair's breadth from spheres, lights whose coordinates are
    not float32 values, coordinates around 1e4, centres that are not float32 values, radii from 0.01 to 100,
    occluders that touch their targets, an un-normalised plane."""
    w, h, depth = 320, 180, 4
    if case == "c3":
        flat = sc.synthetic_scene("c3")
        w, h = 960, 540
    elif case == "lights_among_the_spheres":
        flat = sc.synthetic_scene("c3", n_spheres=6000)
        c, r = flat.spheres['center'], flat.spheres['radius']
        rows = []
        for k, s in enumerate((11, 222, 3333, 4444)):
            # just outside sphere s, along a diagonal: 1.0005 r, 1.02 r, 1.5 r, 3 r from its centre
            d = np.array([0.6, -0.64, 0.48]) * r[s] * (1.0005, 1.02, 1.5, 3.0)[k]
            rows.append((tuple(c[s] + d), (1.0, 0.3 + 0.2 * k, 0.9 - 0.2 * k)))
        flat = _with_extras(flat, lights=_lights(rows))
    elif case == "far_and_unrounded_lights":
        flat = sc.synthetic_scene("c3", n_spheres=5000)
        flat = _with_extras(flat, lights=_lights([((1234.56789, -9876.54321, -4321.123456789), (1, 1, 1)),
                                                   ((0.1, -40.7, 33.3), (1, 0.5, 0.2)),
                                                   ((-1e4 / 3, -1e3 / 7, 5e3 / 9), (0.2, 0.6, 1.0))]))
    elif case == "large_coordinates":
        flat = stress_scene(6000, 6, (40, 20, 40), (0.1, 0.8), scale=100.0, offset=(9000.0, -3000.0, 20000.0))
        flat = _with_extras(flat, lights=_lights([((9500.0, -8000.0, 15000.0), (1, 1, 0.5)),
                                                   ((6000.1, -3500.7, 21000.3), (1, 0, 0.5))]))
        cam = sc.pose_camera(0)
        cam.location[:] = (9000.0, -3500.0, 14000.0)
        cam.screen_width, cam.screen_height = 4.0, 2.25
        flat.camera = cam
    elif case == "double_centres":
        flat = stress_scene(5000, 7, (30, 15, 30), (0.05, 0.7), offset=(0.1, -17.3, 40.7), float_exact=False)
    elif case == "radii_over_four_decades":
        flat = stress_scene(4000, 21, (40, 15, 40), (0.2, 0.6), offset=(0, -16, 45))
        rng = np.random.default_rng(21)
        flat.spheres['radius'] = (10.0 ** rng.uniform(-2, 0.3, len(flat.spheres))).astype(np.float32).astype(np.float64)
        flat.spheres['radius'][:3] = (100.0, 30.0, 12.0)
        flat.spheres['center'][:3] = [(0, -140, 160), (-60, -40, 90), (45, -20, 70)]
    elif case == "touching_pairs":
        flat = stress_scene(3000, 22, (35, 12, 35), (0.3, 0.7), offset=(0, -14, 45))
        c, r = flat.spheres['center'], flat.spheres['radius']
        # every second sphere touches (or slightly overlaps) its predecessor, on the side of the first light
        to_light = np.array([5.0, -20.0, 0.0]) - c[0::2]
        to_light /= np.linalg.norm(to_light, axis=1, keepdims=True)
        c[1::2] = c[0::2] + to_light * ((r[0::2] + r[1::2]) * np.where(np.arange(len(c) // 2) % 2 == 0, 1.0, 0.98))[:, None]
        flat.spheres['center'] = c
    else:
        flat = sc.synthetic_scene("c3", n_spheres=4000)
        planes = np.zeros(2, dtype=_lib.PLANE_DT)
        planes['normal'] = [(0, -1, 0), (0.3, -2.5, 0.4)]
        planes['distance'] = [5, 40]
        planes['material']['colour'] = [(1, 1, 1), (0.4, 0.8, 0.9)]
        planes['material']['specular_power'] = [1, 4]
        planes['material']['shininess'] = [0, 0.5]
        planes['material']['reflectivity'] = [0.01, 0.6]
        flat = _with_extras(flat, planes=planes)
    dev = flat.upload(0)
    cam = getattr(flat, "camera", None) if case == "large_coordinates" else None
    a, sa = dev.render(w, h, depth, fmt="f64", accel="auto", camera=cam, flags=_lib.FLAG_COUNT_TESTS)
    b, sb = dev.render(w, h, depth, fmt="f64", accel="auto", camera=cam, flags=_lib.FLAG_NO_LIGHT_GRID)
    assert np.array_equal(a, b), case
    assert sa["rays"] == sb["rays"] and sa["rays"] > 2 * w * h
    assert sa["gpu_launches"] < sb["gpu_launches"]          # one launch for shadow rays and fold against two or more
    dev.close()


def test_long_cell_lists_equal_the_linear_scan(gpu):
    flat, knot = clustered_scene()
    dev = flat.upload(0)
    rng = np.random.default_rng(5)
    # rays aimed at the knot from all around, and rays that start inside it
    o = knot.mean(axis=0) + rng.normal(size=(40_000, 3)) * 6.0
    t = knot[rng.integers(0, 70, 40_000)] + rng.normal(size=(40_000, 3)) * 0.05
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([np.concatenate([o, d], axis=1),
                           random_rays(rng, 10_000, knot.min(axis=0) - 0.2, knot.max(axis=0) + 0.2)], axis=0)
    og, tg = dev.trace_rays(rays, accel="grid")
    ol, tl = dev.trace_rays(rays, accel="linear")
    assert np.array_equal(og, ol) and np.array_equal(tg, tl)
    assert (og >= 0).mean() > 0.8 and (og[:40_000] < 73).mean() > 0.3          # many of them hit the knot
    cam = sc.pose_camera(0)
    cam.location[:] = (3.3, -13.1, 40.0)
    cam.screen_width, cam.screen_height = 1.0, 0.5625
    a, sa = dev.render(320, 180, 4, fmt="f64", accel="grid", camera=cam)
    b, sb = dev.render(320, 180, 4, fmt="f64", accel="bvh", camera=cam, flags=_lib.FLAG_NO_LIGHT_GRID)
    assert sa["accel_used"] == "grid" and np.array_equal(a, b) and sa["rays"] == sb["rays"]
    dev.close()
