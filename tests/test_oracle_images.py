"""Whole-image pins of the oracle: the two restatements agree bit for bit, the survey's
independent values hold, the committed golden fixture is reproduced."""
import hashlib
import json
import os

import numpy as np
import pytest

from eraytracer_b200 import scene as sc
from helpers import oracle_frame
from oracle import orc, pyoracle as po

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "demo_images.json")))


def demo_arrays():
    s = po.scene()
    kind, f = orc.flatten(s[1:])
    return s, orc.camera_array(s[0]), kind, f


@pytest.mark.parametrize("w,h,d", [(4, 3, 5), (16, 12, 1), (32, 24, 1), (32, 24, 5)])
def test_c_oracle_equals_python_restatement_bit_for_bit(w, h, d):
    s, cam, kind, f = demo_arrays()
    c_rgb, rays_literal, tests = orc.render(cam, kind, f, w, h, d, retrace=True)
    counters = po.Counters()
    py = np.array([p[1] for p in po.raytraced_pixel_list_simple(w, h, s, d, counters)], dtype=np.float64)
    assert np.array_equal(c_rgb, py)
    assert rays_literal == counters.rays and tests == counters.tests


def test_memoised_reflection_is_bit_identical_to_literal_retrace():
    """The reference re-traces the reflection once per light (erl:216-224); tracing it once
    gives the same doubles with a third of the rays."""
    _s, cam, kind, f = demo_arrays()
    lit, rays_lit, _ = orc.render(cam, kind, f, 32, 24, 5, retrace=True)
    memo, rays_memo, _ = orc.render(cam, kind, f, 32, 24, 5, retrace=False)
    assert np.array_equal(lit, memo)
    assert rays_lit == 10824 and rays_memo == 3590


def test_survey_appendix_a_values():
    a = GOLDEN["appendix_a"]
    _s, cam, kind, f = demo_arrays()
    assert orc.lib().orc_focal_length(90, 4) == a["focal_length_90_4"]
    for key, d in (("32x24_d1", 1), ("32x24_d5", 5)):
        rgb, rays, tests = orc.render(cam, kind, f, 32, 24, d, retrace=True)
        q = orc.quantise_image(rgb).reshape(24, 32, 3)
        assert int((q.sum(axis=2) > 0).sum()) == a[key]["nonblack"]
        assert int((q.max(axis=2) >= 255).sum()) == a[key]["saturated"]
        assert q.reshape(-1, 3).sum(axis=0).tolist() == a[key]["sums"]
        for xy, want in a[key]["pixels"].items():
            x, y = map(int, xy.split(","))
            assert q[y, x].tolist() == want
    rgb, _, _ = orc.render(cam, kind, f, 4, 3, 5)
    assert orc.quantise_image(rgb).tolist() == a["4x3_d5"]


@pytest.mark.parametrize("key", sorted(GOLDEN["images"]))
def test_golden_fixture_is_reproduced(key):
    g = GOLDEN["images"][key]
    _s, cam, kind, f = demo_arrays()
    rgb, rays, tests = orc.render(cam, kind, f, g["width"], g["height"], g["depth"])
    assert orc.quantise_image(rgb).reshape(-1).tolist() == g["rgb8"]
    assert hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest() == g["f64_sha256"]
    assert rays == g["rays"]


def test_nearest_hit_map_of_survey_appendix_a():
    """Rows 6-23 of the 32x24 nearest-object map ('.', S, T, P)."""
    expected = """
.........SSSSS.......T..........
........SSSSSSS.....SSSSS.......
........SSSSSSSS..SSSSSSSSS.....
.......SSSSSSSSS.SSSSSSSSSSS....
.......SSSSSSSSS.SSSSSSSSSSS....
......SSSSSSSSSS.SSSSSSSSSSS....
....SSSSSSSSSSS..SSSSSSSSSSS....
PPPSSSSSSSSSSSPPPSSSSSSSSSSSPPPP
PPSSSSSSSSSSSSSPPTSSSSSSSSSSPPPP
PPSSSSSSSSSSSSSPPTTSSSSSSSSSPPPP
PSSSSSSSSSSSSSSPTTTSSSSSSSSPPPPP
PSSSSSSSSSSSSSSTTTTTSSSSSPPPPPPP
PSSSSSSSSSSSSSSTTTTTTTPPPPPPPPPP
PSSSSSSSSSSSSSTTTTTTTPPPPPPPPPPP
PPSSSSSSSSSSSSTTTTTPPPPPPPPPPPPP
PPSSSSSSSSSSPTTTTPPPPPPPPPPPPPPP
PPPSSSSSSSPPPTTPPPPPPPPPPPPPPPPP
PPPPPPPPPPPPTPPPPPPPPPPPPPPPPPPP""".strip().split("\n")
    s = po.scene()
    rows = []
    for y in range(24):
        row = ""
        for x in range(32):
            near = po.nearest_object_intersecting_ray(po.ray_through_pixel(x / 32, y / 24, s[0]), s[1:])
            row += "." if near is None else {"sphere": "S", "triangle": "T", "plane": "P"}[near[1][0]]
        rows.append(row)
    assert rows[:6] == ["." * 32] * 6
    assert rows[6:] == expected


def test_oracle_edge_cases():
    s, cam, kind, f = demo_arrays()
    # depth 0 is black without tracing (erl:186-187)
    rgb, rays, _ = orc.render(cam, kind, f, 8, 6, 0)
    assert not rgb.any() and rays == 0
    # no lights: the fold at erl:211-252 starts and ends at (0,0,0)
    nolights = [e for e in s[1:] if e[0] != 'point_light']
    k2, f2 = orc.flatten(nolights)
    rgb, rays, _ = orc.render(cam, k2, f2, 8, 6, 5)
    assert not rgb.any() and rays == 48
    # camera only: every ray misses (erl:303-305)
    k3, f3 = orc.flatten([])
    rgb, _, _ = orc.render(cam, k3, f3, 4, 4, 3)
    assert not rgb.any()
    # unknown list elements are skipped (erl:357-358, 248-249)
    junk = list(s[1:])
    junk.insert(2, ('fog', 40))
    junk.append('atom')
    k4, f4 = orc.flatten(junk)
    a, _, _ = orc.render(cam, k4, f4, 16, 12, 3)
    b, _, _ = orc.render(cam, kind, f, 16, 12, 3)
    assert np.array_equal(a, b)


def test_quantisation_rule():
    """erl:678-680: trunc toward zero, upper clamp only."""
    L = orc.lib()
    assert L.orc_quantise(0.999999, 255) == 254
    assert L.orc_quantise(1.0, 255) == 255
    assert L.orc_quantise(2.3, 255) == 255
    assert L.orc_quantise(0.0, 255) == 0
    assert L.orc_quantise(-0.5, 255) == -127
    assert po.quantise(-0.5) == -127 and po.quantise(0.5) == 127


def test_synthetic_scene_oracle_small_vs_python():
    """A small synthetic scene through both restatements (3 lights, spheres, plane)."""
    flat = sc.synthetic_scene("c3", n_spheres=40)
    ref, _, _ = oracle_frame(flat, 24, 14, 3)
    # the same scene as Erlang-shaped tuples for the Python restatement
    def col(v):
        return ('colour', float(v[0]), float(v[1]), float(v[2]))
    def vec(v):
        return ('vector', float(v[0]), float(v[1]), float(v[2]))
    def mat(m):
        return ('material', col(m['colour']), float(m['specular_power']), float(m['shininess']),
                float(m['reflectivity']))
    c = flat.camera
    scene = [('camera', vec(c.location), vec(c.rotation), c.fov, ('screen', c.screen_width, c.screen_height))]
    scene += [('point_light', col(l['diffuse_colour']), vec(l['location']), col(l['specular_colour']))
              for l in flat.lights]
    scene += [('sphere', float(s['radius']), vec(s['center']), mat(s['material'])) for s in flat.spheres]
    scene += [('plane', vec(p['normal']), float(p['distance']), mat(p['material'])) for p in flat.planes]
    py = np.array([p[1] for p in po.raytraced_pixel_list_simple(24, 14, scene, 3)]).reshape(14, 24, 3)
    assert np.array_equal(ref, py)
    assert ref.any()


SYNTHETIC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synthetic_images.json")))


@pytest.mark.parametrize("key", sorted(SYNTHETIC["images"]))
def test_synthetic_golden_fixture_is_reproduced(key):
    """tests/golden/synthetic_images.json (random-sphere scenes, made by make_golden_synthetic.py, which
    also checks a lattice of pixels against the pure-Python restatement bit for bit)."""
    from eraytracer_b200 import scene as sc
    from helpers import oracle_scene_from_flat
    g = SYNTHETIC["images"][key]
    flat = sc.synthetic_scene("c3", n_spheres=g["n_spheres"])
    cam, kind, f = oracle_scene_from_flat(flat)
    rgb, rays, tests = orc.render(cam, kind, f, g["width"], g["height"], g["depth"])
    assert orc.quantise_image(rgb).reshape(-1).tolist() == g["rgb8"]
    assert hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest() == g["f64_sha256"]
    assert rays == g["rays"] and tests == g["tests"]
