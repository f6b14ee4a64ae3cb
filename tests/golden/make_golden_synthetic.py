"""Generates tests/golden/synthetic_images.json: frames of seeded random-sphere scenes (the C3
recipe of eraytracer_b200/scene.py at 1200 and 3000 spheres) from the CPU oracle (oracle/oracle.c).

Like demo_images.json these are not outputs of the reference (no Erlang here); they pin the
oracle on scenes large enough for the wavefront / cell-grid / direction-grid path and give the
GPU tests a fixture for it.  Before writing, a lattice of pixels is recomputed by the second,
separately written restatement (oracle/pyoracle.py, pure Python over Erlang-shaped tuples) and
must agree with the C oracle bit for bit.

    python tests/golden/make_golden_synthetic.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from eraytracer_b200 import scene as sc  # noqa: E402
from helpers import oracle_scene_from_flat  # noqa: E402
from oracle import orc, pyoracle  # noqa: E402

CASES = (("c3_1200", 1200, 64, 36, 4), ("c3_3000", 3000, 96, 54, 3))


def erlang_scene(flat):
    """FlatScene -> the reference's tagged tuples, list order = `order` (camera first)."""
    cam = flat.camera
    items = []
    for lt in flat.lights:
        items.append((int(lt['order']), ('point_light', ('colour',) + tuple(map(float, lt['diffuse_colour'])),
                                         ('vector',) + tuple(map(float, lt['location'])),
                                         ('colour',) + tuple(map(float, lt['specular_colour'])))))

    def mat(m):
        return ('material', ('colour',) + tuple(map(float, m['colour'])), float(m['specular_power']),
                float(m['shininess']), float(m['reflectivity']))
    for s in flat.spheres:
        items.append((int(s['order']), ('sphere', float(s['radius']), ('vector',) + tuple(map(float, s['center'])),
                                        mat(s['material']))))
    for p in flat.planes:
        items.append((int(p['order']), ('plane', ('vector',) + tuple(map(float, p['normal'])), float(p['distance']),
                                        mat(p['material']))))
    assert len(flat.triangles) == 0
    items.sort(key=lambda t: t[0])
    assert [o for o, _ in items] == list(range(len(items)))
    camera = ('camera', ('vector',) + tuple(cam.location), ('vector',) + tuple(cam.rotation), cam.fov,
              ('screen', cam.screen_width, cam.screen_height))
    return [camera] + [e for _, e in items]


def main():
    out = {"images": {}}
    for name, n, w, h, d in CASES:
        flat = sc.synthetic_scene("c3", n_spheres=n)
        cam, kind, f = oracle_scene_from_flat(flat)
        rgb, rays, tests = orc.render(cam, kind, f, w, h, d)
        rgb = rgb.reshape(h, w, 3)
        # second restatement on a lattice of pixels: bit-for-bit
        escene = erlang_scene(flat)
        checked = 0
        for y in range(1, h, h // 5):
            for x in range(2, w, w // 6):
                c = pyoracle.trace_ray_through_pixel((x / w, y / h), escene, d)
                assert [float(c[1]), float(c[2]), float(c[3])] == rgb[y, x].tolist(), (name, x, y)
                checked += 1
        q = orc.quantise_image(rgb)
        out["images"][name] = {
            "scene": "synthetic_scene('c3', n_spheres=%d)" % n, "n_spheres": n,
            "width": w, "height": h, "depth": d, "rays": rays, "tests": tests,
            "pyoracle_pixels_checked": checked,
            "rgb8": q.reshape(-1).tolist(),
            "f64_sha256": hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest()}
        print(name, "rays", rays, "nonblack", int((q.reshape(-1, 3).sum(axis=1) > 0).sum()), "pyoracle pixels", checked)
    with open(os.path.join(HERE, "synthetic_images.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote synthetic_images.json")


if __name__ == "__main__":
    main()
