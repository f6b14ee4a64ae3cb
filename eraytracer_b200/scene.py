"""Scene model of the host side: Erlang-shaped records -> flat tables for the C ABI.

Records are the tagged tuples of raytracer.erl:72-81 as `scene_test` pins them
(raytracer.erl:760-801), e.g.
    ('sphere', 4, ('vector', 4, 0, 10), ('material', ('colour', 0, 0.5, 1), 20, 1, 0.1))
A scene is a list with the camera first (raytracer.erl:617-619).  Every numeric
field may be an int or a float (Erlang scene literals mix both, erl:619-664).
"""
import math

import numpy as np

from . import _lib


def demo_scene():
    """The built-in scene of scene/0, raytracer.erl:618-665, values as written there."""
    return [
        ('camera', ('vector', 0, 0, -2), ('vector', 0, 0, 0), 90, ('screen', 4, 3)),
        ('point_light', ('colour', 1, 1, 0.5), ('vector', 5, -2, 0), ('colour', 1, 1, 1)),
        ('point_light', ('colour', 1, 0, 0.5), ('vector', -10, 0, 7), ('colour', 1, 0, 0.5)),
        ('sphere', 4, ('vector', 4, 0, 10),
         ('material', ('colour', 0, 0.5, 1), 20, 1, 0.1)),
        ('sphere', 4, ('vector', -5, 3, 9),
         ('material', ('colour', 1, 0.5, 0), 4, 0.25, 0.5)),
        ('sphere', 4, ('vector', -4.5, -2.5, 14),
         ('material', ('colour', 0.5, 1, 0), 20, 0.25, 0.7)),
        ('triangle', ('vector', -2, 5, 5), ('vector', 4, 5, 10), ('vector', 4, -5, 10),
         ('material', ('colour', 1, 0.5, 0), 4, 0.25, 0.5)),
        ('plane', ('vector', 0, -1, 0), 5,
         ('material', ('colour', 1, 1, 1), 1, 0, 0.01)),
    ]


def _num(x):
    # Erlang numbers only; the atom 'undefined' (hand-built records, erl:1016-1020) is badarg
    if isinstance(x, bool) or not isinstance(x, (int, float, np.integer, np.floating)):
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "not a number in a scene record: %r" % (x,))
    return float(x)


def _v3(rec, tag):
    if not (isinstance(rec, tuple) and len(rec) == 4 and rec[0] == tag):
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "expected a #%s{} record, got %r" % (tag, rec))
    return (_num(rec[1]), _num(rec[2]), _num(rec[3]))


def _material(rec):
    if not (isinstance(rec, tuple) and len(rec) == 5 and rec[0] == 'material'):
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "expected a #material{} record, got %r" % (rec,))
    return (_v3(rec[1], 'colour'), _num(rec[2]), _num(rec[3]), _num(rec[4]))


def camera_struct(rec):
    """#camera{location, rotation, fov, screen=#screen{width,height}} -> _lib.Camera."""
    if isinstance(rec, _lib.Camera):
        return rec
    if not (isinstance(rec, tuple) and len(rec) == 5 and rec[0] == 'camera'):
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "the first scene element must be a #camera{} record")
    scr = rec[4]
    if not (isinstance(scr, tuple) and len(scr) == 3 and scr[0] == 'screen'):
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "expected a #screen{} record in the camera")
    cam = _lib.Camera()
    cam.location[:] = _v3(rec[1], 'vector')
    cam.rotation[:] = _v3(rec[2], 'vector')
    cam.fov = _num(rec[3])
    cam.screen_width = _num(scr[1])
    cam.screen_height = _num(scr[2])
    return cam


class FlatScene:
    """The four element tables plus the camera, ready for ert_scene_create."""

    def __init__(self, camera, lights, spheres, triangles, planes):
        self.camera = camera
        self.lights = lights
        self.spheres = spheres
        self.triangles = triangles
        self.planes = planes

    def upload(self, device=0):
        return _lib.Scene.create(self.camera, self.lights, self.spheres, self.triangles,
                                 self.planes, device=device)

    @property
    def n_elements(self):
        return len(self.lights) + len(self.spheres) + len(self.triangles) + len(self.planes)


def _set_mat(row, mat):
    colour, sp, sh, refl = mat
    row['material']['colour'] = colour
    row['material']['specular_power'] = sp
    row['material']['shininess'] = sh
    row['material']['reflectivity'] = refl


def flatten(scene):
    """[Camera | Rest] (raytracer.erl:180) -> FlatScene.  List positions after the camera
    become `order`; elements that are no known record are skipped like erl:357-358."""
    if not isinstance(scene, (list, tuple)) or len(scene) == 0:
        raise _lib.BadArg(_lib.ERT_ERR_BADARG, "scene must be a non-empty list, camera first")
    camera = camera_struct(scene[0])
    lights, spheres, tris, planes = [], [], [], []
    for order, e in enumerate(scene[1:]):
        tag = e[0] if isinstance(e, tuple) and len(e) else None
        if tag == 'point_light' and len(e) == 4:
            lights.append((order, _v3(e[1], 'colour'), _v3(e[2], 'vector'), _v3(e[3], 'colour')))
        elif tag == 'sphere' and len(e) == 4:
            spheres.append((order, _num(e[1]), _v3(e[2], 'vector'), _material(e[3])))
        elif tag == 'triangle' and len(e) == 5:
            tris.append((order, _v3(e[1], 'vector'), _v3(e[2], 'vector'), _v3(e[3], 'vector'),
                         _material(e[4])))
        elif tag == 'plane' and len(e) == 4:
            planes.append((order, _v3(e[1], 'vector'), _num(e[2]), _material(e[3])))
        # anything else: not an object, not a light — ignored by the reference
    lt = np.zeros(len(lights), dtype=_lib.LIGHT_DT)
    for i, (o, dc, loc, sc) in enumerate(lights):
        lt[i]['diffuse_colour'] = dc
        lt[i]['location'] = loc
        lt[i]['specular_colour'] = sc
        lt[i]['order'] = o
    st = np.zeros(len(spheres), dtype=_lib.SPHERE_DT)
    for i, (o, r, c, m) in enumerate(spheres):
        st[i]['radius'] = r
        st[i]['center'] = c
        _set_mat(st[i], m)
        st[i]['order'] = o
    tt = np.zeros(len(tris), dtype=_lib.TRIANGLE_DT)
    for i, (o, v1, v2, v3, m) in enumerate(tris):
        tt[i]['v1'], tt[i]['v2'], tt[i]['v3'] = v1, v2, v3
        _set_mat(tt[i], m)
        tt[i]['order'] = o
    pt = np.zeros(len(planes), dtype=_lib.PLANE_DT)
    for i, (o, n, d, m) in enumerate(planes):
        pt[i]['normal'] = n
        pt[i]['distance'] = d
        _set_mat(pt[i], m)
        pt[i]['order'] = o
    return FlatScene(camera, lt, st, tt, pt)


# ---------------------------------------------------------------------------
# Synthetic scenes: "randomly generated scene
# (not done)" is on the reference's own to-do list (raytracer.erl:35).
# ---------------------------------------------------------------------------
_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64_uniform(seed, count):
    """count doubles in [0,1): u_k = (splitmix64 output k >> 11) * 2^-53."""
    k = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over='ignore'):
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def _f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


SYNTH_CONFIGS = {
    # name: (n_spheres, (xlo,xhi), (ylo,yhi), (zlo,zhi), (rlo,rhi), seed)
    "c3": (10_000, (-40.0, 40.0), (-30.0, 4.0), (5.0, 85.0), (0.2, 0.8), 0xE7A9C0DE00000003),
    "c4": (1_000_000, (-200.0, 200.0), (-150.0, 4.0), (5.0, 405.0), (0.2, 1.0), 0xE7A9C0DE00000004),
}


def synthetic_scene(name="c3", n_spheres=None, seed=None):
    """Random-sphere scene of SURVEY §8(d): camera (0,0,-2) fov 90 screen 4x2.25, the demo
    floor plane, three point lights, n spheres.  Every value is rounded to float32 so that a
    double-precision CPU evaluation and the GPU start from identical inputs.  List order: lights, spheres, plane."""
    n0, xr, yr, zr, rr, seed0 = SYNTH_CONFIGS[name]
    n = n0 if n_spheres is None else int(n_spheres)
    seed = seed0 if seed is None else seed
    u = splitmix64_uniform(seed, n * 10).reshape(n, 10)
    cam = _lib.Camera()
    cam.location[:] = (0.0, 0.0, -2.0)
    cam.rotation[:] = (0.0, 0.0, 0.0)
    cam.fov = 90.0
    cam.screen_width, cam.screen_height = 4.0, 2.25

    lights = np.zeros(3, dtype=_lib.LIGHT_DT)
    lights['location'] = [(5, -20, 0), (-30, -10, 20), (20, -40, 60)]
    lights['diffuse_colour'] = [(1, 1, 0.5), (1, 0, 0.5), (1, 1, 1)]
    lights['specular_colour'] = [(1, 1, 1), (1, 0, 0.5), (1, 1, 1)]
    lights['order'] = [0, 1, 2]

    sp = np.zeros(n, dtype=_lib.SPHERE_DT)
    sp['center'][:, 0] = _f32(xr[0] + u[:, 0] * (xr[1] - xr[0]))
    sp['center'][:, 1] = _f32(yr[0] + u[:, 1] * (yr[1] - yr[0]))
    sp['center'][:, 2] = _f32(zr[0] + u[:, 2] * (zr[1] - zr[0]))
    sp['radius'] = _f32(rr[0] + u[:, 3] * (rr[1] - rr[0]))
    sp['material']['colour'][:, 0] = _f32(u[:, 4])
    sp['material']['colour'][:, 1] = _f32(u[:, 5])
    sp['material']['colour'][:, 2] = _f32(u[:, 6])
    powers = np.array([1.0, 4.0, 20.0, 50.0])
    sp['material']['specular_power'] = powers[np.minimum((u[:, 7] * 4).astype(np.int64), 3)]
    sp['material']['shininess'] = _f32(u[:, 8])
    sp['material']['reflectivity'] = _f32(u[:, 9] * 0.7)
    sp['order'] = np.arange(3, 3 + n, dtype=np.int32)

    planes = np.zeros(1, dtype=_lib.PLANE_DT)
    planes['normal'] = [(0, -1, 0)]
    planes['distance'] = 5
    planes['material']['colour'] = [(1, 1, 1)]
    planes['material']['specular_power'] = 1
    planes['material']['shininess'] = 0
    planes['material']['reflectivity'] = _f32(0.01)
    planes['order'] = 3 + n
    return FlatScene(cam, lights, sp, np.zeros(0, dtype=_lib.TRIANGLE_DT), planes)


def pose_camera(k, n_poses=64):
    """Camera of pose k for the C5 batch: only the location varies, because
    the reference ignores rotation (raytracer.erl:487)."""
    cam = _lib.Camera()
    a = 2.0 * math.pi * k / n_poses
    cam.location[:] = (4.0 * math.sin(a), -1.0 + math.cos(a) / 2.0, -2.0 - k / 16.0)
    cam.rotation[:] = (0.0, 0.0, 0.0)
    cam.fov = 90.0
    cam.screen_width, cam.screen_height = 4.0, 3.0
    return cam
