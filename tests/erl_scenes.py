"""Scenes of tests/golden/erl_reference.json: used by the generator (which hands them to the reference's own
functions through oracle/erlref.py) and by the tests (which hand the same records to the oracle and the GPU).

Records are the tagged tuples of raytracer.erl:72-81 with Python str tags (eraytracer_b200/scene.py);
`scene_terms` turns them into the evaluator's terms (atoms).
"""
import numpy as np

from eraytracer_b200 import scene as sc


def _mat(c, sp, sh, refl):
    return ('material', ('colour',) + tuple(c), sp, sh, refl)


def stress_scene():
    """Reaches what no reference test pins: two planes (one with an un-normalised normal, erl:461-480 never
    normalises it, so its reflection rays are not unit vectors), triangles incl. one lying IN a plane (exact
    distance ties, decided by list position, erl:319), two identical spheres with different materials, a huge
    and a tiny sphere, three lights, and an element that is no record of the module (skipped, erl:357-358)."""
    return [
        ('camera', ('vector', 0.5, -1, -3), ('vector', 0, 0, 0), 70, ('screen', 4, 3)),
        ('point_light', ('colour', 1, 0.9, 0.6), ('vector', 6, -8, -2), ('colour', 1, 1, 1)),
        ('sphere', 2, ('vector', 0, 0, 8), _mat((0.2, 0.6, 1), 20, 1, 0.3)),
        ('fog', 40, 'not_a_record_of_the_module'),
        ('sphere', 2, ('vector', 0, 0, 8), _mat((1, 0, 0), 4, 0.5, 0.9)),
        ('point_light', ('colour', 0.5, 0, 1), ('vector', -7, -3, 4), ('colour', 0.25, 0.5, 1)),
        ('triangle', ('vector', -6, 4, 30), ('vector', 6, 4, 30), ('vector', 0, -9, 30), _mat((1, 1, 0), 50, 0.75, 0.5)),
        ('plane', ('vector', 0, 0, -1), 30, _mat((0.3, 0.8, 0.4), 1, 0, 0)),
        ('sphere', 1.5, ('vector', 3.5, 1, 6), _mat((1, 0.5, 0), 4, 0.25, 0.6)),
        ('sphere', 1, ('vector', -3, -2, 5), _mat((0.9, 0.9, 0.9), 1, 0.1, 0)),
        ('point_light', ('colour', 1, 1, 1), ('vector', 0, -12, 15), ('colour', 1, 0.5, 0.5)),
        ('sphere', 50, ('vector', 0, 60, 40), _mat((0.4, 0.4, 0.5), 20, 0.5, 0.2)),
        ('sphere', 0.05, ('vector', 1, -1, 2), _mat((1, 1, 1), 50, 1, 0.7)),
        ('triangle', ('vector', -5, 3, 7), ('vector', -1, 3, 12), ('vector', -1, -4, 12), _mat((0, 1, 1), 4, 0.25, 0.5)),
        ('plane', ('vector', 0, -2, 0), 10, _mat((1, 1, 1), 1, 0, 0.4)),
    ]


def pose_scene(k):
    s = sc.demo_scene()
    cam = sc.pose_camera(k)
    s[0] = ('camera', ('vector',) + tuple(float(x) for x in cam.location), ('vector', 0, 0, 0), 90, ('screen', 4, 3))
    return s


def flat_to_tuples(flat):
    """FlatScene -> list of records in list order (camera first)."""
    cam = flat.camera
    out = [('camera', ('vector',) + tuple(float(x) for x in cam.location), ('vector',) + tuple(float(x) for x in cam.rotation),
            float(cam.fov), ('screen', float(cam.screen_width), float(cam.screen_height)))]
    elems = []

    def mat(row):
        m = row['material']
        return _mat([float(x) for x in m['colour']], float(m['specular_power']), float(m['shininess']), float(m['reflectivity']))

    for r in flat.lights:
        elems.append((int(r['order']), ('point_light', ('colour',) + tuple(float(x) for x in r['diffuse_colour']),
                                        ('vector',) + tuple(float(x) for x in r['location']),
                                        ('colour',) + tuple(float(x) for x in r['specular_colour']))))
    for r in flat.spheres:
        elems.append((int(r['order']), ('sphere', float(r['radius']), ('vector',) + tuple(float(x) for x in r['center']), mat(r))))
    for r in flat.triangles:
        elems.append((int(r['order']), ('triangle', ('vector',) + tuple(float(x) for x in r['v1']),
                                        ('vector',) + tuple(float(x) for x in r['v2']),
                                        ('vector',) + tuple(float(x) for x in r['v3']), mat(r))))
    for r in flat.planes:
        elems.append((int(r['order']), ('plane', ('vector',) + tuple(float(x) for x in r['normal']), float(r['distance']), mat(r))))
    elems.sort(key=lambda e: e[0])
    assert [e[0] for e in elems] == list(range(len(elems)))
    return out + [e[1] for e in elems]


def scene_records(name):
    """The scene as Python records (str tags)."""
    if name == "demo":
        return sc.demo_scene()
    if name.startswith("pose"):
        return pose_scene(int(name[4:]))
    if name == "stress":
        return stress_scene()
    if name == "mini_c3":
        return flat_to_tuples(sc.synthetic_scene("c3", n_spheres=40))
    raise KeyError(name)


SCENES = {
    # name: images (w, h, depth) rendered by the reference's raytraced_pixel_list_simple/4, rays traced by its
    # nearest_object_intersecting_ray/2
    "demo": {"images": [(32, 24, 1), (16, 12, 5), (32, 24, 5)], "rays": 200},      # 32x24 depth 1 is config C1 (run.sh)
    "pose5": {"images": [(16, 12, 2)]},
    "pose37": {"images": [(16, 12, 2)]},
    "stress": {"images": [(24, 18, 4)], "rays": 300},
    "mini_c3": {"images": [(16, 9, 3)], "rays": 200},
}


def to_term(m_mod, rec):
    """Python record -> evaluator term (str tags become atoms, recursively)."""
    from oracle import erlref
    if isinstance(rec, tuple):
        return tuple(to_term(m_mod, x) for x in rec)
    if isinstance(rec, str):
        return erlref.Atom(rec)
    if isinstance(rec, (np.floating,)):
        return float(rec)
    if isinstance(rec, (np.integer,)):
        return int(rec)
    return rec


def scene_terms(m_mod, name):
    if name == "demo":
        return m_mod.call("scene")             # the reference's own scene/0 (erl:618-665)
    return [to_term(m_mod, e) for e in scene_records(name)]


def ray_batch(name, n):
    """Seeded rays: from around the camera into the scene, from the lights, and a few axis-parallel ones."""
    recs = scene_records(name)
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)) * 7919)
    cam = np.array(recs[0][1][1:], dtype=np.float64)
    lights = [np.array(e[2][1:], dtype=np.float64) for e in recs[1:] if e[0] == 'point_light']
    rays = np.zeros((n, 6))
    for k in range(n):
        if k % 3 == 2 and lights:
            o = lights[(k // 3) % len(lights)]
            d = rng.uniform(-1, 1, 3)
        else:
            o = cam + (rng.uniform(-0.5, 0.5, 3) if k % 2 else 0.0)
            d = np.array([rng.uniform(-1.2, 1.2), rng.uniform(-0.9, 0.9), 1.0])
        if k % 17 == 0:
            d = np.array([0.0, 0.0, 1.0])
        if k % 5:
            d = d / np.sqrt((d * d).sum())          # the rest stay un-normalised on purpose
        rays[k, :3] = o
        rays[k, 3:] = d
    return rays
