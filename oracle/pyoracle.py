"""pyoracle — second, independent restatement of eraytracer's hot path.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c for the rules).  Pure-Python
loops over Erlang-shaped tagged tuples, so it is only usable on small images.
Its purpose is to pin oracle.c: two restatements written separately from
/root/reference/raytracer.erl must agree bit for bit (Python floats are IEEE
doubles, CPython never contracts a*b+c into an FMA, and math.sqrt/pow/tan are
the same glibc calls BEAM's math BIFs make).

PARITY STATUS: like oracle.c — pinned by the reference's run_tests/0
known-answer values only; plane/triangle/shading/whole-image rows are
"parity unpinned" (no golden vectors exist in the reference).

Records are the tuples `scene_test` pins (raytracer.erl:760-801):
  ('vector',X,Y,Z) ('colour',R,G,B) ('ray',Origin,Direction) ('screen',W,H)
  ('camera',Location,Rotation,Fov,Screen)
  ('material',Colour,SpecularPower,Shininess,Reflectivity)
  ('sphere',Radius,Center,Material) ('triangle',V1,V2,V3,Material)
  ('plane',Normal,Distance,Material)
  ('point_light',DiffuseColour,Location,SpecularColour)
"""
import math

NONE = None


# ---- vector primitives (raytracer.erl:524-573) ---------------------------
def vector_add(a, b):
    return ('vector', a[1] + b[1], a[2] + b[2], a[3] + b[3])


def vector_sub(a, b):
    return ('vector', a[1] - b[1], a[2] - b[2], a[3] - b[3])


def vector_square_mag(v):
    _, x, y, z = v
    return x * x + y * y + z * z


def vector_mag(v):
    return math.sqrt(vector_square_mag(v))


def vector_scalar_mult(v, s):
    _, x, y, z = v
    return ('vector', x * s, y * s, z * s)


def vector_component_mult(a, b):
    return ('vector', a[1] * b[1], a[2] * b[2], a[3] * b[3])


def vector_dot_product(a, b):
    return a[1] * b[1] + a[2] * b[2] + a[3] * b[3]


def vector_cross_product(a, b):
    _, a1, a2, a3 = a
    _, b1, b2, b3 = b
    return ('vector', a2 * b3 - a3 * b2, a3 * b1 - a1 * b3, a1 * b2 - a2 * b1)


def vector_normalize(v):
    mag = vector_mag(v)
    if mag == 0:
        return ('vector', 0, 0, 0)
    return vector_scalar_mult(v, 1 / vector_mag(v))


def vector_neg(v):
    return ('vector', -v[1], -v[2], -v[3])


def vector_bounce_off_plane(vector, normal):
    return vector_add(
        vector_scalar_mult(normal, 2 * vector_dot_product(normal, vector_neg(vector))),
        vector)


def vectors_equal(v1, v2, epsilon=0.0001):
    """raytracer.erl:513-521 (test helper)."""
    return all(v1[i] + epsilon >= v2[i] and v1[i] - epsilon <= v2[i] for i in (1, 2, 3))


def point_on_sphere(sphere, p):
    """raytracer.erl:603-607 (test helper)."""
    _, radius, c, _m = sphere
    return 0.001 > abs(((p[1] - c[1]) * (p[1] - c[1]) + (p[2] - c[2]) * (p[2] - c[2])
                        + (p[3] - c[3]) * (p[3] - c[3])) - radius * radius)


def _max0(x):
    # lists:max([0, X])
    return x if x > 0 else 0


# ---- camera (raytracer.erl:483-511) ---------------------------------------
def focal_length(angle, dimension):
    return dimension / (2 * math.tan(angle * (math.pi / 180) / 2))


def point_on_screen(x, y, camera):
    _, location, _rotation, fov, (_s, screen_width, screen_height) = camera
    acc = location
    for vect in (vector_scalar_mult(('vector', 0, 0, 1), focal_length(fov, screen_width)),
                 ('vector', (x - 0.5) * screen_width, 0, 0),
                 ('vector', 0, (y - 0.5) * screen_height, 0)):
        acc = vector_add(vect, acc)
    return acc


def shoot_ray(frm, through):
    return ('ray', frm, vector_normalize(vector_sub(through, frm)))


def ray_through_pixel(x, y, camera):
    return shoot_ray(camera[1], point_on_screen(x, y, camera))


# ---- intersectors ---------------------------------------------------------
def ray_sphere_intersect(ray, sphere):
    """raytracer.erl:364-397."""
    _, (_, x0, y0, z0), (_, xd, yd, zd) = ray
    _, radius, (_, xc, yc, zc), _m = sphere
    epsilon = 0.001
    a = xd * xd + yd * yd + zd * zd
    b = 2 * (xd * (x0 - xc) + yd * (y0 - yc) + zd * (z0 - zc))
    c = (x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc) + (z0 - zc) * (z0 - zc) - radius * radius
    discriminant = b * b - 4 * a * c
    if discriminant >= epsilon:
        t0 = (-b + math.sqrt(discriminant)) / 2
        t1 = (-b - math.sqrt(discriminant)) / 2
        if t0 >= 0 and t1 >= 0:
            distance = min(t0, t1)
            intersection = vector_add(('vector', x0, y0, z0),
                                      vector_scalar_mult(('vector', xd, yd, zd), distance))
            normal = vector_normalize(vector_sub(intersection, ('vector', xc, yc, zc)))
            return (distance, intersection, normal)
    return NONE


def ray_triangle_intersect(ray, triangle):
    """raytracer.erl:402-455."""
    _, origin, direction = ray
    _, v1, v2, v3, _m = triangle
    epsilon = 0.000001
    edge1 = vector_sub(v2, v1)
    edge2 = vector_sub(v3, v1)
    p = vector_cross_product(direction, edge2)
    determinant = vector_dot_product(edge1, p)
    if determinant < epsilon:
        return NONE
    t = vector_sub(origin, v1)
    u = vector_dot_product(t, p)
    if u < 0 or u > determinant:
        return NONE
    q = vector_cross_product(t, edge1)
    v = vector_dot_product(direction, q)
    if v < 0 or u + v > determinant:
        return NONE
    distance = vector_dot_product(edge2, q) / determinant
    intersection = vector_add(origin, vector_scalar_mult(direction, distance))
    normal = vector_normalize(vector_cross_product(v1, v2))
    return (distance, intersection, normal)


def ray_plane_intersect(ray, plane):
    """raytracer.erl:461-480."""
    _, origin, direction = ray
    _, pnormal, pdistance, _m = plane
    epsilon = 0.001
    vd = vector_dot_product(pnormal, direction)
    if vd < 0:
        v0 = -(vector_dot_product(pnormal, origin) + pdistance)
        distance = v0 / vd
        if distance < epsilon:
            return NONE
        intersection = vector_add(origin, vector_scalar_mult(direction, distance))
        return (distance, intersection, pnormal)
    return NONE


def ray_object_intersect(ray, obj):
    """raytracer.erl:349-359."""
    tag = obj[0] if isinstance(obj, tuple) and obj else None
    if tag == 'sphere' and len(obj) == 4:
        return ray_sphere_intersect(ray, obj)
    if tag == 'triangle' and len(obj) == 5:
        return ray_triangle_intersect(ray, obj)
    if tag == 'plane' and len(obj) == 4:
        return ray_plane_intersect(ray, obj)
    return NONE


class Counters:
    def __init__(self):
        self.rays = 0
        self.tests = 0
        self.negative_nearest = 0


def nearest_object_intersecting_ray(ray, scene, counters=None):
    """raytracer.erl:300-346.  Returns (Index, Object, Distance, Hit, Normal) or None;
    Index is this restatement's addition (position in `scene`)."""
    best = None
    if counters is not None:
        counters.rays += 1
    for i, obj in enumerate(scene):
        if counters is not None and isinstance(obj, tuple) and obj and obj[0] in (
                'sphere', 'triangle', 'plane'):
            counters.tests += 1
        res = ray_object_intersect(ray, obj)
        if res is not NONE:
            if best is None or best[2] > res[0]:
                best = (i, obj, res[0], res[1], res[2])
    if counters is not None and best is not None and best[2] < 0:
        counters.negative_nearest += 1
    return best


# ---- shading (raytracer.erl:186-297) --------------------------------------
def _material(obj):
    return obj[-1]


def diffuse_term(obj, light_location, hit_location, hit_normal):
    colour = _material(obj)[1]
    return vector_scalar_mult(
        ('vector', colour[1], colour[2], colour[3]),
        _max0(vector_dot_product(hit_normal,
                                 vector_normalize(vector_sub(light_location, hit_location)))))


def specular_term(eye_vector, light_location, hit_location, hit_normal, specular_power,
                  shininess, specular_colour):
    return vector_scalar_mult(
        ('vector', specular_colour[1], specular_colour[2], specular_colour[3]),
        shininess * math.pow(
            _max0(vector_dot_product(
                vector_normalize(vector_add(
                    vector_normalize(vector_sub(light_location, hit_location)),
                    vector_neg(eye_vector))),
                hit_normal)),
            specular_power))


def shadow_factor(light_location, hit_location, obj_index, scene, counters):
    shadow_ray = ('ray', light_location,
                  vector_normalize(vector_sub(hit_location, light_location)))
    near = nearest_object_intersecting_ray(shadow_ray, scene, counters)
    # term equality restated as index equality; see oracle.c shadow_factor
    return 1 if near is not None and near[0] == obj_index else 0


def lighting_function(ray, obj_index, obj, hit_location, hit_normal, scene, depth, counters):
    final = ('vector', 0, 0, 0)
    _, colour, specular_power, shininess, reflectivity = _material(obj)
    for elem in scene:
        if not (isinstance(elem, tuple) and len(elem) == 4 and elem[0] == 'point_light'):
            continue
        _, light_colour, light_location, specular_colour = elem
        # literal: the reflection is re-traced for every light (erl:216-224)
        child = pixel_colour_from_ray(
            ('ray', hit_location, vector_bounce_off_plane(ray[2], hit_normal)),
            scene, depth - 1, counters)
        reflection = vector_scalar_mult(('vector', child[1], child[2], child[3]), reflectivity)
        contribution = vector_add(
            diffuse_term(obj, light_location, hit_location, hit_normal),
            specular_term(ray[2], light_location, hit_location, hit_normal,
                          specular_power, shininess, specular_colour))
        final = vector_add(
            final,
            vector_add(reflection,
                       vector_scalar_mult(
                           vector_component_mult(
                               ('vector', light_colour[1], light_colour[2], light_colour[3]),
                               contribution),
                           shadow_factor(light_location, hit_location, obj_index, scene,
                                         counters))))
    return final


def pixel_colour_from_ray(ray, scene, depth, counters=None):
    if depth == 0:
        return ('colour', 0, 0, 0)
    near = nearest_object_intersecting_ray(ray, scene, counters)
    if near is None:
        return ('colour', 0, 0, 0)
    i, obj, _dist, hit_location, hit_normal = near
    v = lighting_function(ray, i, obj, hit_location, hit_normal, scene, depth, counters)
    return ('colour', v[1], v[2], v[3])


def trace_ray_through_pixel(xy, scene, depth, counters=None):
    """raytracer.erl:180-184."""
    x, y = xy
    camera, rest = scene[0], scene[1:]
    return pixel_colour_from_ray(ray_through_pixel(x, y, camera), rest, depth, counters)


def raytraced_pixel_list_simple(width, height, scene, depth, counters=None):
    """raytracer.erl:86-99."""
    if width == 0 and height == 0:
        return 'done'
    assert width > 0 and height > 0
    out = []
    for y in range(height):
        for x in range(width):
            c = trace_ray_through_pixel((x / width, y / height), scene, depth, counters)
            out.append((1, (c[1], c[2], c[3])))
    return out


def quantise(c, max_value=255):
    """raytracer.erl:678-680."""
    return min(math.trunc(c * max_value), max_value)


def scene():
    """raytracer.erl:618-665 — the built-in demo scene, values as written there."""
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
