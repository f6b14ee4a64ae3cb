"""The reference's own run_tests/0 (raytracer.erl:736-1133), restated against BOTH oracle
restatements (oracle/oracle.c through ctypes, oracle/pyoracle.py).  These known-answer
values are the only golden vectors the reference holds for this path."""
import ctypes
import math

import numpy as np
import pytest

from oracle import orc, pyoracle as po


def V(x, y, z):
    return ('vector', x, y, z)


class COracle:
    """oracle.c behind the same function names pyoracle uses."""

    def __init__(self):
        self.L = orc.lib()

    @staticmethod
    def _a(v):
        return orc.vec(v)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def _vv(self, fn, a, b):
        a, b, o = self._a(a), self._a(b), np.zeros(3)
        fn(self._p(a), self._p(b), self._p(o))
        return ('vector',) + tuple(o.tolist())

    def vector_add(self, a, b): return self._vv(self.L.orc_vector_add, a, b)
    def vector_sub(self, a, b): return self._vv(self.L.orc_vector_sub, a, b)
    def vector_cross_product(self, a, b): return self._vv(self.L.orc_vector_cross_product, a, b)
    def vector_bounce_off_plane(self, a, b): return self._vv(self.L.orc_vector_bounce_off_plane, a, b)

    def _v(self, fn, a):
        a, o = self._a(a), np.zeros(3)
        fn(self._p(a), self._p(o))
        return ('vector',) + tuple(o.tolist())

    def vector_normalize(self, a): return self._v(self.L.orc_vector_normalize, a)
    def vector_neg(self, a): return self._v(self.L.orc_vector_neg, a)

    def vector_scalar_mult(self, a, s):
        a, o = self._a(a), np.zeros(3)
        self.L.orc_vector_scalar_mult(self._p(a), float(s), self._p(o))
        return ('vector',) + tuple(o.tolist())

    def vector_square_mag(self, a): return self.L.orc_vector_square_mag(self._p(self._a(a)))
    def vector_mag(self, a): return self.L.orc_vector_mag(self._p(self._a(a)))

    def vector_dot_product(self, a, b):
        a, b = self._a(a), self._a(b)
        return self.L.orc_vector_dot_product(self._p(a), self._p(b))

    def focal_length(self, angle, dim): return self.L.orc_focal_length(float(angle), float(dim))

    def point_on_screen(self, x, y, camera):
        cam, o = orc.camera_array(camera), np.zeros(3)
        self.L.orc_point_on_screen(float(x), float(y), self._p(cam), self._p(o))
        return ('vector',) + tuple(o.tolist())

    def shoot_ray(self, frm, through):
        a, b, o = self._a(frm), self._a(through), np.zeros(6)
        self.L.orc_shoot_ray(self._p(a), self._p(b), self._p(o))
        return ('ray', ('vector',) + tuple(o[:3].tolist()), ('vector',) + tuple(o[3:].tolist()))

    @staticmethod
    def _ray6(ray):
        return np.array(list(ray[1][1:]) + list(ray[2][1:]), dtype=np.float64)

    def ray_sphere_intersect(self, ray, sphere):
        kind, f = orc.flatten([sphere])
        out = np.zeros(7)
        r6 = self._ray6(ray)
        hit = self.L.orc_ray_object_intersect(self._p(r6), int(kind[0]), self._p(f), self._p(out))
        if not hit:
            return None
        return (out[0], ('vector',) + tuple(out[1:4].tolist()), ('vector',) + tuple(out[4:7].tolist()))

    def nearest_object_intersecting_ray(self, ray, scene):
        kind, f = orc.flatten(scene)
        out = np.zeros(7)
        r6 = self._ray6(ray)
        i = self.L.orc_nearest_object_intersecting_ray(
            self._p(r6), len(kind), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            self._p(f), self._p(out))
        if i < 0:
            return None
        return (i, scene[i], out[0], ('vector',) + tuple(out[1:4].tolist()),
                ('vector',) + tuple(out[4:7].tolist()))


@pytest.fixture(params=["c", "py"])
def o(request):
    return COracle() if request.param == "c" else po


def eq(a, b, eps=0.0001):
    return po.vectors_equal(a, b, eps)


def test_scene_test():
    """scene_test erl:760-801: the record -> tuple layout (the NIF's decode contract)."""
    assert po.scene() == [
        ('camera', ('vector', 0, 0, -2), ('vector', 0, 0, 0), 90, ('screen', 4, 3)),
        ('point_light', ('colour', 1, 1, 0.5), ('vector', 5, -2, 0), ('colour', 1, 1, 1)),
        ('point_light', ('colour', 1, 0, 0.5), ('vector', -10, 0, 7), ('colour', 1, 0, 0.5)),
        ('sphere', 4, ('vector', 4, 0, 10), ('material', ('colour', 0, 0.5, 1), 20, 1, 0.1)),
        ('sphere', 4, ('vector', -5, 3, 9), ('material', ('colour', 1, 0.5, 0), 4, 0.25, 0.5)),
        ('sphere', 4, ('vector', -4.5, -2.5, 14), ('material', ('colour', 0.5, 1, 0), 20, 0.25, 0.7)),
        ('triangle', ('vector', -2, 5, 5), ('vector', 4, 5, 10), ('vector', 4, -5, 10),
         ('material', ('colour', 1, 0.5, 0), 4, 0.25, 0.5)),
        ('plane', ('vector', 0, -1, 0), 5, ('material', ('colour', 1, 1, 1), 1, 0, 0.01)),
    ]


def test_vector_equality():
    """vector_equality_test erl:828-844."""
    v1, v2 = V(0, 0, 0), V(1234, -234, 0)
    assert eq(v1, v1) and eq(v2, v2) and not eq(v1, v2) and not eq(v2, v1)
    assert eq(V(0.0983, 0.0214, 0.12342), V(0.0984, 0.0213, 0.12341), 0.0001)
    assert eq(V(10 / 3, -10 / 6, 8 / 7), V(3.3, -1.6, 1.1), 0.1)


def test_vector_addition(o):
    """vector_addition_test erl:847-866."""
    v0 = o.vector_add(V(3, 7, -3), V(0, -24, 123))
    assert v0[1:] == (3, -17, 120)
    v1 = V(5, 0, 984)
    v2 = o.vector_add(v1, v1)
    assert v2[1:] == (10, 0, 1968)
    v3 = V(908, -98, 234)
    assert eq(v3, o.vector_add(v3, V(0, 0, 0)))


def test_vector_subtraction(o):
    """vector_subtraction_test erl:868-883."""
    v1, v2, v3, v4 = V(0, 0, 0), V(8390, -2098, 939), V(1, 1, 1), V(-1, -1, -1)
    assert eq(v1, o.vector_sub(v1, v1))
    assert eq(v3, o.vector_sub(v3, v1))
    assert not eq(v3, o.vector_sub(v1, v3))
    assert eq(v4, o.vector_sub(v4, v1))
    assert not eq(v4, o.vector_sub(v1, v4))
    assert eq(o.vector_add(v2, v4), o.vector_sub(v2, v3))


def test_vector_square_mag_and_mag(o):
    """vector_square_mag_test erl:885-895, vector_mag_test erl:897-907."""
    assert o.vector_square_mag(V(0, 0, 0)) == 0
    assert o.vector_square_mag(V(1, 1, 1)) == 3
    assert o.vector_square_mag(V(3, -4, 0)) == 25
    assert o.vector_mag(V(0, 0, 0)) == 0
    assert o.vector_mag(V(1, 1, 1)) == math.sqrt(3)
    assert o.vector_mag(V(3, -4, 0)) == 5


def test_vector_scalar_multiplication(o):
    """vector_scalar_multiplication_test erl:909-923."""
    z, ones, v3 = V(0, 0, 0), V(1, 1, 1), V(3, -4, 0)
    assert eq(z, o.vector_scalar_mult(z, 45))
    assert eq(z, o.vector_scalar_mult(z, -13))
    assert eq(z, o.vector_scalar_mult(v3, 0))
    assert eq(V(4, 4, 4), o.vector_scalar_mult(ones, 4))
    assert eq(v3, o.vector_scalar_mult(v3, 1))
    assert not eq(v3, o.vector_scalar_mult(v3, -3))


def test_vector_dot_product(o):
    """vector_dot_product_test erl:925-939."""
    v1, v2 = V(1, 3, -5), V(4, -2, -1)
    assert o.vector_dot_product(v1, v2) == 3
    assert o.vector_dot_product(v2, v2) == o.vector_square_mag(v2)
    assert o.vector_dot_product(V(0, 0, 0), v1) == 0
    assert o.vector_dot_product(V(1, 0, 0), V(0, 1, 0)) == 0


def test_vector_cross_product(o):
    """vector_cross_product_test erl:941-976 (incl. distributivity and the Jacobi identity)."""
    z, x, y, zz = V(0, 0, 0), V(1, 0, 0), V(0, 1, 0), V(0, 0, 1)
    v5, v6, v7, v8, v9 = V(1, 2, 3), V(4, 5, 6), V(-3, 6, -3), V(-1, 0, 0), V(-9, 8, 433)
    cp = o.vector_cross_product
    assert eq(z, cp(x, x))
    assert eq(z, cp(x, v8))
    assert eq(x, cp(y, zz))
    assert eq(v7, cp(v5, v6))
    assert eq(cp(v7, o.vector_add(v8, v9)), o.vector_add(cp(v7, v8), cp(v7, v9)))
    assert eq(z, o.vector_add(o.vector_add(cp(v7, cp(v8, v9)), cp(v8, cp(v9, v7))),
                              cp(v9, cp(v7, v8))))


def test_vector_normalization(o):
    """vector_normalization_test erl:978-990 (the zero vector maps to zero)."""
    z, x = V(0, 0, 0), V(1, 0, 0)
    assert eq(z, o.vector_normalize(z))
    assert eq(x, o.vector_normalize(x))
    assert eq(x, o.vector_normalize(V(5, 0, 0)))
    assert eq(x, o.vector_normalize(o.vector_scalar_mult(x, 324)))


def test_vector_negation(o):
    """vector_negation_test erl:992-1000."""
    v = V(4, -5, 6)
    assert eq(V(0, 0, 0), o.vector_neg(V(0, 0, 0)))
    assert eq(v, o.vector_neg(o.vector_neg(v)))


def test_ray_shooting(o):
    """ray_shooting_test erl:1002-1011."""
    assert eq(o.shoot_ray(V(0, 0, 0), V(1, 0, 0))[2], V(1, 0, 0))


def test_ray_sphere_intersection(o):
    """ray_sphere_intersection_test erl:1013-1034: t == 7.0; tangent and outside rays miss."""
    sphere = ('sphere', 3, V(0, 0, 10), ('material', ('colour', 0.4, 0.4, 0.4), 1, 0, 0))
    d = V(0, 0, 1)
    hit = o.ray_sphere_intersect(('ray', V(0, 0, 0), d), sphere)
    assert hit[0] == 7.0
    assert o.ray_sphere_intersect(('ray', V(3, 0, 0), d), sphere) is None
    assert o.ray_sphere_intersect(('ray', V(4, 0, 0), d), sphere) is None


def test_point_on_screen(o):
    """point_on_screen_test erl:1036-1066."""
    cam1 = ('camera', V(0, 0, 0), V(0, 0, 0), 90, ('screen', 1, 1))
    cam2 = ('camera', V(0, 0, 0), V(0, 0, 0), 90, ('screen', 640, 480))
    assert eq(V(0, 0, 0.5), o.point_on_screen(0.5, 0.5, cam1))
    assert eq(V(-0.5, -0.5, 0.5), o.point_on_screen(0, 0, cam1))
    assert eq(V(0.5, 0.5, 0.5), o.point_on_screen(1, 1, cam1))
    assert eq(o.point_on_screen(0, 0, cam2), V(-320, -240, 320))
    assert eq(o.point_on_screen(1, 1, cam2), V(320, 240, 320))
    assert eq(o.point_on_screen(0.5, 0.5, cam2), V(0, 0, 320))


def test_nearest_object_intersecting_ray(o):
    """nearest_object_intersecting_ray_test erl:1068-1097."""
    def sph(z, b):
        return ('sphere', 5, V(0, 0, z), ('material', ('colour', 0, 0, b), 1, 0, 0))
    scene = [sph(10, 0.03), sph(20, 0.06), sph(30, 0.09), sph(-10, -0.4)]
    ray = ('ray', V(0, 0, 0), V(0, 0, 1))
    res = o.nearest_object_intersecting_ray(ray, scene)
    _i, obj, dist, hit, normal = res
    assert obj == scene[0] and dist == 5
    assert eq(normal, po.vector_neg(ray[2]))
    assert po.point_on_sphere(scene[0], hit)


def test_focal_length(o):
    """focal_length_test erl:1099-1113 (tolerance 0.1)."""
    for focal, angle in ((13, 108), (15, 100.4), (18, 90), (21, 81.2)):
        v = o.focal_length(angle, 36)
        assert focal + 0.1 >= v >= focal - 0.1
    assert o.focal_length(90, 4) == 2.0000000000000004


def test_vector_bounce_off_plane(o):
    """vector_bounce_off_plane_test erl:1115-1133."""
    v1, v2 = V(1, 1, 0), V(0, -1, 0)
    assert eq(o.vector_bounce_off_plane(v1, o.vector_normalize(v2)), V(1, -1, 0))
    assert eq(o.vector_bounce_off_plane(v2, o.vector_normalize(v1)), V(1, 0, 0))
