"""ctypes binding for oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module.  The product package must never do so.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

K_UNKNOWN, K_LIGHT, K_SPHERE, K_TRIANGLE, K_PLANE = 0, 1, 2, 3, 4
STRIDE = 16

_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        up = ctypes.POINTER(ctypes.c_uint64)
        for name in ("orc_vector_add", "orc_vector_sub", "orc_vector_cross_product",
                     "orc_vector_bounce_off_plane"):
            getattr(L, name).argtypes = [dp, dp, dp]
            getattr(L, name).restype = None
        for name in ("orc_vector_normalize", "orc_vector_neg"):
            getattr(L, name).argtypes = [dp, dp]
            getattr(L, name).restype = None
        L.orc_vector_scalar_mult.argtypes = [dp, ctypes.c_double, dp]
        L.orc_vector_scalar_mult.restype = None
        for name in ("orc_vector_square_mag", "orc_vector_mag"):
            getattr(L, name).argtypes = [dp]
            getattr(L, name).restype = ctypes.c_double
        L.orc_vector_dot_product.argtypes = [dp, dp]
        L.orc_vector_dot_product.restype = ctypes.c_double
        L.orc_focal_length.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_focal_length.restype = ctypes.c_double
        L.orc_point_on_screen.argtypes = [ctypes.c_double, ctypes.c_double, dp, dp]
        L.orc_point_on_screen.restype = None
        L.orc_shoot_ray.argtypes = [dp, dp, dp]
        L.orc_shoot_ray.restype = None
        L.orc_ray_through_pixel.argtypes = [ctypes.c_double, ctypes.c_double, dp, dp]
        L.orc_ray_through_pixel.restype = None
        L.orc_ray_object_intersect.argtypes = [dp, ctypes.c_int, dp, dp]
        L.orc_ray_object_intersect.restype = ctypes.c_int
        L.orc_nearest_object_intersecting_ray.argtypes = [dp, ctypes.c_int, ip, dp, dp]
        L.orc_nearest_object_intersecting_ray.restype = ctypes.c_int
        L.orc_nearest_batch.argtypes = [ctypes.c_int64, dp, ctypes.c_int, ip, dp, ip, dp]
        L.orc_nearest_batch.restype = None
        L.orc_pixel_colour_from_ray.argtypes = [dp, ctypes.c_int, ip, dp, ctypes.c_int,
                                                ctypes.c_int, dp, up]
        L.orc_pixel_colour_from_ray.restype = None
        L.orc_quantise.argtypes = [ctypes.c_double, ctypes.c_int]
        L.orc_quantise.restype = ctypes.c_int
        L.orc_render.argtypes = [dp, ctypes.c_int, ip, dp, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int64, ip, ip, dp,
                                 ctypes.c_int, up]
        L.orc_render.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def vec(v):
    """('vector',X,Y,Z) or ('colour',R,G,B) or a 3-sequence -> float64[3]."""
    if isinstance(v, tuple) and len(v) == 4 and isinstance(v[0], str):
        v = v[1:]
    return np.ascontiguousarray(np.array(v, dtype=np.float64))


def _mat(m):
    _tag, colour, sp, sh, refl = m
    return [colour[1], colour[2], colour[3], sp, sh, refl]


def camera_array(cam):
    _t, loc, rot, fov, (_s, sw, sh) = cam
    return np.array([loc[1], loc[2], loc[3], rot[1], rot[2], rot[3], fov, sw, sh],
                    dtype=np.float64)


def flatten(elements):
    """Scene list *after the camera* (Erlang-shaped tuples) -> (kind int32[n], f float64[n,16])."""
    n = len(elements)
    kind = np.zeros(n, dtype=np.int32)
    f = np.zeros((n, STRIDE), dtype=np.float64)
    for i, e in enumerate(elements):
        tag = e[0] if isinstance(e, tuple) and e else None
        if tag == 'point_light' and len(e) == 4:
            kind[i] = K_LIGHT
            f[i, 0:9] = [e[1][1], e[1][2], e[1][3], e[2][1], e[2][2], e[2][3],
                         e[3][1], e[3][2], e[3][3]]
        elif tag == 'sphere' and len(e) == 4:
            kind[i] = K_SPHERE
            f[i, 0:10] = [e[1], e[2][1], e[2][2], e[2][3]] + _mat(e[3])
        elif tag == 'triangle' and len(e) == 5:
            kind[i] = K_TRIANGLE
            f[i, 0:15] = [e[1][1], e[1][2], e[1][3], e[2][1], e[2][2], e[2][3],
                          e[3][1], e[3][2], e[3][3]] + _mat(e[4])
        elif tag == 'plane' and len(e) == 4:
            kind[i] = K_PLANE
            f[i, 0:10] = [e[1][1], e[1][2], e[1][3], e[2]] + _mat(e[3])
        else:
            kind[i] = K_UNKNOWN
    return kind, f


def flatten_arrays(lights=None, spheres=None, planes=None, triangles=None, order=None):
    """Flat numpy tables -> oracle encoding without going through tuples (large scenes).

    lights (nl,9), spheres (ns,10) [radius, center3, material6], planes (np,10),
    triangles (nt,15).  Scene order: lights, spheres, triangles, planes unless `order`
    gives, for each of those rows in that concatenation, its list position."""
    parts, kinds = [], []
    for arr, k in ((lights, K_LIGHT), (spheres, K_SPHERE), (triangles, K_TRIANGLE),
                   (planes, K_PLANE)):
        if arr is None or len(arr) == 0:
            continue
        a = np.zeros((len(arr), STRIDE), dtype=np.float64)
        a[:, :arr.shape[1]] = arr
        parts.append(a)
        kinds.append(np.full(len(arr), k, dtype=np.int32))
    if not parts:
        return np.zeros(0, np.int32), np.zeros((0, STRIDE))
    f = np.concatenate(parts)
    kind = np.concatenate(kinds)
    if order is not None:
        order = np.asarray(order)
        f2 = np.empty_like(f)
        k2 = np.empty_like(kind)
        f2[order] = f
        k2[order] = kind
        f, kind = f2, k2
    return np.ascontiguousarray(kind), np.ascontiguousarray(f)


def render(cam, kind, f, width, height, depth, pixels=None, nthreads=None, retrace=False):
    """Returns (rgb float64[npix,3], rays, tests).  pixels: optional (xs, ys) int arrays."""
    L = lib()
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    kind = np.ascontiguousarray(kind, dtype=np.int32)
    f = np.ascontiguousarray(f, dtype=np.float64)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    if pixels is None:
        npix = width * height
        xs = ys = None
        xp = yp = None
    else:
        xs = np.ascontiguousarray(pixels[0], dtype=np.int32)
        ys = np.ascontiguousarray(pixels[1], dtype=np.int32)
        npix = len(xs)
        xp, yp = _ip(xs), _ip(ys)
    out = np.zeros((npix, 3), dtype=np.float64)
    counters = (ctypes.c_uint64 * 2)()
    rc = L.orc_render(_dp(cam), len(kind), _ip(kind), _dp(f), width, height, depth,
                      1 if retrace else 0, npix, xp, yp, _dp(out), int(nthreads), counters)
    if rc != 0:
        raise ValueError("orc_render rejected its arguments")
    return out, int(counters[0]), int(counters[1])


def quantise_image(rgb):
    """erl:678-680 on an array of unclamped doubles -> int64 array (no lower clamp)."""
    return np.minimum(np.trunc(rgb * 255.0), 255.0).astype(np.int64)


def nearest_batch(rays6, kind, f):
    L = lib()
    rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
    n = len(rays6)
    idx = np.zeros(n, dtype=np.int32)
    t = np.zeros(n, dtype=np.float64)
    L.orc_nearest_batch(n, _dp(rays6), len(kind), _ip(kind), _dp(f), _ip(idx), _dp(t))
    return idx, t
