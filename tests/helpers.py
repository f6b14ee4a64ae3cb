"""Shared helpers of the test-suite: oracle-side views of scenes and the parity metric."""
import numpy as np

from oracle import orc


def oracle_scene_from_flat(flat):
    """FlatScene (product tables) -> (cam[9], kind[n], f[n,16]) in the oracle's encoding."""
    cam = np.array(list(flat.camera.location) + list(flat.camera.rotation)
                   + [flat.camera.fov, flat.camera.screen_width, flat.camera.screen_height],
                   dtype=np.float64)
    orders = np.concatenate([flat.lights['order'], flat.spheres['order'],
                             flat.triangles['order'], flat.planes['order']]).astype(np.int64)
    n = int(orders.max()) + 1 if len(orders) else 0
    kind = np.zeros(n, dtype=np.int32)
    f = np.zeros((n, orc.STRIDE), dtype=np.float64)

    def mat(t):
        m = t['material']
        return np.concatenate([m['colour'], m['specular_power'][:, None], m['shininess'][:, None],
                               m['reflectivity'][:, None]], axis=1)

    if len(flat.lights):
        o = flat.lights['order']
        kind[o] = orc.K_LIGHT
        f[o, 0:9] = np.concatenate([flat.lights['diffuse_colour'], flat.lights['location'],
                                    flat.lights['specular_colour']], axis=1)
    if len(flat.spheres):
        o = flat.spheres['order']
        kind[o] = orc.K_SPHERE
        f[o, 0:10] = np.concatenate([flat.spheres['radius'][:, None], flat.spheres['center'],
                                     mat(flat.spheres)], axis=1)
    if len(flat.triangles):
        o = flat.triangles['order']
        kind[o] = orc.K_TRIANGLE
        f[o, 0:15] = np.concatenate([flat.triangles['v1'], flat.triangles['v2'],
                                     flat.triangles['v3'], mat(flat.triangles)], axis=1)
    if len(flat.planes):
        o = flat.planes['order']
        kind[o] = orc.K_PLANE
        f[o, 0:10] = np.concatenate([flat.planes['normal'], flat.planes['distance'][:, None],
                                     mat(flat.planes)], axis=1)
    return cam, kind, f


def oracle_frame(flat, width, height, depth, pixels=None, nthreads=None):
    cam, kind, f = oracle_scene_from_flat(flat)
    rgb, rays, tests = orc.render(cam, kind, f, width, height, depth, pixels=pixels,
                                  nthreads=nthreads)
    if pixels is None:
        rgb = rgb.reshape(height, width, 3)
    return rgb, rays, tests


def quantise(frame):
    """raytracer.erl:678-680."""
    return np.minimum(np.trunc(np.asarray(frame, dtype=np.float64) * 255.0), 255.0).astype(np.int64)


def parity_report(gpu_q, ref_q):
    """The north-star image metric on 8-bit channels."""
    d = np.abs(np.asarray(gpu_q, dtype=np.int64) - np.asarray(ref_q, dtype=np.int64))
    per_px = d.reshape(-1, 3).max(axis=1)
    return {"max": int(per_px.max()) if per_px.size else 0,
            "frac_le1": float((per_px <= 1).mean()) if per_px.size else 1.0,
            "n_diff": int((per_px > 0).sum())}


def assert_image_parity(gpu_q, ref_q):
    """<= 1 LSB per 8-bit channel on >= 99.9 % of pixels, no pixel off by more than 2."""
    r = parity_report(gpu_q, ref_q)
    assert r["max"] <= 2, r
    assert r["frac_le1"] >= 0.999, r
    return r


def assert_double_parity(gpu, ref, rtol=1e-9, atol=1e-12):
    """Tighter than the image metric: the GPU colour in double must agree with the oracle to
    re-association noise.  A single flipped hit/shadow decision is 1e-3..1 and fails this."""
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(gpu - ref)
    tol = atol + rtol * np.abs(ref)
    bad = err > tol
    assert not bad.any(), ("%d channel values differ, worst %.3e at %s"
                           % (int(bad.sum()), float(err.max()),
                              np.unravel_index(int(err.argmax()), err.shape)))


def clustered_scene():
    """3000 ordinary spheres plus a knot of 70 small ones inside one cell of the cell grid: that cell's list runs
    through both of its blocks, the overflow groups and more than one survivor mask (32 < count <= 127)."""
    from eraytracer_b200 import scene as sc
    rng = np.random.default_rng(31)
    flat = sc.synthetic_scene("c3", n_spheres=3000, seed=31)
    c = np.stack([rng.uniform(-b, b, 3000) for b in (30, 12, 30)], axis=1) + np.asarray((0.0, -14.0, 45.0))
    r = rng.uniform(0.2, 0.6, 3000)
    knot = np.array([3.3, -13.1, 44.2]) + rng.uniform(-0.35, 0.35, (70, 3))
    c[:70] = knot
    r[:70] = rng.uniform(0.03, 0.08, 70)
    flat.spheres['center'] = c.astype(np.float32).astype(np.float64)
    flat.spheres['radius'] = r.astype(np.float32).astype(np.float64)
    return flat, flat.spheres['center'][:70].copy()
