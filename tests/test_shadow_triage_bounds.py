"""CPU check of the FP32 shadow triage (shadow_blocked, eraytracer_b200/csrc/ert_wavefront.cuh).

The triage declares a shadow ray blocked without any FP64 when its float32 arithmetic proves that a sphere S is met
before the target.  Here the same operations are restated in numpy float32 (one rounding per operation, the device
code is compiled with -fmad=false; rsqrtf is modelled as the correctly rounded value pushed up to 2 ulp either way)
and held, over millions of random configurations built to sit on the margins, against the reference's own test
evaluated in double (raytracer.erl:364-397, literal operation order, and in long double for the grazing cases):

    triage says "blocked"  ==>  S is hit (Discriminant >= 0.001, both roots >= 0) and Distance(S) < Distance(target)
                               whenever the target is hit at all.

This pins the proof's margins without a GPU; the GPU suites compare the kernel's decisions with the literal path.
"""
import numpy as np
import pytest

U = np.float32(2.0 ** -24)
F = np.float32


def sphere_exact(O, D, C, r):
    """raytracer.erl:364-397 in double, literal order; returns (hit, t)."""
    ox, oy, oz = O[:, 0] - C[:, 0], O[:, 1] - C[:, 1], O[:, 2] - C[:, 2]
    a = D[:, 0] * D[:, 0] + D[:, 1] * D[:, 1] + D[:, 2] * D[:, 2]
    b = 2.0 * (D[:, 0] * ox + D[:, 1] * oy + D[:, 2] * oz)
    c = ox * ox + oy * oy + oz * oz - r * r
    disc = b * b - 4.0 * a * c
    ok = disc >= 0.001
    sq = np.sqrt(np.where(ok, disc, 0.0))
    t0, t1 = (-b + sq) / 2.0, (-b - sq) / 2.0
    hit = ok & (t0 >= 0.0) & (t1 >= 0.0)
    return hit, np.minimum(t0, t1)


def filter_sphere(C, r):
    """make_filter_sphere (ert_api.cu): float centre, R >= r^2 with its slack, per-sphere pad_c and eta_c."""
    cf = C.astype(np.float32)
    eta = 2.0 * np.max(np.abs(C - cf.astype(np.float64)), axis=1)
    pad = 16.0 * r * eta + 4.0 * eta * eta / float(U)
    R = r * r * (1.0 + 3.814697265625e-6) + pad
    Rf = R.astype(np.float32)
    Rf = np.where(Rf.astype(np.float64) < R, np.nextafter(Rf, F(np.inf)), Rf)
    Rf = np.nextafter(Rf, F(np.inf))
    up = lambda x: np.where(x.astype(np.float32).astype(np.float64) < x, np.nextafter(x.astype(np.float32), F(np.inf)), x.astype(np.float32))
    return cf, Rf, up(pad), up(eta)


def triage_blocked(O, P, cf, Rf, tcf, tRf, eta_c_max, pad_c_max, rng):
    """shadow_blocked() for a sphere target and ONE candidate, operation for operation in float32."""
    u = U
    f = (P - O).astype(np.float32)                                   # (float)(P.x - lt[3]): double subtraction, one rounding
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    l2 = fx * fx + fy * fy + fz * fz
    inv = (F(1.0) / np.sqrt(l2)).astype(np.float32)
    inv = (inv * (F(1.0) + rng.integers(-2, 3, len(inv)).astype(np.float32) * F(2.0 ** -23))).astype(np.float32)   # rsqrtf: <= 2 ulp
    dx, dy, dz = fx * inv, fy * inv, fz * inv
    of = O.astype(np.float32)
    ox, oy, oz = of[:, 0], of[:, 1], of[:, 2]
    ec0 = F(1.01) * (eta_c_max + F(1.75) * u * np.maximum(np.maximum(np.abs(ox), np.abs(oy)), np.abs(oz)))
    tx, ty, tz = tcf[:, 0] - ox, tcf[:, 1] - oy, tcf[:, 2] - oz
    trho2 = tx * tx + ty * ty + tz * tz
    tb = dx * tx + dy * ty + dz * tz
    s_lo = tb - (ec0 + F(26.0) * u * np.sqrt(trho2)) - np.sqrt(tRf) * (F(1.0) + F(4.0) * u) - F(4.0) * u * np.abs(tb) - F(1e-12) * trho2 - F(1e-7)
    cx, cy, cz = cf[:, 0] - ox, cf[:, 1] - oy, cf[:, 2] - oz
    rho2 = cx * cx + cy * cy + cz * cz
    rho = np.sqrt(rho2)
    b = dx * cx + dy * cy + dz * cz
    qx, qy, qz = cy * dz - cz * dy, cz * dx - cx * dz, cx * dy - cy * dx
    err = ec0 + F(26.0) * u * rho
    qmax = np.sqrt(qx * qx + qy * qy + qz * qz) * (F(1.0) + F(8.0) * u) + err
    noise = F(1e-12) * rho2
    T = ((Rf - pad_c_max) * (F(1.0) - F(7.62939453125e-6)) - F(8.0) * u * Rf) - qmax * qmax * (F(1.0) + F(4.0) * u) - F(8.0) * u * Rf
    sure = (T > F(0.0004) + noise) & (b - err > np.sqrt(Rf) * (F(1.0) + F(4.0) * u))
    s_up = b + err - np.sqrt(np.maximum(T, F(0))) * (F(1.0) - F(4.0) * u)
    s_up = s_up + (F(4.0) * u * b + F(4e-6) * np.abs(s_up) + noise + F(1e-7))
    for arr in (l2, inv, dx, ec0, s_lo, T, s_up):
        assert arr.dtype == np.float32
    return sure & (s_up < s_lo)


def configs(rng, n, scale, offset, exact_f32):
    """Light O, target sphere T with the hit P on it, candidate S placed ON THE MARGINS: its centre at a distance from
    the ray between 0.9 and 1.05 radii in a third of the cases (grazing), and at a distance along the ray that puts its
    entry point within a few ulp-scaled margins of the target's in another third."""
    O = offset + rng.normal(size=(n, 3)) * scale * 10.0 ** rng.uniform(-1, 1, (n, 1))
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    L = scale * 10.0 ** rng.uniform(-0.5, 2.0, n)                   # distance light -> hit
    P = O + dirs * L[:, None]
    rT = scale * 10.0 ** rng.uniform(-2, 0.5, n)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    CT = P - nrm * rT[:, None]                                       # P lies on the target
    rS = scale * 10.0 ** rng.uniform(-2, 0.5, n)
    # perpendicular offset of S from the ray, and position along it
    perp = np.cross(dirs, rng.normal(size=(n, 3))); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    kind = rng.integers(0, 3, n)
    q = np.where(kind == 0, rS * rng.uniform(0.9, 1.05, n), rS * rng.uniform(0.0, 0.999, n))
    along = L * rng.uniform(0.02, 1.2, n)
    near = kind == 1                                                 # entry point of S next to the target's
    half = np.sqrt(np.maximum(rS * rS - q * q, 0.0))
    along = np.where(near, L + half + rT * rng.uniform(-2.2, 0.2, n) + scale * rng.normal(size=n) * 10.0 ** rng.uniform(-7, -2, n), along)
    CS = O + dirs * along[:, None] + perp * q[:, None]
    if exact_f32:
        O, CT, CS = (x.astype(np.float32).astype(np.float64) for x in (O, CT, CS))
        rT, rS = (x.astype(np.float32).astype(np.float64) for x in (rT, rS))
        P = P.astype(np.float32).astype(np.float64)                  # the hit then sits off the sphere by a rounding: harmless
    return O, P, CT, rT, CS, rS


@pytest.mark.parametrize("scale,offset,exact_f32", [(1.0, 0.0, True), (1.0, 0.0, False), (0.05, 3.0, False),
                                                    (50.0, 1.0e4, True), (30.0, -2.5e4, False), (1.0, 400.0, False)])
def test_blocked_by_the_triage_means_blocked_by_the_reference(scale, offset, exact_f32):
    rng = np.random.default_rng(int(abs(offset)) + int(scale * 100) + (7 if exact_f32 else 0))
    n = 400_000
    said = wrong = 0
    for _ in range(3):
        O, P, CT, rT, CS, rS = configs(rng, n, scale, np.full(3, offset), exact_f32)
        cf, Rf, pad, eta = filter_sphere(CS, rS)
        tcf, tRf, tpad, teta = filter_sphere(CT, rT)
        blocked = triage_blocked(O, P, cf, Rf, tcf, tRf, np.maximum(eta, teta), np.maximum(pad, tpad), rng)
        # the reference: literal FP64 direction (erl:257-260, 554-560) and literal sphere tests
        d = P - O
        mag = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        D = d * (1.0 / mag)[:, None]
        hitS, tS = sphere_exact(O, D, CS, rS)
        hitT, tT = sphere_exact(O, D, CT, rT)
        ref_blocked = ~hitT | (hitS & (tS < tT))                      # nearest != target (erl:263)
        bad = blocked & ~ref_blocked
        said += int(blocked.sum())
        wrong += int(bad.sum())
        assert not bad.any(), ("triage blocked a ray the reference lights", O[bad][0], P[bad][0], CS[bad][0], rS[bad][0], CT[bad][0], rT[bad][0])
    # the check has teeth: the triage does settle a good share of these margin cases
    assert said > 0.05 * 3 * n, said


def plane_exact(O, D, n, dist):
    """raytracer.erl:461-480 in double."""
    vd = n[:, 0] * D[:, 0] + n[:, 1] * D[:, 1] + n[:, 2] * D[:, 2]
    v0 = -((n[:, 0] * O[:, 0] + n[:, 1] * O[:, 1] + n[:, 2] * O[:, 2]) + dist)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = v0 / vd
    hit = (vd < 0.0) & (t >= 0.001)
    return hit, t


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (40.0, 1.0e4), (0.3, -50.0)])
def test_blocked_by_the_triage_with_a_plane_target(scale, offset):
    """The plane branch of shadow_blocked(): the target's lower bound is the distance to the hit, refused for grazing rays."""
    rng = np.random.default_rng(11 + int(abs(offset)))
    n = 400_000
    u = U
    said = 0
    for _ in range(2):
        O = offset + rng.normal(size=(n, 3)) * scale * 5.0
        nrm = rng.normal(size=(n, 3)) * 10.0 ** rng.uniform(-0.5, 0.5, (n, 1))            # un-normalised normals
        dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        L = scale * 10.0 ** rng.uniform(-0.5, 2.0, n)
        P = O + dirs * L[:, None]
        dist = -(nrm[:, 0] * P[:, 0] + nrm[:, 1] * P[:, 1] + nrm[:, 2] * P[:, 2])         # P lies on the plane
        rS = scale * 10.0 ** rng.uniform(-2, 0.5, n)
        perp = np.cross(dirs, rng.normal(size=(n, 3))); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
        q = np.where(rng.random(n) < 0.4, rS * rng.uniform(0.9, 1.05, n), rS * rng.uniform(0.0, 0.999, n))
        half = np.sqrt(np.maximum(rS * rS - q * q, 0.0))
        along = np.where(rng.random(n) < 0.5, L + half + scale * rng.normal(size=n) * 10.0 ** rng.uniform(-7, -1, n), L * rng.uniform(0.02, 1.2, n))
        CS = O + dirs * along[:, None] + perp * q[:, None]
        cf, Rf, pad, eta = filter_sphere(CS, rS)
        # shadow_blocked(), plane target
        f = (P - O).astype(np.float32)
        fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
        l2 = fx * fx + fy * fy + fz * fz
        inv = (F(1.0) / np.sqrt(l2)).astype(np.float32)
        inv = (inv * (F(1.0) + rng.integers(-2, 3, n).astype(np.float32) * F(2.0 ** -23))).astype(np.float32)
        dx, dy, dz = fx * inv, fy * inv, fz * inv
        of = O.astype(np.float32)
        ox, oy, oz = of[:, 0], of[:, 1], of[:, 2]
        ec0 = F(1.01) * (eta + F(1.75) * u * np.maximum(np.maximum(np.abs(ox), np.abs(oy)), np.abs(oz)))
        nf = nrm.astype(np.float32)
        nd = nf[:, 0] * dx + nf[:, 1] * dy + nf[:, 2] * dz
        ok = np.abs(nd) > F(1e-3) * np.sqrt(nf[:, 0] * nf[:, 0] + nf[:, 1] * nf[:, 1] + nf[:, 2] * nf[:, 2])
        s_lo = l2 * inv * (F(1.0) - F(1e-5)) - F(1e-7)
        cx, cy, cz = cf[:, 0] - ox, cf[:, 1] - oy, cf[:, 2] - oz
        rho2 = cx * cx + cy * cy + cz * cz
        rho = np.sqrt(rho2)
        b = dx * cx + dy * cy + dz * cz
        qx, qy, qz = cy * dz - cz * dy, cz * dx - cx * dz, cx * dy - cy * dx
        err = ec0 + F(26.0) * u * rho
        qmax = np.sqrt(qx * qx + qy * qy + qz * qz) * (F(1.0) + F(8.0) * u) + err
        noise = F(1e-12) * rho2
        T = ((Rf - pad) * (F(1.0) - F(7.62939453125e-6)) - F(8.0) * u * Rf) - qmax * qmax * (F(1.0) + F(4.0) * u) - F(8.0) * u * Rf
        sure = (T > F(0.0004) + noise) & (b - err > np.sqrt(Rf) * (F(1.0) + F(4.0) * u))
        s_up = b + err - np.sqrt(np.maximum(T, F(0))) * (F(1.0) - F(4.0) * u)
        s_up = s_up + (F(4.0) * u * b + F(4e-6) * np.abs(s_up) + noise + F(1e-7))
        blocked = ok & sure & (s_up < s_lo)
        # the reference
        d = P - O
        mag = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        D = d * (1.0 / mag)[:, None]
        hitS, tS = sphere_exact(O, D, CS, rS)
        hitP, tP = plane_exact(O, D, nrm, dist)
        ref_blocked = ~hitP | (hitS & (tS < tP))
        bad = blocked & ~ref_blocked
        said += int(blocked.sum())
        assert not bad.any(), ("triage blocked a ray the reference lights", O[bad][0], P[bad][0], CS[bad][0], rS[bad][0], nrm[bad][0])
    assert said > 0.05 * 2 * n, said
