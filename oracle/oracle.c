/*
 * oracle.c — CPU restatement of eraytracer's per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (eraytracer_b200/,
 * the C-ABI library, the NIF) may include, link or call this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and there only as the checker / the timed CPU arm.
 *
 * What it follows: /root/reference/raytracer.erl lines 180-614 (the hot path)
 * plus the quantisation rule at 678-680 and the pixel sampling at 95-97.
 * Every function cites the lines it restates.  All arithmetic is IEEE double
 * in the literal operation order of the Erlang source; build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math
 * so that no FMA contraction or re-association happens (BEAM evaluates each
 * '*' and '+' separately).  sqrt is correctly rounded; pow/tan come from
 * glibc, the same libm a local BEAM's math BIFs would call.
 *
 * PARITY STATUS: the reference cannot be executed in this image (no Erlang).
 * The oracle is pinned against every known-answer value in the reference's
 * own run_tests/0 (vector algebra, focal length, point_on_screen, shoot_ray,
 * ray/sphere, nearest object, bounce — tests/test_oracle_kats.py) and is
 * cross-checked bit-for-bit against an independent pure-Python restatement
 * (oracle/pyoracle.py).  Plane/triangle intersection, shading, shadows,
 * reflection and whole images have NO golden vectors in the reference:
 * for those rows this oracle is "parity unpinned" (source text only).
 *
 * Scene encoding used by this file (chosen for the tests, not by the
 * reference): the scene list *after the camera* is an array of elements in
 * list order, kind[i] in {0 unknown, 1 point_light, 2 sphere, 3 triangle,
 * 4 plane} and 16 doubles f[i][0..15]:
 *   point_light: diffuse rgb 0-2, location 3-5, specular rgb 6-8   (erl:81)
 *   sphere:      radius 0, center 1-3, material 4-9                (erl:78)
 *   triangle:    v1 0-2, v2 3-5, v3 6-8, material 9-14             (erl:79)
 *   plane:       normal 0-2, distance 3, material 4-9              (erl:80)
 *   material:    colour rgb, specular_power, shininess, reflectivity (erl:77)
 * Camera: location 0-2, rotation 3-5 (ignored, erl:487), fov 6,
 *         screen width 7, screen height 8                          (erl:75-76)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_STRIDE 16
enum { K_UNKNOWN = 0, K_LIGHT = 1, K_SPHERE = 2, K_TRIANGLE = 3, K_PLANE = 4 };

typedef struct { double x, y, z; } vec;
typedef struct { vec o, d; } ray;

typedef struct {
    int n;
    const int32_t *kind;
    const double *f;          /* n * ORC_STRIDE */
} scene_t;

typedef struct {
    uint64_t rays;            /* nearest-object scans (primary+reflection+shadow) */
    uint64_t tests;           /* ray_object_intersect calls on sphere/triangle/plane */
} counters_t;

/* ---- vector primitives: erl:524-573 ------------------------------------ */
static inline vec v3(double x, double y, double z) { vec r = {x, y, z}; return r; }
/* erl:524-527 */
static inline vec v_add(vec a, vec b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
/* erl:529-532 */
static inline vec v_sub(vec a, vec b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
/* erl:534-535 */
static inline double v_square_mag(vec a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
/* erl:537-538 */
static inline double v_mag(vec a) { return sqrt(v_square_mag(a)); }
/* erl:540-541 */
static inline vec v_scale(vec a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
/* erl:543-544 */
static inline vec v_cmul(vec a, vec b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
/* erl:546-547 */
static inline double v_dot(vec a, vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* erl:549-552 */
static inline vec v_cross(vec a, vec b)
{
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* erl:554-560: reciprocal then three multiplies; zero vector maps to zero */
static inline vec v_normalize(vec a)
{
    double mag = v_mag(a);
    if (mag == 0) return v3(0, 0, 0);
    return v_scale(a, 1 / v_mag(a));
}
/* erl:562-563 */
static inline vec v_neg(vec a) { return v3(-a.x, -a.y, -a.z); }
/* erl:568-573 */
static inline vec v_bounce(vec v, vec n)
{
    return v_add(v_scale(n, 2 * v_dot(n, v_neg(v))), v);
}
/* lists:max([0, X]) as used at erl:275 and erl:290 */
static inline double max0(double x) { return x > 0 ? x : 0; }

/* ---- camera: erl:483-511 ------------------------------------------------ */
/* erl:483-484 */
static double focal_length(double angle, double dimension)
{
    return dimension / (2 * tan(angle * (M_PI / 180) / 2));
}

/* erl:486-503; the foldl adds each list element to the accumulator that
 * starts at the camera location: add(V1,loc), add(V2,.), add(V3,.) */
static vec point_on_screen(double x, double y, const double *cam)
{
    double sw = cam[7], sh = cam[8];
    vec acc = v3(cam[0], cam[1], cam[2]);
    acc = v_add(v_scale(v3(0, 0, 1), focal_length(cam[6], sw)), acc);
    acc = v_add(v3((x - 0.5) * sw, 0, 0), acc);
    acc = v_add(v3(0, (y - 0.5) * sh, 0), acc);
    return acc;
}

/* erl:506-507 */
static ray shoot_ray(vec from, vec through)
{
    ray r;
    r.o = from;
    r.d = v_normalize(v_sub(through, from));
    return r;
}

/* erl:510-511 */
static ray ray_through_pixel(double x, double y, const double *cam)
{
    return shoot_ray(v3(cam[0], cam[1], cam[2]), point_on_screen(x, y, cam));
}

/* ---- intersectors ------------------------------------------------------- */
typedef struct { double t; vec p, n; } hit_t;

/* erl:364-397 */
static int ray_sphere_intersect(ray r, const double *s, hit_t *h)
{
    double radius = s[0], xc = s[1], yc = s[2], zc = s[3];
    double x0 = r.o.x, y0 = r.o.y, z0 = r.o.z;
    double xd = r.d.x, yd = r.d.y, zd = r.d.z;
    double epsilon = 0.001;
    double a = xd * xd + yd * yd + zd * zd;
    double b = 2 * (xd * (x0 - xc) + yd * (y0 - yc) + zd * (z0 - zc));
    double c = (x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc) + (z0 - zc) * (z0 - zc)
               - radius * radius;
    double disc = b * b - 4 * a * c;
    if (disc >= epsilon) {
        double t0 = (-b + sqrt(disc)) / 2;
        double t1 = (-b - sqrt(disc)) / 2;
        if (t0 >= 0 && t1 >= 0) {
            double dist = t0 < t1 ? t0 : t1;          /* lists:min([T0,T1]) */
            vec p = v_add(v3(x0, y0, z0), v_scale(v3(xd, yd, zd), dist));
            h->t = dist;
            h->p = p;
            h->n = v_normalize(v_sub(p, v3(xc, yc, zc)));
            return 1;
        }
    }
    return 0;
}

/* erl:402-455 */
static int ray_triangle_intersect(ray r, const double *tr, hit_t *h)
{
    double epsilon = 0.000001;
    vec v1 = v3(tr[0], tr[1], tr[2]), v2 = v3(tr[3], tr[4], tr[5]), v3_ = v3(tr[6], tr[7], tr[8]);
    vec edge1 = v_sub(v2, v1);
    vec edge2 = v_sub(v3_, v1);
    vec p = v_cross(r.d, edge2);
    double det = v_dot(edge1, p);
    if (det < epsilon) return 0;
    vec t = v_sub(r.o, v1);
    double u = v_dot(t, p);
    if (u < 0 || u > det) return 0;
    vec q = v_cross(t, edge1);
    double v = v_dot(r.d, q);
    if (v < 0 || u + v > det) return 0;
    double dist = v_dot(edge2, q) / det;
    h->t = dist;
    h->p = v_add(r.o, v_scale(r.d, dist));
    h->n = v_normalize(v_cross(v1, v2));       /* erl:448-451: position vectors */
    return 1;
}

/* erl:461-480 */
static int ray_plane_intersect(ray r, const double *pl, hit_t *h)
{
    double epsilon = 0.001;
    vec n = v3(pl[0], pl[1], pl[2]);
    double vd = v_dot(n, r.d);
    if (vd < 0) {
        double v0 = -(v_dot(n, r.o) + pl[3]);
        double dist = v0 / vd;
        if (dist < epsilon) return 0;
        h->t = dist;
        h->p = v_add(r.o, v_scale(r.d, dist));
        h->n = n;                                 /* erl:476: un-normalised */
        return 1;
    }
    return 0;
}

/* erl:349-359 */
static int ray_object_intersect(ray r, int kind, const double *f, hit_t *h, counters_t *c)
{
    switch (kind) {
    case K_SPHERE:   if (c) c->tests++; return ray_sphere_intersect(r, f, h);
    case K_TRIANGLE: if (c) c->tests++; return ray_triangle_intersect(r, f, h);
    case K_PLANE:    if (c) c->tests++; return ray_plane_intersect(r, f, h);
    default:         return 0;
    }
}

/* erl:300-346: linear scan, strict '>' replacement => earlier element wins ties */
static int nearest_object(ray r, const scene_t *sc, hit_t *best, counters_t *c)
{
    int best_i = -1;
    if (c) c->rays++;
    for (int i = 0; i < sc->n; i++) {
        hit_t h;
        if (ray_object_intersect(r, sc->kind[i], sc->f + (size_t)i * ORC_STRIDE, &h, c)) {
            if (best_i < 0 || best->t > h.t) {
                *best = h;
                best_i = i;
            }
        }
    }
    return best_i;
}

/* material accessors erl:575-601 */
static const double *material_of(int kind, const double *f)
{
    return kind == K_TRIANGLE ? f + 9 : f + 4;
}

static vec pixel_colour_from_ray(ray r, const scene_t *sc, int depth, int retrace, counters_t *c);

/* erl:256-267; "same term" is restated as "same list index": two term-equal
 * objects are geometrically identical, so the earlier one wins both the
 * shading scan and this scan (strict '>'), and index equality agrees. */
static double shadow_factor(vec light, vec hit, int obj, const scene_t *sc, counters_t *c)
{
    ray s;
    hit_t h;
    s.o = light;
    s.d = v_normalize(v_sub(hit, light));
    return nearest_object(s, sc, &h, c) == obj ? 1 : 0;
}

/* erl:272-279 */
static vec diffuse_term(const double *mat, vec light, vec hit, vec normal)
{
    return v_scale(v3(mat[0], mat[1], mat[2]),
                   max0(v_dot(normal, v_normalize(v_sub(light, hit)))));
}

/* erl:285-297 */
static vec specular_term(vec eye, vec light, vec hit, vec normal, double spec_power,
                         double shininess, vec spec_colour)
{
    double base = max0(v_dot(v_normalize(v_add(v_normalize(v_sub(light, hit)), v_neg(eye))),
                             normal));
    return v_scale(spec_colour, shininess * pow(base, spec_power));
}

/* erl:209-252.  The reference evaluates the reflection inside the per-light
 * fold, i.e. re-traces the identical reflection ray once per light.  The
 * value is the same each time (pure function), so with retrace==0 it is
 * traced once and reused — bit-identical output, fewer rays.  retrace==1
 * re-traces literally (for ray accounting and for checking that claim). */
static vec lighting_function(ray r, int obj, hit_t h, const scene_t *sc, int depth,
                             int retrace, counters_t *c)
{
    const double *f = sc->f + (size_t)obj * ORC_STRIDE;
    const double *mat = material_of(sc->kind[obj], f);
    vec final = v3(0, 0, 0);
    vec child = v3(0, 0, 0);
    int have_child = 0;
    for (int i = 0; i < sc->n; i++) {
        if (sc->kind[i] != K_LIGHT) continue;
        const double *l = sc->f + (size_t)i * ORC_STRIDE;
        vec light_colour = v3(l[0], l[1], l[2]);
        vec light_loc = v3(l[3], l[4], l[5]);
        vec spec_colour = v3(l[6], l[7], l[8]);
        if (retrace || !have_child) {
            ray rr;
            rr.o = h.p;
            rr.d = v_bounce(r.d, h.n);
            child = pixel_colour_from_ray(rr, sc, depth - 1, retrace, c);
            have_child = 1;
        }
        vec reflection = v_scale(child, mat[5]);
        vec contribution = v_add(diffuse_term(mat, light_loc, h.p, h.n),
                                 specular_term(r.d, light_loc, h.p, h.n, mat[3], mat[4],
                                               spec_colour));
        final = v_add(final,
                      v_add(reflection,
                            v_scale(v_cmul(light_colour, contribution),
                                    shadow_factor(light_loc, h.p, obj, sc, c))));
    }
    return final;
}

/* erl:186-203 */
static vec pixel_colour_from_ray(ray r, const scene_t *sc, int depth, int retrace, counters_t *c)
{
    hit_t h;
    int obj;
    if (depth == 0) return v3(0, 0, 0);
    obj = nearest_object(r, sc, &h, c);
    if (obj < 0) return v3(0, 0, 0);               /* BACKGROUND_COLOUR erl:82 */
    return lighting_function(r, obj, h, sc, depth, retrace, c);
}

/* ======================= exported test entry points ====================== */
#define EXPORT __attribute__((visibility("default")))

EXPORT void orc_vector_add(const double *a, const double *b, double *o)
{ vec r = v_add(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT void orc_vector_sub(const double *a, const double *b, double *o)
{ vec r = v_sub(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT double orc_vector_square_mag(const double *a) { return v_square_mag(v3(a[0], a[1], a[2])); }
EXPORT double orc_vector_mag(const double *a) { return v_mag(v3(a[0], a[1], a[2])); }
EXPORT void orc_vector_scalar_mult(const double *a, double s, double *o)
{ vec r = v_scale(v3(a[0], a[1], a[2]), s); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT double orc_vector_dot_product(const double *a, const double *b)
{ return v_dot(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2])); }
EXPORT void orc_vector_cross_product(const double *a, const double *b, double *o)
{ vec r = v_cross(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT void orc_vector_normalize(const double *a, double *o)
{ vec r = v_normalize(v3(a[0], a[1], a[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT void orc_vector_neg(const double *a, double *o)
{ vec r = v_neg(v3(a[0], a[1], a[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT void orc_vector_bounce_off_plane(const double *v, const double *n, double *o)
{ vec r = v_bounce(v3(v[0], v[1], v[2]), v3(n[0], n[1], n[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }

EXPORT double orc_focal_length(double angle, double dimension) { return focal_length(angle, dimension); }
EXPORT void orc_point_on_screen(double x, double y, const double *cam, double *o)
{ vec r = point_on_screen(x, y, cam); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
EXPORT void orc_shoot_ray(const double *from, const double *through, double *o)
{
    ray r = shoot_ray(v3(from[0], from[1], from[2]), v3(through[0], through[1], through[2]));
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
}
EXPORT void orc_ray_through_pixel(double x, double y, const double *cam, double *o)
{
    ray r = ray_through_pixel(x, y, cam);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
}

static ray ray_from(const double *r6)
{
    ray r;
    r.o = v3(r6[0], r6[1], r6[2]);
    r.d = v3(r6[3], r6[4], r6[5]);
    return r;
}
static void hit_out(const hit_t *h, double *o7)
{
    o7[0] = h->t; o7[1] = h->p.x; o7[2] = h->p.y; o7[3] = h->p.z;
    o7[4] = h->n.x; o7[5] = h->n.y; o7[6] = h->n.z;
}

/* returns 1 on hit and fills out[7] = {t, P, N}; 0 == the atom 'none' */
EXPORT int orc_ray_object_intersect(const double *ray6, int kind, const double *f, double *out7)
{
    hit_t h;
    if (!ray_object_intersect(ray_from(ray6), kind, f, &h, NULL)) return 0;
    hit_out(&h, out7);
    return 1;
}

/* returns the list index of the nearest object, or -1 for 'none' */
EXPORT int orc_nearest_object_intersecting_ray(const double *ray6, int n, const int32_t *kind,
                                               const double *f, double *out7)
{
    scene_t sc = {n, kind, f};
    hit_t h;
    int i = nearest_object(ray_from(ray6), &sc, &h, NULL);
    if (i >= 0) hit_out(&h, out7);
    return i;
}

/* batch version used to check the GPU ray-batch entry point */
EXPORT void orc_nearest_batch(int64_t nrays, const double *rays6, int n, const int32_t *kind,
                              const double *f, int32_t *out_idx, double *out_t)
{
    scene_t sc = {n, kind, f};
    for (int64_t k = 0; k < nrays; k++) {
        hit_t h;
        int i = nearest_object(ray_from(rays6 + 6 * k), &sc, &h, NULL);
        out_idx[k] = i;
        out_t[k] = i >= 0 ? h.t : 0.0;
    }
}

EXPORT void orc_pixel_colour_from_ray(const double *ray6, int n, const int32_t *kind,
                                      const double *f, int depth, int retrace, double *rgb,
                                      uint64_t *counters2)
{
    scene_t sc = {n, kind, f};
    counters_t c = {0, 0};
    vec col = pixel_colour_from_ray(ray_from(ray6), &sc, depth, retrace, &c);
    rgb[0] = col.x; rgb[1] = col.y; rgb[2] = col.z;
    if (counters2) { counters2[0] = c.rays; counters2[1] = c.tests; }
}

/* erl:678-680: min(trunc(C*MaxValue), MaxValue), no lower clamp */
EXPORT int orc_quantise(double c, int max_value)
{
    double v = trunc(c * max_value);
    return v < max_value ? (int)v : max_value;
}

/* ---- threaded pixel-list render ----------------------------------------- */
typedef struct {
    const double *cam;
    scene_t sc;
    int width, height, depth, retrace, nthreads;
    int64_t npix;
    const int32_t *xs, *ys;     /* NULL => full frame, row-major (erl:90-99) */
    double *out;                /* npix*3 */
    int64_t next;               /* work counter */
    pthread_mutex_t mu;
    counters_t total;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    counters_t c = {0, 0};
    int64_t chunk = j->npix / ((int64_t)j->nthreads * 16);
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    for (;;) {
        int64_t lo, hi;
        pthread_mutex_lock(&j->mu);
        lo = j->next;
        j->next += chunk;
        pthread_mutex_unlock(&j->mu);
        if (lo >= j->npix) break;
        hi = lo + chunk < j->npix ? lo + chunk : j->npix;
        for (int64_t k = lo; k < hi; k++) {
            int x = j->xs ? j->xs[k] : (int)(k % j->width);
            int y = j->ys ? j->ys[k] : (int)(k / j->width);
            /* erl:95-97 / 112 / 176: coordinates passed as X/Width, Y/Height */
            ray r = ray_through_pixel((double)x / (double)j->width,
                                      (double)y / (double)j->height, j->cam);
            /* erl:180-184 */
            vec col = pixel_colour_from_ray(r, &j->sc, j->depth, j->retrace, &c);
            j->out[3 * k] = col.x; j->out[3 * k + 1] = col.y; j->out[3 * k + 2] = col.z;
        }
    }
    pthread_mutex_lock(&j->mu);
    j->total.rays += c.rays;
    j->total.tests += c.tests;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

/* Renders npix pixels.  xs/ys NULL => the whole width*height frame in the
 * row-major order of raytraced_pixel_list_simple (npix is then ignored).
 * out: npix*3 doubles, unclamped {R,G,B} as colour_to_pixel returns them
 * (erl:613-614).  counters2: {rays, object tests} or NULL. */
EXPORT int orc_render(const double *cam, int n, const int32_t *kind, const double *f,
                      int width, int height, int depth, int retrace, int64_t npix,
                      const int32_t *xs, const int32_t *ys, double *out, int nthreads,
                      uint64_t *counters2)
{
    job_t j;
    pthread_t th[256];
    if (width <= 0 || height <= 0 || depth < 0) return -1;
    memset(&j, 0, sizeof j);
    j.cam = cam;
    j.sc.n = n; j.sc.kind = kind; j.sc.f = f;
    j.width = width; j.height = height; j.depth = depth; j.retrace = retrace;
    j.xs = xs; j.ys = ys;
    j.npix = (xs && ys) ? npix : (int64_t)width * height;
    j.out = out;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    j.nthreads = nthreads;
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
    worker(&j);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&j.mu);
    if (counters2) { counters2[0] = j.total.rays; counters2[1] = j.total.tests; }
    return 0;
}
