/*
 * raytracer_gpu_nif.c — Erlang NIF over the C ABI of include/ert_b200.h.
 *
 * Glue only: Erlang term -> ert_scene_desc / ert_render_params -> ert_* call ->
 * Erlang term.  All logic lives behind ert_b200.h.  Every NIF that touches the
 * GPU is a dirty CPU-bound NIF (a 4K frame blocks for milliseconds to seconds).
 * No ports, no CPU fallback: errors come back as {error, {Class, Code, Msg}}
 * and undecodable terms raise badarg.
 *
 * Written against the OTP erl_nif API without an OTP installation in this
 * image (no erl_nif.h here).  tests/mock_erl/ holds a minimal stand-in header
 * and term library so that this file is compiled and its decode/encode paths
 * are executed by the test-suite; building against a real OTP only needs
 *   gcc -shared -fPIC -I$ERL_ROOT/usr/include -Iinclude c_src/raytracer_gpu_nif.c \
 *       -Leraytracer_b200/lib -lert_b200 -o priv/raytracer_gpu.so
 *
 * Scene wire format = the record tuples scene_test pins (raytracer.erl:760-801):
 *   {camera,{vector,X,Y,Z},{vector,..},Fov,{screen,W,H}}
 *   {point_light,{colour,R,G,B},{vector,..},{colour,..}}
 *   {sphere,Radius,{vector,..},{material,{colour,..},SpecPow,Shininess,Reflectivity}}
 *   {triangle,V1,V2,V3,Material}      {plane,{vector,..},Distance,Material}
 * Numbers may be Erlang integers or floats (raytracer.erl:619-664 mixes them).
 * List elements that are none of these are skipped (raytracer.erl:357-358).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "erl_nif.h"
#include "ert_b200.h"

static ErlNifResourceType *scene_rt;

typedef struct {
    ert_scene *scene;
} scene_res;

static void scene_dtor(ErlNifEnv *env, void *obj)
{
    scene_res *r = (scene_res *)obj;
    (void)env;
    if (r->scene) ert_scene_destroy(r->scene);
    r->scene = NULL;
}

static int load(ErlNifEnv *env, void **priv, ERL_NIF_TERM info)
{
    (void)priv; (void)info;
    scene_rt = enif_open_resource_type(env, NULL, "ert_b200_scene", scene_dtor, ERL_NIF_RT_CREATE, NULL);
    return scene_rt ? 0 : 1;
}

static int upgrade(ErlNifEnv *env, void **priv, void **old_priv, ERL_NIF_TERM info)
{
    (void)old_priv;
    return load(env, priv, info);
}

/* ---- term helpers --------------------------------------------------------- */
static int is_atom_named(ErlNifEnv *env, ERL_NIF_TERM t, const char *name)
{
    char buf[32];
    if (!enif_is_atom(env, t)) return 0;
    if (enif_get_atom(env, t, buf, sizeof buf, ERL_NIF_LATIN1) <= 0) return 0;
    return strcmp(buf, name) == 0;
}

/* enif_get_double fails on integers: try both */
static int get_number(ErlNifEnv *env, ERL_NIF_TERM t, double *out)
{
    ErlNifSInt64 i;
    if (enif_get_double(env, t, out)) return 1;
    if (enif_get_int64(env, t, &i)) { *out = (double)i; return 1; }
    return 0;
}

/* {Tag, A, B, C} */
static int get_tagged3(ErlNifEnv *env, ERL_NIF_TERM t, const char *tag, double out[3])
{
    int arity;
    const ERL_NIF_TERM *e;
    if (!enif_get_tuple(env, t, &arity, &e) || arity != 4 || !is_atom_named(env, e[0], tag)) return 0;
    return get_number(env, e[1], &out[0]) && get_number(env, e[2], &out[1]) && get_number(env, e[3], &out[2]);
}

static int get_material(ErlNifEnv *env, ERL_NIF_TERM t, ert_material *m)
{
    int arity;
    const ERL_NIF_TERM *e;
    if (!enif_get_tuple(env, t, &arity, &e) || arity != 5 || !is_atom_named(env, e[0], "material")) return 0;
    return get_tagged3(env, e[1], "colour", m->colour) && get_number(env, e[2], &m->specular_power) &&
           get_number(env, e[3], &m->shininess) && get_number(env, e[4], &m->reflectivity);
}

static int get_camera(ErlNifEnv *env, ERL_NIF_TERM t, ert_camera *c)
{
    int arity, sa;
    const ERL_NIF_TERM *e, *s;
    if (!enif_get_tuple(env, t, &arity, &e) || arity != 5 || !is_atom_named(env, e[0], "camera")) return 0;
    if (!get_tagged3(env, e[1], "vector", c->location) || !get_tagged3(env, e[2], "vector", c->rotation) ||
        !get_number(env, e[3], &c->fov))
        return 0;
    if (!enif_get_tuple(env, e[4], &sa, &s) || sa != 3 || !is_atom_named(env, s[0], "screen")) return 0;
    return get_number(env, s[1], &c->screen_width) && get_number(env, s[2], &c->screen_height);
}

typedef struct {
    ert_scene_desc desc;
    ert_point_light *lights;
    ert_sphere *spheres;
    ert_triangle *triangles;
    ert_plane *planes;
} decoded_scene;

static void decoded_free(decoded_scene *d)
{
    free(d->lights); free(d->spheres); free(d->triangles); free(d->planes);
    memset(d, 0, sizeof *d);
}

/* [Camera | Rest] (raytracer.erl:180) -> tables; returns 1 on success, 0 => badarg */
static int decode_scene(ErlNifEnv *env, ERL_NIF_TERM list, decoded_scene *d)
{
    unsigned len;
    ERL_NIF_TERM head, tail;
    int32_t order = 0;
    memset(d, 0, sizeof *d);
    if (!enif_get_list_length(env, list, &len) || len == 0) return 0;
    if (!enif_get_list_cell(env, list, &head, &tail) || !get_camera(env, head, &d->desc.camera)) return 0;
    d->lights = calloc(len, sizeof *d->lights);
    d->spheres = calloc(len, sizeof *d->spheres);
    d->triangles = calloc(len, sizeof *d->triangles);
    d->planes = calloc(len, sizeof *d->planes);
    if (!d->lights || !d->spheres || !d->triangles || !d->planes) { decoded_free(d); return 0; }
    while (enif_get_list_cell(env, tail, &head, &tail)) {
        int arity;
        const ERL_NIF_TERM *e;
        int32_t pos = order++;
        if (!enif_get_tuple(env, head, &arity, &e) || arity < 1) continue;          /* unknown element */
        if (arity == 4 && is_atom_named(env, e[0], "point_light")) {
            ert_point_light *l = &d->lights[d->desc.n_lights];
            if (!get_tagged3(env, e[1], "colour", l->diffuse_colour) || !get_tagged3(env, e[2], "vector", l->location) ||
                !get_tagged3(env, e[3], "colour", l->specular_colour))
                goto bad;
            l->order = pos;
            d->desc.n_lights++;
        } else if (arity == 4 && is_atom_named(env, e[0], "sphere")) {
            ert_sphere *s = &d->spheres[d->desc.n_spheres];
            if (!get_number(env, e[1], &s->radius) || !get_tagged3(env, e[2], "vector", s->center) ||
                !get_material(env, e[3], &s->material))
                goto bad;
            s->order = pos;
            d->desc.n_spheres++;
        } else if (arity == 5 && is_atom_named(env, e[0], "triangle")) {
            ert_triangle *t = &d->triangles[d->desc.n_triangles];
            if (!get_tagged3(env, e[1], "vector", t->v1) || !get_tagged3(env, e[2], "vector", t->v2) ||
                !get_tagged3(env, e[3], "vector", t->v3) || !get_material(env, e[4], &t->material))
                goto bad;
            t->order = pos;
            d->desc.n_triangles++;
        } else if (arity == 4 && is_atom_named(env, e[0], "plane")) {
            ert_plane *p = &d->planes[d->desc.n_planes];
            if (!get_tagged3(env, e[1], "vector", p->normal) || !get_number(env, e[2], &p->distance) ||
                !get_material(env, e[3], &p->material))
                goto bad;
            p->order = pos;
            d->desc.n_planes++;
        }
        /* any other tuple: neither object nor light, skipped like erl:357-358 / 248-249 */
    }
    d->desc.lights = d->lights;
    d->desc.spheres = d->spheres;
    d->desc.triangles = d->triangles;
    d->desc.planes = d->planes;
    return 1;
bad:
    decoded_free(d);
    return 0;
}

static ERL_NIF_TERM make_error(ErlNifEnv *env, int code)
{
    const char *cls = code == ERT_ERR_NO_DEVICE ? "no_device" : code == ERT_ERR_CUDA ? "cuda" :
                      code == ERT_ERR_NOMEM ? "nomem" : "badarg";
    return enif_make_tuple2(env, enif_make_atom(env, "error"),
                            enif_make_tuple3(env, enif_make_atom(env, cls), enif_make_int(env, code),
                                             enif_make_string(env, ert_last_error(), ERL_NIF_LATIN1)));
}

/* ---- NIFs ------------------------------------------------------------------ */
/* device_count() -> {ok, N} | {error, _} */
static ERL_NIF_TERM nif_device_count(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[])
{
    int n = 0, rc;
    (void)argc; (void)argv;
    rc = ert_device_count(&n);
    if (rc != ERT_OK) return make_error(env, rc);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), enif_make_int(env, n));
}

/* scene_info(Scene) -> {NLights, NSpheres, NTriangles, NPlanes}; badarg on malformed scenes.
 * Needs no GPU: lets the host validate a scene term. */
static ERL_NIF_TERM nif_scene_info(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[])
{
    decoded_scene d;
    ERL_NIF_TERM out;
    (void)argc;
    if (!decode_scene(env, argv[0], &d)) return enif_make_badarg(env);
    out = enif_make_tuple4(env, enif_make_int64(env, d.desc.n_lights), enif_make_int64(env, d.desc.n_spheres),
                           enif_make_int64(env, d.desc.n_triangles), enif_make_int64(env, d.desc.n_planes));
    decoded_free(&d);
    return out;
}

/* scene_upload(Scene, Device) -> {ok, Handle} | {error, _}   (dirty) */
static ERL_NIF_TERM nif_scene_upload(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[])
{
    decoded_scene d;
    int device, rc;
    ert_scene *scene = NULL;
    scene_res *res;
    ERL_NIF_TERM term;
    (void)argc;
    if (!enif_get_int(env, argv[1], &device)) return enif_make_badarg(env);
    if (!decode_scene(env, argv[0], &d)) return enif_make_badarg(env);
    rc = ert_scene_create(&d.desc, device, &scene);
    decoded_free(&d);
    if (rc == ERT_ERR_BADARG) return enif_make_badarg(env);
    if (rc != ERT_OK) return make_error(env, rc);
    res = enif_alloc_resource(scene_rt, sizeof *res);
    if (!res) { ert_scene_destroy(scene); return make_error(env, ERT_ERR_NOMEM); }
    res->scene = scene;
    term = enif_make_resource(env, res);
    enif_release_resource(res);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

/* Opts: proplist of {format, rgb8|f32|f64} | {accel, auto|exact|linear|bvh|bvh_mega|grid} |
 *       {part, {BandRows, NParts, Part}} | {camera, CameraRecord} */
static int decode_opts(ErlNifEnv *env, ERL_NIF_TERM opts, ert_render_params *p, ert_camera *cam)
{
    ERL_NIF_TERM head, tail = opts;
    while (enif_get_list_cell(env, tail, &head, &tail)) {
        int arity;
        const ERL_NIF_TERM *e;
        if (!enif_get_tuple(env, head, &arity, &e) || arity != 2) return 0;
        if (is_atom_named(env, e[0], "format")) {
            if (is_atom_named(env, e[1], "rgb8")) p->format = ERT_FMT_RGB8;
            else if (is_atom_named(env, e[1], "f32")) p->format = ERT_FMT_F32;
            else if (is_atom_named(env, e[1], "f64")) p->format = ERT_FMT_F64;
            else return 0;
        } else if (is_atom_named(env, e[0], "accel")) {
            if (is_atom_named(env, e[1], "auto")) p->accel = ERT_ACCEL_AUTO;
            else if (is_atom_named(env, e[1], "exact")) p->accel = ERT_ACCEL_EXACT;
            else if (is_atom_named(env, e[1], "linear")) p->accel = ERT_ACCEL_LINEAR;
            else if (is_atom_named(env, e[1], "bvh")) p->accel = ERT_ACCEL_BVH;
            else if (is_atom_named(env, e[1], "bvh_mega")) p->accel = ERT_ACCEL_BVH_MEGAKERNEL;
            else if (is_atom_named(env, e[1], "grid")) p->accel = ERT_ACCEL_GRID;
            else return 0;
        } else if (is_atom_named(env, e[0], "part")) {
            int a3;
            const ERL_NIF_TERM *q;
            if (!enif_get_tuple(env, e[1], &a3, &q) || a3 != 3) return 0;
            if (!enif_get_int(env, q[0], &p->band_rows) || !enif_get_int(env, q[1], &p->n_parts) ||
                !enif_get_int(env, q[2], &p->part))
                return 0;
        } else if (is_atom_named(env, e[0], "camera")) {
            if (!get_camera(env, e[1], cam)) return 0;
            p->camera = cam;
        } else {
            return 0;
        }
    }
    return enif_is_empty_list(env, tail);
}

static size_t elem_size(int format) { return format == ERT_FMT_RGB8 ? 1 : format == ERT_FMT_F32 ? 4 : 8; }

static int decode_render_args(ErlNifEnv *env, const ERL_NIF_TERM argv[], scene_res **res, ert_render_params *p,
                              ert_camera *cam)
{
    memset(p, 0, sizeof *p);
    if (!enif_get_resource(env, argv[0], scene_rt, (void **)res) || !(*res)->scene) return 0;
    if (!enif_get_int(env, argv[1], &p->width) || !enif_get_int(env, argv[2], &p->height) ||
        !enif_get_int(env, argv[3], &p->depth))
        return 0;
    p->format = ERT_FMT_RGB8;
    p->accel = ERT_ACCEL_AUTO;
    p->n_parts = 1;
    return decode_opts(env, argv[4], p, cam);
}

/* render(Handle, Width, Height, Depth, Opts) -> {ok, FrameBinary} | {error, _}   (dirty)
 * The binary is the full row-major frame (Y = 0 first), 3 channels per pixel. */
static ERL_NIF_TERM nif_render(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[])
{
    scene_res *res;
    ert_render_params p;
    ert_camera cam;
    ERL_NIF_TERM bin;
    unsigned char *buf;
    size_t bytes;
    int rc;
    (void)argc;
    if (!decode_render_args(env, argv, &res, &p, &cam)) return enif_make_badarg(env);
    if (p.width <= 0 || p.height <= 0 || p.depth < 0) return enif_make_badarg(env);   /* guards erl:89 */
    bytes = (size_t)p.width * (size_t)p.height * 3 * elem_size(p.format);
    buf = enif_make_new_binary(env, bytes, &bin);
    if (!buf) return make_error(env, ERT_ERR_NOMEM);
    memset(buf, 0, bytes);
    rc = ert_render(res->scene, &p, buf, bytes);
    if (rc == ERT_ERR_BADARG) return enif_make_badarg(env);
    if (rc != ERT_OK) return make_error(env, rc);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), bin);
}

/* render_pixel_list(Handle, Width, Height, Depth, Opts) -> [{Index, {R,G,B}}] | {error, _}   (dirty)
 * The reference's own return type (raytracer.erl:86-99): unclamped floats, row-major,
 * Index = X + Y*Width as the concurrent driver numbers pixels (erl:112).  For small images:
 * the list costs ~100 bytes of BEAM heap per pixel. */
static ERL_NIF_TERM nif_render_pixel_list(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[])
{
    scene_res *res;
    ert_render_params p;
    ert_camera cam;
    double *buf;
    size_t n, bytes;
    ERL_NIF_TERM list;
    int rc;
    (void)argc;
    if (!decode_render_args(env, argv, &res, &p, &cam)) return enif_make_badarg(env);
    if (p.width <= 0 || p.height <= 0 || p.depth < 0) return enif_make_badarg(env);
    p.format = ERT_FMT_F64;
    n = (size_t)p.width * (size_t)p.height;
    bytes = n * 3 * sizeof(double);
    buf = malloc(bytes);
    if (!buf) return make_error(env, ERT_ERR_NOMEM);
    rc = ert_render(res->scene, &p, buf, bytes);
    if (rc != ERT_OK) {
        free(buf);
        return rc == ERT_ERR_BADARG ? enif_make_badarg(env) : make_error(env, rc);
    }
    list = enif_make_list(env, 0);
    while (n-- > 0) {
        ERL_NIF_TERM px = enif_make_tuple3(env, enif_make_double(env, buf[3 * n]), enif_make_double(env, buf[3 * n + 1]),
                                           enif_make_double(env, buf[3 * n + 2]));
        list = enif_make_list_cell(env, enif_make_tuple2(env, enif_make_int64(env, (ErlNifSInt64)n), px), list);
    }
    free(buf);
    return list;
}

static ErlNifFunc nif_funcs[] = {
    {"device_count", 0, nif_device_count, 0},
    {"scene_info", 1, nif_scene_info, 0},
    {"scene_upload", 2, nif_scene_upload, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"render", 5, nif_render, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"render_pixel_list", 5, nif_render_pixel_list, ERL_NIF_DIRTY_JOB_CPU_BOUND},
};

ERL_NIF_INIT(raytracer_gpu, nif_funcs, load, NULL, upgrade, NULL)
