/*
 * erl_nif.h — MOCK of the OTP NIF API, for tests only.
 *
 * Erlang/OTP is not installed in this image, so c_src/raytracer_gpu_nif.c cannot be built
 * against the real header.  This stand-in declares exactly the subset of the erl_nif API the
 * NIF uses, with the real names and argument orders, backed by a tiny heap-term model
 * (mock_erl_nif.c), so that the NIF is compiled with -Wall -Werror and its term decode /
 * encode paths are executed by the test-suite.  It is not part of the product.
 *
 * Every prototype below restates the signature printed in the erl_nif(3) manual page of OTP 26
 * (section "Exports"), from memory — nothing here was checked against an OTP installation:
 *   int enif_is_atom(ErlNifEnv* env, ERL_NIF_TERM term)
 *   int enif_is_empty_list(ErlNifEnv* env, ERL_NIF_TERM term)
 *   int enif_get_atom(ErlNifEnv* env, ERL_NIF_TERM term, char* buf, unsigned size, ErlNifCharEncoding encode)
 *   int enif_get_double(ErlNifEnv* env, ERL_NIF_TERM term, double* dp)          (fails on integers)
 *   int enif_get_int(ErlNifEnv* env, ERL_NIF_TERM term, int* ip)
 *   int enif_get_int64(ErlNifEnv* env, ERL_NIF_TERM term, ErlNifSInt64* ip)
 *   int enif_get_tuple(ErlNifEnv* env, ERL_NIF_TERM term, int* arity, const ERL_NIF_TERM** array)
 *   int enif_get_list_cell(ErlNifEnv* env, ERL_NIF_TERM list, ERL_NIF_TERM* head, ERL_NIF_TERM* tail)
 *   int enif_get_list_length(ErlNifEnv* env, ERL_NIF_TERM term, unsigned* len)
 *   ERL_NIF_TERM enif_make_atom(ErlNifEnv* env, const char* name)
 *   ERL_NIF_TERM enif_make_int(ErlNifEnv* env, int i) / enif_make_int64(ErlNifEnv* env, ErlNifSInt64 i)
 *   ERL_NIF_TERM enif_make_double(ErlNifEnv* env, double d)                     (d must be finite)
 *   ERL_NIF_TERM enif_make_tuple(ErlNifEnv* env, unsigned cnt, ...) and the enif_make_tupleN macros
 *   ERL_NIF_TERM enif_make_list(ErlNifEnv* env, unsigned cnt, ...) / enif_make_list_cell(env, head, tail)
 *   ERL_NIF_TERM enif_make_string(ErlNifEnv* env, const char* string, ErlNifCharEncoding encoding)
 *   unsigned char* enif_make_new_binary(ErlNifEnv* env, size_t size, ERL_NIF_TERM* termp)
 *   ERL_NIF_TERM enif_make_badarg(ErlNifEnv* env)
 *   ErlNifResourceType* enif_open_resource_type(ErlNifEnv* env, const char* module_str, const char* name,
 *        ErlNifResourceDtor* dtor, ErlNifResourceFlags flags, ErlNifResourceFlags* tried)
 *        (ERL_NIF_RT_CREATE alone fails when the type exists; upgrade passes CREATE | TAKEOVER)
 *   void* enif_alloc_resource(ErlNifResourceType* type, unsigned size) / void enif_release_resource(void* obj)
 *   ERL_NIF_TERM enif_make_resource(ErlNifEnv* env, void* obj)
 *   int enif_get_resource(ErlNifEnv* env, ERL_NIF_TERM term, ErlNifResourceType* type, void** objp)
 *   ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv* env, void* obj, const void* data, size_t size)
 *   ERL_NIF_INIT(MODULE, ErlNifFunc funcs[], load, NULL, upgrade, unload); ErlNifFunc = {name, arity, fptr, flags}
 *        with flags 0, ERL_NIF_DIRTY_JOB_CPU_BOUND or ERL_NIF_DIRTY_JOB_IO_BOUND
 */
#ifndef MOCK_ERL_NIF_H
#define MOCK_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

typedef uintptr_t ERL_NIF_TERM;
typedef int64_t ErlNifSInt64;
typedef struct enif_environment_t ErlNifEnv;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef void ErlNifResourceDtor(ErlNifEnv *, void *);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef enum { ERL_NIF_LATIN1 = 1 } ErlNifCharEncoding;
#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2

typedef struct {
    const char *name;
    unsigned arity;
    ERL_NIF_TERM (*fptr)(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]);
    unsigned flags;
} ErlNifFunc;

typedef struct {
    const char *name;
    int num_of_funcs;
    ErlNifFunc *funcs;
    int (*load)(ErlNifEnv *, void **priv_data, ERL_NIF_TERM load_info);
    int (*reload)(ErlNifEnv *, void **priv_data, ERL_NIF_TERM load_info);
    int (*upgrade)(ErlNifEnv *, void **priv_data, void **old_priv_data, ERL_NIF_TERM load_info);
    void (*unload)(ErlNifEnv *, void *priv_data);
} ErlNifEntry;

#define ERL_NIF_INIT(NAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD)                       \
    __attribute__((visibility("default"))) ErlNifEntry *nif_init(void)                 \
    {                                                                                  \
        static ErlNifEntry entry = {#NAME, sizeof(FUNCS) / sizeof(*FUNCS), FUNCS,      \
                                    LOAD, RELOAD, UPGRADE, UNLOAD};                    \
        return &entry;                                                                 \
    }

int enif_is_atom(ErlNifEnv *, ERL_NIF_TERM);
int enif_is_empty_list(ErlNifEnv *, ERL_NIF_TERM);
int enif_get_atom(ErlNifEnv *, ERL_NIF_TERM, char *buf, unsigned len, ErlNifCharEncoding);
int enif_get_double(ErlNifEnv *, ERL_NIF_TERM, double *);
int enif_get_int(ErlNifEnv *, ERL_NIF_TERM, int *);
int enif_get_int64(ErlNifEnv *, ERL_NIF_TERM, ErlNifSInt64 *);
int enif_get_tuple(ErlNifEnv *, ERL_NIF_TERM, int *arity, const ERL_NIF_TERM **array);
int enif_get_list_cell(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM *head, ERL_NIF_TERM *tail);
int enif_get_list_length(ErlNifEnv *, ERL_NIF_TERM, unsigned *len);
ERL_NIF_TERM enif_make_atom(ErlNifEnv *, const char *);
ERL_NIF_TERM enif_make_int(ErlNifEnv *, int);
ERL_NIF_TERM enif_make_int64(ErlNifEnv *, ErlNifSInt64);
ERL_NIF_TERM enif_make_double(ErlNifEnv *, double);
ERL_NIF_TERM enif_make_tuple(ErlNifEnv *, unsigned cnt, ...);
#define enif_make_tuple2(env, a, b) enif_make_tuple(env, 2, a, b)
#define enif_make_tuple3(env, a, b, c) enif_make_tuple(env, 3, a, b, c)
#define enif_make_tuple4(env, a, b, c, d) enif_make_tuple(env, 4, a, b, c, d)
#define enif_make_tuple5(env, a, b, c, d, e) enif_make_tuple(env, 5, a, b, c, d, e)
ERL_NIF_TERM enif_make_list(ErlNifEnv *, unsigned cnt, ...);
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv *, ERL_NIF_TERM head, ERL_NIF_TERM tail);
ERL_NIF_TERM enif_make_string(ErlNifEnv *, const char *, ErlNifCharEncoding);
unsigned char *enif_make_new_binary(ErlNifEnv *, size_t size, ERL_NIF_TERM *termp);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *);
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *, const char *module_str, const char *name,
                                            ErlNifResourceDtor *dtor, ErlNifResourceFlags flags,
                                            ErlNifResourceFlags *tried);
void *enif_alloc_resource(ErlNifResourceType *, size_t size);
void enif_release_resource(void *obj);
ERL_NIF_TERM enif_make_resource(ErlNifEnv *, void *obj);
int enif_get_resource(ErlNifEnv *, ERL_NIF_TERM, ErlNifResourceType *, void **objp);
/* erl_nif(3): "ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv* env, void* obj, const void* data, size_t size)" —
 * a binary term whose bytes are `data`, owned by resource `obj`; the resource is kept until the binary is collected */
ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv *, void *obj, const void *data, size_t size);

/* ---- mock-only helpers for the test harness (not in OTP) ----------------- */
ErlNifEnv *mock_env_new(void);
int mock_is_badarg(ERL_NIF_TERM);
int mock_is_binary(ERL_NIF_TERM, const unsigned char **data, size_t *size);
int mock_is_string(ERL_NIF_TERM, const char **s);
void mock_resource_gc(ERL_NIF_TERM);        /* drops the term's reference, runs the destructor at zero */
#endif
