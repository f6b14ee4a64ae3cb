/* Heap-term model behind tests/mock_erl/erl_nif.h (tests only; leaks by design). */
#include "erl_nif.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

enum { T_ATOM = 1, T_INT, T_FLOAT, T_TUPLE, T_CONS, T_NIL, T_BIN, T_RES, T_BADARG, T_STR };

typedef struct term {
    int tag;
    union {
        char atom[64];
        int64_t i;
        double f;
        struct { int arity; ERL_NIF_TERM *e; } tup;
        struct { ERL_NIF_TERM head, tail; } cons;
        struct { size_t size; unsigned char *data; } bin;
        void *res;
        char *str;
    } u;
    void *owner;            /* resource that owns a resource binary's bytes */
} term;

struct enif_environment_t { int unused; };
struct enif_resource_type_t { ErlNifResourceDtor *dtor; };
typedef struct { ErlNifResourceType *type; int refs; } res_hdr;

static term *T(ERL_NIF_TERM t) { return (term *)t; }
static ERL_NIF_TERM mk(int tag)
{
    term *t = calloc(1, sizeof *t);
    t->tag = tag;
    return (ERL_NIF_TERM)t;
}

ErlNifEnv *mock_env_new(void) { return calloc(1, sizeof(ErlNifEnv)); }
int enif_is_atom(ErlNifEnv *e, ERL_NIF_TERM t) { (void)e; return T(t)->tag == T_ATOM; }
int enif_is_empty_list(ErlNifEnv *e, ERL_NIF_TERM t) { (void)e; return T(t)->tag == T_NIL; }
int enif_get_atom(ErlNifEnv *e, ERL_NIF_TERM t, char *buf, unsigned len, ErlNifCharEncoding enc)
{
    (void)e; (void)enc;
    if (T(t)->tag != T_ATOM || strlen(T(t)->u.atom) + 1 > len) return 0;
    strcpy(buf, T(t)->u.atom);
    return (int)strlen(buf) + 1;
}
int enif_get_double(ErlNifEnv *e, ERL_NIF_TERM t, double *d)
{
    (void)e;
    if (T(t)->tag != T_FLOAT) return 0;      /* like OTP: fails on integers */
    *d = T(t)->u.f;
    return 1;
}
int enif_get_int64(ErlNifEnv *e, ERL_NIF_TERM t, ErlNifSInt64 *i)
{
    (void)e;
    if (T(t)->tag != T_INT) return 0;
    *i = T(t)->u.i;
    return 1;
}
int enif_get_int(ErlNifEnv *e, ERL_NIF_TERM t, int *i)
{
    (void)e;
    if (T(t)->tag != T_INT || T(t)->u.i > 2147483647LL || T(t)->u.i < -2147483648LL) return 0;
    *i = (int)T(t)->u.i;
    return 1;
}
int enif_get_tuple(ErlNifEnv *e, ERL_NIF_TERM t, int *arity, const ERL_NIF_TERM **array)
{
    (void)e;
    if (T(t)->tag != T_TUPLE) return 0;
    *arity = T(t)->u.tup.arity;
    *array = T(t)->u.tup.e;
    return 1;
}
int enif_get_list_cell(ErlNifEnv *e, ERL_NIF_TERM t, ERL_NIF_TERM *head, ERL_NIF_TERM *tail)
{
    (void)e;
    if (T(t)->tag != T_CONS) return 0;
    *head = T(t)->u.cons.head;
    *tail = T(t)->u.cons.tail;
    return 1;
}
int enif_get_list_length(ErlNifEnv *e, ERL_NIF_TERM t, unsigned *len)
{
    unsigned n = 0;
    (void)e;
    while (T(t)->tag == T_CONS) { n++; t = T(t)->u.cons.tail; }
    if (T(t)->tag != T_NIL) return 0;
    *len = n;
    return 1;
}
ERL_NIF_TERM enif_make_atom(ErlNifEnv *e, const char *name)
{
    ERL_NIF_TERM t = mk(T_ATOM);
    (void)e;
    strncpy(T(t)->u.atom, name, sizeof T(t)->u.atom - 1);
    return t;
}
ERL_NIF_TERM enif_make_int64(ErlNifEnv *e, ErlNifSInt64 i) { ERL_NIF_TERM t = mk(T_INT); (void)e; T(t)->u.i = i; return t; }
ERL_NIF_TERM enif_make_int(ErlNifEnv *e, int i) { return enif_make_int64(e, i); }
ERL_NIF_TERM enif_make_double(ErlNifEnv *e, double d) { ERL_NIF_TERM t = mk(T_FLOAT); (void)e; T(t)->u.f = d; return t; }
ERL_NIF_TERM enif_make_tuple(ErlNifEnv *e, unsigned cnt, ...)
{
    ERL_NIF_TERM t = mk(T_TUPLE);
    va_list ap;
    (void)e;
    T(t)->u.tup.arity = (int)cnt;
    T(t)->u.tup.e = calloc(cnt ? cnt : 1, sizeof(ERL_NIF_TERM));
    va_start(ap, cnt);
    for (unsigned i = 0; i < cnt; i++) T(t)->u.tup.e[i] = va_arg(ap, ERL_NIF_TERM);
    va_end(ap);
    return t;
}
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv *e, ERL_NIF_TERM head, ERL_NIF_TERM tail)
{
    ERL_NIF_TERM t = mk(T_CONS);
    (void)e;
    T(t)->u.cons.head = head;
    T(t)->u.cons.tail = tail;
    return t;
}
ERL_NIF_TERM enif_make_list(ErlNifEnv *e, unsigned cnt, ...)
{
    ERL_NIF_TERM items[64], t = mk(T_NIL);
    va_list ap;
    va_start(ap, cnt);
    for (unsigned i = 0; i < cnt && i < 64; i++) items[i] = va_arg(ap, ERL_NIF_TERM);
    va_end(ap);
    for (unsigned i = cnt; i-- > 0;) t = enif_make_list_cell(e, items[i], t);
    return t;
}
ERL_NIF_TERM enif_make_string(ErlNifEnv *e, const char *s, ErlNifCharEncoding enc)
{
    ERL_NIF_TERM t = mk(T_STR);
    (void)e; (void)enc;
    T(t)->u.str = strdup(s ? s : "");
    return t;
}
unsigned char *enif_make_new_binary(ErlNifEnv *e, size_t size, ERL_NIF_TERM *termp)
{
    ERL_NIF_TERM t = mk(T_BIN);
    (void)e;
    T(t)->u.bin.size = size;
    T(t)->u.bin.data = malloc(size ? size : 1);
    *termp = t;
    return T(t)->u.bin.data;
}
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *e) { (void)e; return mk(T_BADARG); }
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *e, const char *m, const char *name, ErlNifResourceDtor *dtor,
                                            ErlNifResourceFlags flags, ErlNifResourceFlags *tried)
{
    /* like OTP: CREATE alone fails for a type that exists, TAKEOVER alone for one that does not */
    static struct { char name[64]; ErlNifResourceType *rt; } known[8];
    ErlNifResourceType *rt;
    int k;
    (void)e; (void)m;
    for (k = 0; k < 8 && known[k].rt; k++) {
        if (strcmp(known[k].name, name) == 0) {
            if (!(flags & ERL_NIF_RT_TAKEOVER)) return NULL;
            known[k].rt->dtor = dtor;
            if (tried) *tried = ERL_NIF_RT_TAKEOVER;
            return known[k].rt;
        }
    }
    if (!(flags & ERL_NIF_RT_CREATE) || k == 8) return NULL;
    rt = calloc(1, sizeof *rt);
    rt->dtor = dtor;
    strncpy(known[k].name, name, sizeof known[k].name - 1);
    known[k].rt = rt;
    if (tried) *tried = ERL_NIF_RT_CREATE;
    return rt;
}
void *enif_alloc_resource(ErlNifResourceType *type, size_t size)
{
    res_hdr *h = calloc(1, sizeof(res_hdr) + size);
    h->type = type;
    h->refs = 1;
    return h + 1;
}
static void res_unref(void *obj)
{
    res_hdr *h = (res_hdr *)obj - 1;
    if (--h->refs == 0) {
        if (h->type->dtor) h->type->dtor(NULL, obj);
        free(h);
    }
}
void enif_release_resource(void *obj) { res_unref(obj); }
ERL_NIF_TERM enif_make_resource(ErlNifEnv *e, void *obj)
{
    ERL_NIF_TERM t = mk(T_RES);
    (void)e;
    ((res_hdr *)obj - 1)->refs++;
    T(t)->u.res = obj;
    return t;
}
int enif_get_resource(ErlNifEnv *e, ERL_NIF_TERM t, ErlNifResourceType *type, void **objp)
{
    (void)e;
    if (T(t)->tag != T_RES || !T(t)->u.res || ((res_hdr *)T(t)->u.res - 1)->type != type) return 0;
    *objp = T(t)->u.res;
    return 1;
}
ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv *e, void *obj, const void *data, size_t size)
{
    /* the bytes stay where they are; the term holds a reference on the resource (dropped by mock_resource_gc) */
    ERL_NIF_TERM t = mk(T_BIN);
    (void)e;
    ((res_hdr *)obj - 1)->refs++;
    T(t)->u.bin.size = size;
    T(t)->u.bin.data = (unsigned char *)data;
    T(t)->owner = obj;
    return t;
}
int mock_is_badarg(ERL_NIF_TERM t) { return T(t)->tag == T_BADARG; }
int mock_is_binary(ERL_NIF_TERM t, const unsigned char **data, size_t *size)
{
    if (T(t)->tag != T_BIN) return 0;
    *data = T(t)->u.bin.data;
    *size = T(t)->u.bin.size;
    return 1;
}
int mock_is_string(ERL_NIF_TERM t, const char **s)
{
    if (T(t)->tag != T_STR) return 0;
    *s = T(t)->u.str;
    return 1;
}
void mock_resource_gc(ERL_NIF_TERM t)
{
    if (T(t)->tag == T_RES && T(t)->u.res) {
        res_unref(T(t)->u.res);
        T(t)->u.res = NULL;
    } else if (T(t)->tag == T_BIN && T(t)->owner) {
        res_unref(T(t)->owner);
        T(t)->owner = NULL;
        T(t)->u.bin.data = NULL;
        T(t)->u.bin.size = 0;
    }
}
