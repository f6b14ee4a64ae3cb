"""Drives c_src/raytracer_gpu_nif.c through the mock erl_nif (tests/mock_erl) with ctypes."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "mock_erl", "_build")
SO = os.path.join(BUILD, "raytracer_gpu_mock.so")


class ErlNifFunc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("arity", ctypes.c_uint),
                ("fptr", ctypes.CFUNCTYPE(ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_size_t))),
                ("flags", ctypes.c_uint)]


class ErlNifEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("num_of_funcs", ctypes.c_int),
                ("funcs", ctypes.POINTER(ErlNifFunc)),
                ("load", ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)),
                ("reload", ctypes.c_void_p), ("upgrade", ctypes.c_void_p), ("unload", ctypes.c_void_p)]


class Badarg(Exception):
    pass


class Atom(str):
    pass


class Resource:
    def __init__(self, term):
        self.term = term


class MockBeam:
    def __init__(self):
        os.makedirs(BUILD, exist_ok=True)
        lib_dir = os.path.join(ROOT, "eraytracer_b200", "lib")
        srcs = [os.path.join(ROOT, "c_src", "raytracer_gpu_nif.c"),
                os.path.join(ROOT, "tests", "mock_erl", "mock_erl_nif.c")]
        if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
            subprocess.check_call(
                ["gcc", "-Wall", "-Wextra", "-Werror", "-O1", "-fPIC", "-shared",
                 "-I" + os.path.join(ROOT, "tests", "mock_erl"), "-I" + os.path.join(ROOT, "include")]
                + srcs + ["-L" + lib_dir, "-lert_b200", "-Wl,-rpath," + lib_dir, "-o", SO])
        L = self.L = ctypes.CDLL(SO)
        T = ctypes.c_size_t
        L.mock_env_new.restype = ctypes.c_void_p
        L.nif_init.restype = ctypes.POINTER(ErlNifEntry)
        for name, args in (("enif_make_atom", [ctypes.c_void_p, ctypes.c_char_p]),
                           ("enif_make_int64", [ctypes.c_void_p, ctypes.c_int64]),
                           ("enif_make_double", [ctypes.c_void_p, ctypes.c_double]),
                           ("enif_make_list_cell", [ctypes.c_void_p, T, T])):
            getattr(L, name).restype = T
            getattr(L, name).argtypes = args
        L.enif_make_tuple.restype = T
        L.enif_make_list.restype = T
        L.enif_get_tuple.argtypes = [ctypes.c_void_p, T, ctypes.POINTER(ctypes.c_int),
                                     ctypes.POINTER(ctypes.POINTER(T))]
        L.enif_get_list_cell.argtypes = [ctypes.c_void_p, T, ctypes.POINTER(T), ctypes.POINTER(T)]
        L.enif_get_atom.argtypes = [ctypes.c_void_p, T, ctypes.c_char_p, ctypes.c_uint, ctypes.c_int]
        L.enif_get_double.argtypes = [ctypes.c_void_p, T, ctypes.POINTER(ctypes.c_double)]
        L.enif_get_int64.argtypes = [ctypes.c_void_p, T, ctypes.POINTER(ctypes.c_int64)]
        L.enif_is_empty_list.argtypes = [ctypes.c_void_p, T]
        L.enif_is_atom.argtypes = [ctypes.c_void_p, T]
        L.mock_is_badarg.argtypes = [T]
        L.mock_is_binary.argtypes = [T, ctypes.POINTER(ctypes.POINTER(ctypes.c_ubyte)),
                                     ctypes.POINTER(ctypes.c_size_t)]
        L.mock_is_string.argtypes = [T, ctypes.POINTER(ctypes.c_char_p)]
        L.mock_resource_gc.argtypes = [T]
        self.env = L.mock_env_new()
        self.entry = L.nif_init().contents
        assert self.entry.load(self.env, None, 0) == 0
        self.funcs = {}
        for i in range(self.entry.num_of_funcs):
            f = self.entry.funcs[i]
            self.funcs[(f.name.decode(), f.arity)] = f

    # ---- Python -> term
    def term(self, v):
        L, env = self.L, self.env
        if isinstance(v, Resource):
            return v.term
        if isinstance(v, bool):
            return L.enif_make_atom(env, b"true" if v else b"false")
        if isinstance(v, str):
            return L.enif_make_atom(env, v.encode())
        if isinstance(v, int):
            return L.enif_make_int64(env, v)
        if isinstance(v, float):
            return L.enif_make_double(env, v)
        if isinstance(v, tuple):
            elems = [ctypes.c_size_t(self.term(e)) for e in v]
            return L.enif_make_tuple(ctypes.c_void_p(env), ctypes.c_uint(len(elems)), *elems)
        if isinstance(v, list):
            t = L.enif_make_list(ctypes.c_void_p(env), ctypes.c_uint(0))
            for e in reversed(v):
                t = L.enif_make_list_cell(env, self.term(e), t)
            return t
        raise TypeError(v)

    # ---- term -> Python
    def value(self, t):
        L, env = self.L, self.env
        if L.mock_is_badarg(t):
            raise Badarg()
        d = ctypes.c_double()
        if L.enif_get_double(env, t, ctypes.byref(d)):
            return d.value
        i = ctypes.c_int64()
        if L.enif_get_int64(env, t, ctypes.byref(i)):
            return i.value
        if L.enif_is_atom(env, t):
            buf = ctypes.create_string_buffer(64)
            L.enif_get_atom(env, t, buf, 64, 1)
            return Atom(buf.value.decode())
        arity = ctypes.c_int()
        arr = ctypes.POINTER(ctypes.c_size_t)()
        if L.enif_get_tuple(env, t, ctypes.byref(arity), ctypes.byref(arr)):
            return tuple(self.value(arr[k]) for k in range(arity.value))
        if L.enif_is_empty_list(env, t):
            return []
        h, tl = ctypes.c_size_t(), ctypes.c_size_t()
        if L.enif_get_list_cell(env, t, ctypes.byref(h), ctypes.byref(tl)):
            out = []
            while True:
                out.append(self.value(h.value))
                t = tl.value
                if not L.enif_get_list_cell(env, t, ctypes.byref(h), ctypes.byref(tl)):
                    break
            return out
        data = ctypes.POINTER(ctypes.c_ubyte)()
        size = ctypes.c_size_t()
        if L.mock_is_binary(t, ctypes.byref(data), ctypes.byref(size)):
            return ctypes.string_at(data, size.value)
        s = ctypes.c_char_p()
        if L.mock_is_string(t, ctypes.byref(s)):
            return s.value.decode()
        return Resource(t)

    def call(self, name, *args):
        f = self.funcs[(name, len(args))]
        argv = (ctypes.c_size_t * max(len(args), 1))(*[self.term(a) for a in args])
        return self.value(f.fptr(self.env, len(args), argv))

    def gc(self, resource):
        self.L.mock_resource_gc(resource.term)
