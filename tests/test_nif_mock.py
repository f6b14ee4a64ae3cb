"""The NIF glue (c_src/raytracer_gpu_nif.c) compiled against the mock erl_nif and executed:
term decode (the wire format scene_test pins, raytracer.erl:760-801), badarg behaviour,
and — on the GPU box — upload/render through the NIF entry points."""
import numpy as np
import pytest

import eraytracer_b200 as ert
from eraytracer_b200 import scene as sc
from nif_mock import Badarg, MockBeam, Resource


@pytest.fixture(scope="module")
def beam():
    ert.load()
    return MockBeam()


def test_nif_table(beam):
    assert beam.entry.name == b"raytracer_gpu"
    table = {k: f.flags for k, f in beam.funcs.items()}
    assert set(table) == {("device_count", 0), ("scene_info", 1), ("scene_upload", 2), ("scene_clone", 2), ("render", 5),
                          ("render_pixel_list", 5), ("frame_alloc", 3), ("render_into", 4), ("frame_binary", 1)}
    # everything that can block runs on a dirty scheduler: the host-side build is CPU-bound (1), waiting for the
    # GPU is IO-bound (2)
    assert table[("scene_upload", 2)] == 1
    for name in (("scene_clone", 2), ("render", 5), ("render_pixel_list", 5), ("render_into", 4), ("frame_alloc", 3)):
        assert table[name] == 2, name


def test_hot_code_upgrade_takes_the_resource_types_over(beam):
    # load() ran once in the fixture: the types exist.  A second plain load must fail (ERL_NIF_RT_CREATE on an
    # existing type), upgrade() must succeed (CREATE | TAKEOVER).
    import ctypes
    assert beam.entry.load(beam.env, None, 0) != 0
    up = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)(beam.entry.upgrade)
    assert up(beam.env, None, None, 0) == 0


def test_a_million_element_scene_is_decoded_with_exact_tables(beam):
    # decode_scene sizes its tables by a counting pass; a list of junk must not allocate per-kind tables of its length
    n = 200000
    scene = [sc.demo_scene()[0]] + [('fog', k) for k in range(n)] + sc.demo_scene()[1:]
    assert beam.call("scene_info", scene) == (2, 3, 1, 1)


def test_scene_decode_accepts_the_reference_records(beam):
    assert beam.call("scene_info", sc.demo_scene()) == (2, 3, 1, 1)
    floats = [tuple(float(x) if isinstance(x, int) else x for x in e) if e[0] == 'sphere' else e
              for e in sc.demo_scene()]
    assert beam.call("scene_info", floats) == (2, 3, 1, 1)
    with_junk = sc.demo_scene() + [('fog', 40), 'atom', 7, ('sphere', 1)]
    assert beam.call("scene_info", with_junk) == (2, 3, 1, 1)
    assert beam.call("scene_info", [sc.demo_scene()[0]]) == (0, 0, 0, 0)


@pytest.mark.parametrize("bad", [
    [],                                                     # no camera
    'not_a_list',
    [('sphere', 1, ('vector', 0, 0, 0), ('material', ('colour', 1, 1, 1), 1, 0, 0))],
    [sc.demo_scene()[0], ('sphere', 'undefined', ('vector', 0, 0, 0), ('material', ('colour', 1, 1, 1), 1, 0, 0))],
    [sc.demo_scene()[0], ('sphere', 3, ('vector', 0, 0, 10),
                          ('material', ('colour', 0.4, 0.4, 0.4), 'undefined', 'undefined', 'undefined'))],
    [sc.demo_scene()[0], ('plane', ('vector', 0, 1), 5, ('material', ('colour', 1, 1, 1), 1, 0, 0))],
    [sc.demo_scene()[0], ('point_light', ('colour', 1, 1, 1), ('colour', 0, 0, 0), ('colour', 1, 1, 1))],
    [('camera', ('vector', 0, 0, -2), ('vector', 0, 0, 0), 90, ('screen', 4))],
])
def test_malformed_scenes_raise_badarg(beam, bad):
    with pytest.raises(Badarg):
        beam.call("scene_info", bad)
    with pytest.raises(Badarg):
        beam.call("scene_upload", bad, 0)


def _has_gpu():
    try:
        return ert.device_count() > 0
    except ert.ErtError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_upload_without_a_gpu_is_an_error_tuple(beam):
    res = beam.call("scene_upload", sc.demo_scene(), 0)
    assert res[0] == "error" and res[1][0] in ("no_device", "cuda") and isinstance(res[1][2], str)
    res = beam.call("device_count")
    assert res[0] == "error"


@pytest.mark.gpu
def test_render_through_the_nif_matches_the_c_abi(beam, gpu):
    ok, n = beam.call("device_count")
    assert ok == "ok" and n >= 1
    ok, handle = beam.call("scene_upload", sc.demo_scene(), 0)
    assert ok == "ok" and isinstance(handle, Resource)
    dev = sc.flatten(sc.demo_scene()).upload(0)
    w, h, depth = 48, 36, 5
    want8, _ = dev.render(w, h, depth, fmt="rgb8")
    want64, _ = dev.render(w, h, depth, fmt="f64")
    ok, frame = beam.call("render", handle, w, h, depth, [])
    assert ok == "ok" and frame == want8.tobytes()
    ok, frame = beam.call("render", handle, w, h, depth, [("format", "f64"), ("accel", "bvh")])
    assert frame == want64.tobytes()
    pixels = beam.call("render_pixel_list", handle, w, h, depth, [])
    assert [p[0] for p in pixels] == list(range(w * h))
    assert np.array_equal(np.array([p[1] for p in pixels]).reshape(h, w, 3), want64)
    # row-band parts and a per-call camera
    cam = ('camera', ('vector', 1, -1, -3), ('vector', 0, 0, 0), 90, ('screen', 4, 3))
    parts = []
    for part in range(3):
        ok, fr = beam.call("render", handle, w, h, 1, [("part", (4, 3, part)), ("camera", cam)])
        parts.append(np.frombuffer(fr, dtype=np.uint8).reshape(h, w, 3))
    stitched = np.zeros((h, w, 3), dtype=np.uint8)
    for part in range(3):
        rows = [y for y in range(h) if (y // 4) % 3 == part]
        stitched[rows] = parts[part][rows]
    want_cam, _ = dev.render(w, h, 1, fmt="rgb8", camera=sc.camera_struct(cam))
    assert np.array_equal(stitched, want_cam)
    # guards of erl:89 and unknown options
    for args in ((handle, 0, 4, 1, []), (handle, 4, -1, 1, []), (handle, 4, 4, -1, []),
                 (handle, 4, 4, 1, [("format", "jpeg")]), (handle, 4, 4, 1, [("part", (4, 2, 5))]),
                 ("not_a_handle", 4, 4, 1, [])):
        with pytest.raises(Badarg):
            beam.call("render", *args)
    # one shared page-locked frame filled part by part through a clone of the scene (what render_binary/5 does)
    ok, clone = beam.call("scene_clone", handle, 0)
    assert ok == "ok" and isinstance(clone, Resource)
    ok, frame = beam.call("frame_alloc", w, h, "rgb8")
    assert ok == "ok" and isinstance(frame, Resource)
    for part, hdl in ((0, handle), (1, clone), (2, handle)):
        assert beam.call("render_into", hdl, frame, 1, [("part", (4, 3, part)), ("camera", cam)]) == "ok"
    assert beam.call("frame_binary", frame) == want_cam.tobytes()
    with pytest.raises(Badarg):
        beam.call("render_into", handle, frame, 1, [("format", "f64")])        # the frame is rgb8
    with pytest.raises(Badarg):
        beam.call("scene_clone", "not_a_handle", 0)
    beam.gc(clone)
    beam.gc(frame)
    beam.gc(handle)          # the resource destructor frees the device scene
    with pytest.raises(Badarg):
        beam.call("render", handle, 4, 4, 1, [])
    dev.close()
