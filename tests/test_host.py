"""CPU tests of the host side: the C-ABI library loads and exports what include/ert_b200.h
declares, struct layouts match, scene flattening, the writer, the partition logic.
No compute call is made here (there is no GPU in this container and no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import eraytracer_b200 as ert
from eraytracer_b200 import _lib, multigpu, ppm, raytracer
from eraytracer_b200 import scene as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ert_b200.h")


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"ERT_API\s+(?:const\s+char\s*\*|int)\s*(ert_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    L = ert.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libert_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names
    assert L.ert_abi_version() == 3


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "ert_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ert_material), sizeof(ert_camera),
         sizeof(ert_point_light), sizeof(ert_sphere), sizeof(ert_triangle), sizeof(ert_plane),
         sizeof(ert_scene_desc), sizeof(ert_render_params), sizeof(ert_stats));
  printf("%zu %zu %zu %zu\n", offsetof(ert_sphere, order), offsetof(ert_triangle, order),
         offsetof(ert_render_params, camera), offsetof(ert_stats, accel_used));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = list(map(int, out))
    assert sizes[:9] == [_lib.MATERIAL_DT.itemsize, ctypes.sizeof(_lib.Camera), _lib.LIGHT_DT.itemsize,
                         _lib.SPHERE_DT.itemsize, _lib.TRIANGLE_DT.itemsize, _lib.PLANE_DT.itemsize,
                         ctypes.sizeof(_lib.SceneDesc), ctypes.sizeof(_lib.RenderParams),
                         ctypes.sizeof(_lib.Stats)]
    assert sizes[9] == _lib.SPHERE_DT.fields['order'][1]
    assert sizes[10] == _lib.TRIANGLE_DT.fields['order'][1]
    assert sizes[11] == _lib.RenderParams.camera.offset
    assert sizes[12] == _lib.Stats.accel_used.offset


def _has_gpu():
    try:
        return ert.device_count() > 0
    except ert.ErtError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly_instead_of_falling_back():
    flat = sc.flatten(sc.demo_scene())
    with pytest.raises(ert.ErtError) as e:
        flat.upload(0)
    assert e.value.code in (_lib.ERT_ERR_NO_DEVICE, _lib.ERT_ERR_CUDA)
    with pytest.raises(ert.ErtError):
        raytracer.raytraced_pixel_list_gpu(4, 3, sc.demo_scene(), 1)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under eraytracer_b200/ may import, link,
    open or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "eraytracer_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|oracle\.|liboracle|pyoracle|\borc\b)")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn)).read()
                m = pat.search(text)
                assert m is None, "%s refers to the oracle: %r" % (os.path.join(dirpath, fn), m.group(0))


def test_flatten_demo_scene():
    flat = sc.flatten(sc.demo_scene())
    assert (len(flat.lights), len(flat.spheres), len(flat.triangles), len(flat.planes)) == (2, 3, 1, 1)
    assert flat.lights['order'].tolist() == [0, 1]
    assert flat.spheres['order'].tolist() == [2, 3, 4]
    assert flat.triangles['order'].tolist() == [5] and flat.planes['order'].tolist() == [6]
    assert flat.spheres['radius'].tolist() == [4.0, 4.0, 4.0]          # integers are promoted
    assert flat.spheres['center'][2].tolist() == [-4.5, -2.5, 14.0]
    assert flat.spheres['material']['reflectivity'].tolist() == [0.1, 0.5, 0.7]
    assert flat.planes['normal'][0].tolist() == [0.0, -1.0, 0.0] and flat.planes['distance'][0] == 5.0
    assert list(flat.camera.location) == [0.0, 0.0, -2.0] and flat.camera.fov == 90.0
    assert (flat.camera.screen_width, flat.camera.screen_height) == (4.0, 3.0)


def test_flatten_skips_unknown_elements_and_keeps_list_positions():
    s = sc.demo_scene()
    s.insert(3, ('fog', 40))
    s.append('an_atom')
    flat = sc.flatten(s)
    assert flat.spheres['order'].tolist() == [3, 4, 5]     # position 2 is the unknown tuple
    assert flat.planes['order'].tolist() == [7]


@pytest.mark.parametrize("bad", [
    [],
    [('sphere', 1, ('vector', 0, 0, 0), ('material', ('colour', 1, 1, 1), 1, 0, 0))],   # no camera first
    [sc.demo_scene()[0], ('sphere', 'undefined', ('vector', 0, 0, 0), ('material', ('colour', 1, 1, 1), 1, 0, 0))],
    [sc.demo_scene()[0], ('sphere', 3, ('vector', 0, 0, 10), ('material', ('colour', 0.4, 0.4, 0.4), 'undefined', 'undefined', 'undefined'))],
    [sc.demo_scene()[0], ('plane', ('vector', 0, 1), 5, ('material', ('colour', 1, 1, 1), 1, 0, 0))],
])
def test_flatten_rejects_malformed_scenes_with_badarg(bad):
    with pytest.raises(ert.BadArg):
        sc.flatten(bad)


def test_synthetic_scene_generator():
    a = sc.synthetic_scene("c3")
    b = sc.synthetic_scene("c3")
    assert len(a.spheres) == 10_000 and len(a.lights) == 3 and len(a.planes) == 1
    assert np.array_equal(a.spheres, b.spheres)                      # deterministic
    c = a.spheres['center']
    assert c[:, 0].min() >= -40 and c[:, 0].max() <= 40
    assert c[:, 1].min() >= -30 and c[:, 1].max() <= 4.0001
    assert c[:, 2].min() >= 5 and c[:, 2].max() <= 85
    assert a.spheres['radius'].min() >= 0.2 - 1e-6 and a.spheres['radius'].max() <= 0.8 + 1e-6
    assert ((c[:, 1] + a.spheres['radius']) <= 5).all()              # nothing pokes through the floor
    for field in ('radius',):
        v = a.spheres[field]
        assert np.array_equal(v, v.astype(np.float32).astype(np.float64))   # float32-exact inputs
    assert np.array_equal(c, c.astype(np.float32).astype(np.float64))
    assert set(np.unique(a.spheres['material']['specular_power'])) <= {1.0, 4.0, 20.0, 50.0}
    assert a.spheres['order'].tolist() == list(range(3, 10_003)) and a.planes['order'][0] == 10_003
    # first splitmix64 draw of the documented seed
    u = sc.splitmix64_uniform(0xE7A9C0DE00000003, 3)
    assert 0 <= u.min() and u.max() < 1
    small = sc.synthetic_scene("c4", n_spheres=7)
    assert len(small.spheres) == 7


def test_splitmix64_matches_the_published_algorithm():
    # reference implementation of splitmix64 in plain Python integers
    def sm(seed, n):
        out, x, mask = [], seed, (1 << 64) - 1
        for _ in range(n):
            x = (x + 0x9E3779B97F4A7C15) & mask
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
            z = z ^ (z >> 31)
            out.append((z >> 11) * 2.0 ** -53)
        return out
    assert sc.splitmix64_uniform(1234567, 5).tolist() == sm(1234567, 5)


def test_write_pixels_to_ppm_layout(tmp_path, capsys):
    """raytracer.erl:668-685: header lines, then 'R G B ' per pixel on one line."""
    p = tmp_path / "t.ppm"
    pixels = [(0, (0.0, 0.5, 1.0)), (1, (2.3, 0.999999, -0.5)), (2, (0.2, 0.2, 0.2)), (3, (1, 0, 0))]
    assert ppm.write_pixels_to_ppm(2, 2, 255, pixels, str(p)) == 'ok'
    assert p.read_text() == "P3\n2 2\n255\n0 127 255 255 254 -127 51 51 51 255 0 0 "
    assert "file opened" in capsys.readouterr().out
    assert ppm.write_pixels_to_ppm(2, 2, 255, pixels, str(tmp_path / "no" / "dir.ppm")) == 'ok'
    assert "error opening file" in capsys.readouterr().out


def test_frame_writer_matches_the_list_writer(tmp_path):
    rng = np.random.default_rng(1)
    frame = rng.uniform(0, 1.4, size=(5, 7, 3))
    pixels = [(i, tuple(px)) for i, px in enumerate(frame.reshape(-1, 3).tolist())]
    a, b, c = tmp_path / "a.ppm", tmp_path / "b.ppm", tmp_path / "c.ppm"
    ppm.write_pixels_to_ppm(7, 5, 255, pixels, str(a))
    q = ppm.quantise(frame).astype(np.uint8)
    ppm.write_frame_to_ppm(q, str(b), "P3")
    assert a.read_bytes() == b.read_bytes()
    ppm.write_frame_to_ppm(q, str(c), "P6")
    assert c.read_bytes() == b"P6\n7 5\n255\n" + q.tobytes()


def test_tracing_function_and_guards():
    assert raytracer.tracing_function('gpu') is raytracer.raytraced_pixel_list_gpu
    assert raytracer.tracing_function('gpu_distributed') is raytracer.raytraced_pixel_list_gpu_distributed
    for cpu in ('simple', 'concurrent', 'distributed'):
        with pytest.raises(NotImplementedError):
            raytracer.tracing_function(cpu)
    with pytest.raises(ValueError):
        raytracer.tracing_function('nonsense')
    assert raytracer.raytraced_pixel_list_gpu(0, 0, None, None) == 'done'      # erl:86-87
    for w, h in ((0, 3), (-1, 2), (4, 0)):
        with pytest.raises(ValueError):                                        # guards erl:89
            raytracer.raytraced_pixel_list_gpu(w, h, sc.demo_scene(), 1)
    assert raytracer.scene() == sc.demo_scene()


@pytest.mark.parametrize("height,band_rows,n_parts", [(2160, 8, 8), (47, 5, 3), (24, 8, 4), (7, 16, 2),
                                                     (1080, 0, 1), (100, 8, 1)])
def test_part_rows_partition_the_frame(height, band_rows, n_parts):
    seen = np.zeros(height, dtype=int)
    for part in range(max(n_parts, 1)):
        rows = multigpu.part_rows(height, band_rows, n_parts, part)
        assert np.all(np.diff(rows) > 0)
        seen[rows] += 1
        if n_parts > 1 and band_rows > 0:
            assert np.all((rows // band_rows) % n_parts == part)
    assert np.all(seen == 1)


def test_default_band_rows_and_assembly():
    assert multigpu.default_band_rows(2160, 8) == 8
    assert multigpu.default_band_rows(2160, 1) == 0
    assert multigpu.default_band_rows(24, 8) == 1
    h, w, n, br = 37, 5, 3, 4
    full = np.arange(h * w * 3, dtype=np.float64).reshape(h, w, 3)
    parts = [full[multigpu.part_rows(h, br, n, p)] for p in range(n)]
    assert np.array_equal(multigpu.assemble_parts(h, w, br, n, parts, np.float64), full)


def test_pose_camera():
    c0, c16 = sc.pose_camera(0), sc.pose_camera(16)
    assert list(c0.location) == [0.0, -0.5, -2.0]
    assert abs(c16.location[0] - 4.0) < 1e-12 and abs(c16.location[2] + 3.0) < 1e-12


def test_bench_reference_arm_prints_one_json_line():
    """The driver's contract for `bench.py --impl reference`: exactly one JSON line on stdout
    (whatever libraries print goes to stderr), same metric and config keys as the GPU arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
