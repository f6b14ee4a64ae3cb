"""BASELINE.json's configurations at their FULL sizes, through the path a user gets (accel="auto"), held to the
CPU oracle directly (crops and lattices where a whole frame is out of the oracle's reach).

  C3  10 k spheres, 3840x2160, depth 5: a 64x64 crop and a 48x27 lattice (5392 px) of the default frame
  C4  1 M spheres, 3840x2160, depth 5: a 32x32 lattice (1024 px, ~17 k rays x 1 M objects on the CPU)
  C5  demo scene 7680x4320 depth 1 (run-distributed.sh:2), 8 of the 64 camera poses: WHOLE frames against the
      oracle (33 M pixels each), plus the 2-part and 8-part row-band assembly at that size

Bar: RGB8 bit-equal, doubles within 1e-9 relative, and the north-star image metric.
"""
import numpy as np
import pytest

from eraytracer_b200 import _lib, multigpu
from eraytracer_b200 import scene as sc
from helpers import assert_double_parity, assert_image_parity, oracle_frame, quantise

pytestmark = pytest.mark.gpu
W4K, H4K = 3840, 2160


def _lattice(w, h, nx, ny):
    xs = ((np.arange(nx) + 0.5) * w / nx).astype(np.int32)
    ys = ((np.arange(ny) + 0.5) * h / ny).astype(np.int32)
    gx, gy = np.meshgrid(xs, ys)
    return gx.reshape(-1).astype(np.int32), gy.reshape(-1).astype(np.int32)


def _check_pixels(flat, full64, full8, xs, ys, depth):
    ref, _, _ = oracle_frame(flat, W4K, H4K, depth, pixels=(xs, ys))
    got = full64[ys, xs]
    assert_double_parity(got, ref)
    want8 = np.clip(quantise(ref), 0, 255)
    assert np.array_equal(quantise(got), quantise(ref))
    assert np.array_equal(full8[ys, xs].astype(np.int64), want8)
    assert_image_parity(full8[ys, xs], want8)


def test_c3_4k_default_path_against_the_oracle(gpu):
    flat = sc.synthetic_scene("c3")
    dev = flat.upload(0)
    depth = 5
    full64, st = dev.render(W4K, H4K, depth, fmt="f64", accel="auto")
    assert st["accel_used"] == "grid" and st["pixels"] == W4K * H4K
    full8, _ = dev.render(W4K, H4K, depth, fmt="rgb8", accel="auto")
    assert np.array_equal(full8.astype(np.int64), np.clip(quantise(full64), 0, 255))
    # a 64x64 crop in the busiest part of the frame, and a lattice over all of it
    cx, cy = np.meshgrid(np.arange(1888, 1952, dtype=np.int32), np.arange(1180, 1244, dtype=np.int32))
    _check_pixels(flat, full64, full8, cx.reshape(-1), cy.reshape(-1), depth)
    _check_pixels(flat, full64, full8, *_lattice(W4K, H4K, 48, 27), depth)
    # the other strategies give the same frame bit for bit
    bvh, sb = dev.render(W4K, H4K, depth, fmt="f64", accel="bvh")
    assert np.array_equal(bvh, full64) and sb["rays"] == st["rays"]
    dev.close()


def test_c4_4k_default_path_against_the_oracle(gpu):
    flat = sc.synthetic_scene("c4")
    dev = flat.upload(0)
    depth = 5
    full64, st = dev.render(W4K, H4K, depth, fmt="f64", accel="auto")
    assert st["accel_used"] == "grid" and st["pixels"] == W4K * H4K and st["rays"] > 8 * W4K * H4K
    full8, _ = dev.render(W4K, H4K, depth, fmt="rgb8", accel="auto")
    assert np.array_equal(full8.astype(np.int64), np.clip(quantise(full64), 0, 255))
    _check_pixels(flat, full64, full8, *_lattice(W4K, H4K, 32, 32), depth)          # 1024 pixels
    dev.close()


C5_W, C5_H, C5_DEPTH = 7680, 4320, 1
C5_POSES = (0, 9, 18, 27, 36, 45, 54, 63)


def test_c5_8k_pose_frames_equal_the_oracle(gpu):
    flat = sc.flatten(sc.demo_scene())
    dev = flat.upload(0)
    for k in C5_POSES:
        cam = sc.pose_camera(k)
        flat_k = sc.FlatScene(cam, flat.lights, flat.spheres, flat.triangles, flat.planes)
        ref, _, _ = oracle_frame(flat_k, C5_W, C5_H, C5_DEPTH)
        want8 = np.clip(quantise(ref), 0, 255).astype(np.uint8)
        got8, st = dev.render(C5_W, C5_H, C5_DEPTH, fmt="rgb8", camera=cam)
        assert st["pixels"] == C5_W * C5_H
        assert np.array_equal(got8, want8), "pose %d: %d pixels differ" % (k, int((got8 != want8).any(axis=2).sum()))
        if k in (0, 36):
            got64, _ = dev.render(C5_W, C5_H, C5_DEPTH, fmt="f64", camera=cam)
            assert_double_parity(got64, ref)
            del got64
        del ref, want8, got8
    dev.close()


@pytest.mark.parametrize("n_parts", (2, 8))
def test_c5_8k_row_band_parts_assemble_the_frame(gpu, n_parts):
    dev = sc.flatten(sc.demo_scene()).upload(0)
    cam = sc.pose_camera(21)
    whole, st = dev.render(C5_W, C5_H, C5_DEPTH, fmt="rgb8", camera=cam)
    band_rows = multigpu.default_band_rows(C5_H, n_parts)
    frame = _lib.PinnedFrame(C5_W * C5_H * 3)
    frame.array(np.uint8, (C5_H, C5_W, 3))[:] = 0
    rays = 0
    for part in range(n_parts):
        dev.render_async(C5_W, C5_H, C5_DEPTH, slot=part % _lib.MAX_SLOTS, fmt="rgb8", camera=cam, band_rows=band_rows,
                         n_parts=n_parts, part=part, host_ptr=frame.ptr, host_bytes=C5_W * C5_H * 3)
        dev.wait(part % _lib.MAX_SLOTS)
        rays += dev.stats(part % _lib.MAX_SLOTS)["rays"]
    assert np.array_equal(frame.array(np.uint8, (C5_H, C5_W, 3)), whole)
    assert rays == st["rays"]
    frame.close()
    dev.close()


@pytest.mark.parametrize("kind", ("c3", "c4"))
def test_4k_shadow_triage_equals_the_literal_shadow_path(gpu, kind):
    """Every shadow ray of the full-size frames (76 M on C3, 102 M on C4): the default path (direction grids, FP32
    triage, packed literal path, fused fold) against shadow rays that walk the BVH on the literal path."""
    flat = sc.synthetic_scene(kind)
    dev = flat.upload(0)
    a, sa = dev.render(W4K, H4K, 5, fmt="rgb8", accel="auto")
    b, sb = dev.render(W4K, H4K, 5, fmt="rgb8", accel="auto", flags=_lib.FLAG_NO_LIGHT_GRID)
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]
    # and the unclamped doubles of every fourth row
    a64 = np.zeros((H4K, W4K, 3), dtype=np.float64)
    b64 = np.zeros((H4K, W4K, 3), dtype=np.float64)
    dev.render(W4K, H4K, 5, fmt="f64", accel="auto", band_rows=1, n_parts=4, part=1, out=a64)
    dev.render(W4K, H4K, 5, fmt="f64", accel="auto", band_rows=1, n_parts=4, part=1, flags=_lib.FLAG_NO_LIGHT_GRID, out=b64)
    assert np.array_equal(a64, b64) and np.any(a64[1::4] != 0)
    dev.close()
