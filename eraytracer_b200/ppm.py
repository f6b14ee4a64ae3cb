"""PPM output: the reference's writer and a binary-framebuffer writer.

write_pixels_to_ppm/5 (raytracer.erl:668-685) consumes the W*H-element pixel list with
one io:format per pixel; it defines the quantisation rule that image parity is judged on
(min(trunc(C*MaxValue), MaxValue), no lower clamp, erl:678-680).  write_frame_to_ppm takes
the RGB8 framebuffer the GPU produces and writes the same bytes without the list.
"""
import math

import numpy as np


def write_pixels_to_ppm(width, height, max_value, pixels, filename):
    """Byte-for-byte what raytracer.erl:668-685 writes: 'P3', 'W H', 'Max' on their own
    lines, then every pixel as 'R G B ' on one line with no trailing newline."""
    try:
        f = open(filename, "w", newline="\n")
    except OSError:
        print("error opening file")                  # erl:683-684
        return 'ok'
    with f:
        print("file opened")
        f.write("P3\n")
        f.write("%d %d\n" % (width, height))
        f.write("%d\n" % max_value)
        out = []
        for _num, (r, g, b) in pixels:
            out.append("%d %d %d " % (min(math.trunc(r * max_value), max_value),
                                      min(math.trunc(g * max_value), max_value),
                                      min(math.trunc(b * max_value), max_value)))
        f.write("".join(out))
    return 'ok'


def quantise(frame, max_value=255):
    """erl:678-680 on an unclamped float frame -> int64 array (negative values kept)."""
    return np.minimum(np.trunc(np.asarray(frame, dtype=np.float64) * max_value),
                      max_value).astype(np.int64)


def write_frame_to_ppm(frame_rgb8, filename, kind="P3"):
    """Writes an (H, W, 3) uint8 frame.  kind='P3' gives the reference's exact text layout
    (for frames without negative channels); 'P6' is the binary format for large frames."""
    frame = np.ascontiguousarray(frame_rgb8, dtype=np.uint8)
    h, w, _ = frame.shape
    if kind == "P6":
        with open(filename, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h))
            f.write(frame.tobytes())
        return 'ok'
    if kind != "P3":
        raise ValueError("kind must be 'P3' or 'P6'")
    lut = np.array([("%d " % v).encode() for v in range(256)], dtype=object)
    with open(filename, "wb") as f:
        f.write(b"P3\n%d %d\n255\n" % (w, h))
        flat = frame.reshape(-1)
        step = 3 * 65536
        for i in range(0, len(flat), step):
            f.write(b"".join(lut[flat[i:i + step]].tolist()))
    return 'ok'
