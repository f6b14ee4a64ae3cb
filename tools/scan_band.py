"""Renders a 4-row band of a workload with the brute-force scan (accel=linear): the driver for ncu captures of wf_scan."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraytracer_b200 import scene as sc, _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flat = sc.synthetic_scene(kind)
dev = flat.upload(0)
w, h, depth = 3840, 2160, 5
for k in range(reps):
    dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="linear", band_rows=4, n_parts=h // 4, part=h // 8,
                     flags=_lib.FLAG_TIME_KERNELS)
    dev.wait(0)
    st = dev.stats(0)
    print("band %s: kernel %.2f ms, path %.2f ms, shadow %.2f ms, rays %d" % (kind, st["kernel_ms"], st["path_ms"], st["shadow_ms"], st["rays"]), flush=True)
dev.close()
