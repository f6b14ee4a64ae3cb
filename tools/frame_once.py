"""Renders one workload frame a few times (the driver for ncu captures): python tools/frame_once.py [workload] [reps] [accel]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraytracer_b200 import scene as sc, _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
accel = sys.argv[3] if len(sys.argv) > 3 else "auto"
flat = sc.synthetic_scene(kind)
dev = flat.upload(0)
for k in range(reps):
    _lib.l2_flush(0)
    dev.render_async(3840, 2160, 5, slot=0, fmt="rgb8", accel=accel, flags=_lib.FLAG_TIME_KERNELS)
    dev.wait(0)
    st = dev.stats(0)
    print("%s frame: kernel %.3f ms, path %.3f, shadow %.3f, other %.3f, launches %d" % (kind, st["kernel_ms"], st["path_ms"], st["shadow_ms"], st["other_ms"], st["gpu_launches"]), flush=True)
dev.close()
