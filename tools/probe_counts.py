"""Development probe: per-depth test counters of the path kernels (library built with -DERT_PROBE prints them)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraytracer_b200 import scene as sc, _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
depths = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 5]
flat = sc.synthetic_scene(kind)
dev = flat.upload(0)
for depth in depths:
    sys.stderr.write("depth %d\n" % depth); sys.stderr.flush()
    dev.render_async(3840, 2160, depth, slot=0, fmt="rgb8", accel="auto", flags=_lib.FLAG_COUNT_TESTS)
    dev.wait(0)
    st = dev.stats(0)
    print(depth, st["rays"], st["bounce_path_rays"][:depth], st["bounce_hits"][:depth], flush=True)
dev.close()
