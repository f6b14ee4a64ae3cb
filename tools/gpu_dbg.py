import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import scene as sc
for spec in sys.argv[1:]:
    w, h, d, n = map(int, spec.split(","))
    flat = sc.synthetic_scene("c3", n_spheres=n)
    dev = flat.upload(0)
    a, st = dev.render(w, h, d, fmt="f64", accel="linear")
    b, st2 = dev.render(w, h, d, fmt="f64", accel="bvh")
    bad = np.argwhere(np.abs(a - b).max(axis=2) > 0)
    print(spec, "linear vs bvh: max|d| =", np.abs(a - b).max(), "bad px", len(bad), bad[:6].tolist(), "rays", st["rays"], st2["rays"], flush=True)
    dev.close()
