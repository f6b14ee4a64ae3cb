import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import scene as sc
accels = sys.argv[1:] or ["bvh"]
flat = sc.flatten(sc.demo_scene())
dev = flat.upload(0)
for accel in accels:
    for depth in (1, 2, 3):
        frame, st = dev.render(32, 24, depth, fmt="f64", accel=accel)
        print(accel, depth, st["rays"], float(frame.sum()), flush=True)
dev.close()
