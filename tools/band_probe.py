"""Kernel time of every part of an N-part C4 frame for several band heights (one GPU renders the parts in turn).

usage: python tools/band_probe.py [n_parts] [band_rows ...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eraytracer_b200 import _lib, multigpu, scene as sc

n_parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bands = [int(x) for x in sys.argv[2:]] or [multigpu.default_band_rows(2160, n_parts)]
w, h, depth = 3840, 2160, 5
dev = sc.synthetic_scene("c4").upload(0)
for band in bands:
    per_part = []
    for part in range(n_parts):
        ms = []
        for i in range(6):
            _lib.l2_flush(0)
            dev.render_async(w, h, depth, slot=0, fmt="rgb8", accel="auto", band_rows=band, n_parts=n_parts, part=part)
            dev.wait(0)
            if i >= 2:
                ms.append(dev.stats(0)["kernel_ms"])
        per_part.append(float(np.median(ms)))
    print("band %4d rows, %d parts: max %.3f mean %.3f ms  %s" % (band, n_parts, max(per_part), float(np.mean(per_part)),
          " ".join("%.2f" % x for x in per_part)), flush=True)
dev.close()
