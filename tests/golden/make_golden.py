"""Generates tests/golden/demo_images.json from the CPU oracle (oracle/oracle.c).

The reference cannot run here (no Erlang), so these are not outputs of the reference: they
pin the oracle against regressions and give the GPU tests a fixture that needs no oracle
build.  `appendix_a` holds the values an INDEPENDENT restatement produced during the survey
(SURVEY.md Appendix A); make_golden.py refuses to write the file if the oracle disagrees.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import orc, pyoracle  # noqa: E402

APPENDIX_A = {
    "focal_length_90_4": 2.0000000000000004,
    "32x24_d1": {"nonblack": 449, "saturated": 44, "sums": [51449, 46290, 20930], "rays": 1708,
                 "tests": 8540,
                 "pixels": {"10,8": [140, 181, 0], "12,10": [174, 255, 20], "22,9": [18, 123, 130],
                            "24,12": [150, 255, 194], "21,6": [204, 109, 7], "16,14": [38, 0, 19],
                            "17,15": [88, 45, 0], "18,18": [44, 22, 0], "5,20": [103, 54, 3],
                            "28,20": [163, 163, 81], "31,23": [199, 199, 99], "13,23": [164, 164, 82],
                            "3,13": [165, 0, 11], "0,0": [0, 0, 0]}},
    "32x24_d5": {"nonblack": 456, "saturated": 108, "sums": [67026, 59986, 29232],
                 "rays_reference_equivalent": 10824,
                 "pixels": {"12,10": [255, 255, 209], "17,15": [255, 222, 1], "18,18": [222, 111, 0],
                            "5,20": [255, 187, 155], "13,23": [168, 168, 84], "10,8": [140, 181, 0],
                            "22,9": [18, 123, 130], "24,12": [150, 255, 194], "21,6": [204, 109, 7],
                            "16,14": [38, 0, 19], "28,20": [163, 163, 81], "31,23": [199, 199, 99],
                            "3,13": [165, 0, 11]}},
    "4x3_d5": [[0, 0, 0]] * 5 + [[125, 71, 0], [0, 0, 0], [0, 106, 105], [81, 0, 40],
                                 [255, 150, 25], [255, 196, 1], [37, 92, 73]],
}


def main():
    sc = pyoracle.scene()
    cam = orc.camera_array(sc[0])
    kind, f = orc.flatten(sc[1:])
    out = {"appendix_a": APPENDIX_A, "images": {}}
    for w, h, d in ((4, 3, 5), (16, 12, 1), (32, 24, 1), (32, 24, 5), (64, 48, 5)):
        rgb, rays, tests = orc.render(cam, kind, f, w, h, d)
        q = orc.quantise_image(rgb)
        out["images"]["%dx%d_d%d" % (w, h, d)] = {
            "width": w, "height": h, "depth": d, "rays": rays, "tests": tests,
            "rgb8": q.reshape(-1).tolist(),
            "f64_sha256": hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest()}
    # cross-check against the survey's independent restatement before writing
    a = APPENDIX_A
    for key in ("32x24_d1", "32x24_d5"):
        img = out["images"][key]
        q = np.array(img["rgb8"]).reshape(24, 32, 3)
        assert q.reshape(-1, 3).sum(axis=0).tolist() == a[key]["sums"], key
        for xy, rgbv in a[key]["pixels"].items():
            x, y = map(int, xy.split(","))
            assert q[y, x].tolist() == rgbv, (key, xy)
    assert np.array(out["images"]["4x3_d5"]["rgb8"]).reshape(-1, 3).tolist() == a["4x3_d5"]
    with open(os.path.join(HERE, "demo_images.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote demo_images.json")


if __name__ == "__main__":
    main()
