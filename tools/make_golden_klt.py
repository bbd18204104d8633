#!/usr/bin/env python3
"""Dump KLT golden vectors from the REFERENCE ITSELF: AlvaAR's own FeatureTracker::fbKltTracking
(src/slam/src/feature_tracker.cpp, compiled unmodified into oracle/_ref/libalva_ref.so) on pyramids built by the vendored
OpenCV 4.5.5's buildOpticalFlowPyramid, and cv::calcOpticalFlowPyrLK with the flags the reference uses.
Run in the build container only; tests/golden/klt.npz is committed."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from alvaar_b200 import synth  # noqa: E402
from klt_util import klt_points  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
R.ref_fb_klt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
R.ref_klt_lk.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                    C.c_void_p, C.c_int]


def main():
    R.ref_config(0, 1)
    w, h, n = 200, 150, 300
    fr, _ = synth.make_frames(2, w, h, seed=5, rgba=False)
    a, b = np.ascontiguousarray(fr[0]), np.ascontiguousarray(fr[1])
    pts, pri = klt_points(w, h, n, seed=5)
    d = {"prev": a, "cur": b, "pts": pts, "priors": pri}
    for levels in (1, 3):
        q = pri.copy()
        good = np.zeros(n, np.uint8)
        d["pyr_levels"] = R.ref_fb_klt(P(a), P(b), w, h, 9, 3, levels, 30.0, 0.5, P(pts), P(q), P(good), n)
        d[f"fb{levels}_pos"], d[f"fb{levels}_good"] = q, good
        for ui in (0, 1):
            nx, st, er = pri.copy(), np.zeros(n, np.uint8), np.zeros(n, np.float32)
            R.ref_klt_lk(P(a), P(b), w, h, 9, 3, levels, 30, 0.01, ui, P(pts), P(nx), P(st), P(er), n)
            d[f"lk{levels}_{ui}_pos"], d[f"lk{levels}_{ui}_status"], d[f"lk{levels}_{ui}_err"] = nx, st, er
        print("levels", levels, "good", int(good.sum()), "of", n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "klt.npz"), **d)


if __name__ == "__main__":
    main()
