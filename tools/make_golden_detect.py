#!/usr/bin/env python3
"""Dump detector golden vectors from the REFERENCE ITSELF: AlvaAR's FeatureExtractor::detectFeaturePoints
(src/slam/src/feature_extractor.cpp compiled unmodified against the vendored OpenCV 4.5.5 into oracle/_ref/libalva_ref.so,
cv::setNumThreads(1), baseline dispatch) plus the float intermediate of one cell.  tests/golden/detect.npz is committed."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
R.ref_detect_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_int]


def main():
    R.ref_config(0, 1)
    d = {}
    for tag, (w, h, cs, seed, ncur) in {"a": (320, 240, 40, 1, 0), "b": (320, 240, 40, 2, 12), "c": (243, 201, 30, 3, 5)}.items():
        fr, _ = synth.make_frames(1, w, h, seed=seed, rgba=False)
        img = np.ascontiguousarray(fr[0])
        rng = np.random.default_rng(seed)
        cur = np.stack([rng.uniform(0, w - 1, ncur), rng.uniform(0, h - 1, ncur)], 1).astype(np.float32) if ncur else np.zeros((0, 2), np.float32)
        roi = np.array([20, 20, w - 40, h - 40], np.int32)
        out = np.zeros((1024, 2), np.float32)
        n = R.ref_detect_points(P(img), w, h, cs, P(cur), ncur, P(roi), 0.001, P(out), 1024)
        hm = np.zeros((cs, cs), np.float32)
        bl = np.zeros((cs, cs), np.uint8)
        R.ref_min_eig_cell(P(img), w, h, cs, cs, cs, P(hm), P(bl))
        d.update({f"{tag}_img": img, f"{tag}_cell": cs, f"{tag}_cur": cur, f"{tag}_roi": roi, f"{tag}_pts": out[:n].copy(),
                  f"{tag}_hmap11": hm, f"{tag}_blur11": bl})
        print(tag, "points", n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "detect.npz"), **d)


if __name__ == "__main__":
    main()
