#!/usr/bin/env python3
"""Dump pose golden vectors from the REFERENCE ITSELF: AlvaAR's MultiViewGeometry::p3pRansac / ceresPnP
(src/slam/src/multi_view_geometry.cpp compiled unmodified with the vendored OpenGV 1.0 / Ceres 2.0 into
oracle/_ref/libalva_ref.so; sampler seed pinned through doRandom = false).  tests/golden/pose.npz is committed."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pose_util import make_pose_problem  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
f32 = C.c_float
R.ref_p3p_lmeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, f32, f32, f32, C.c_void_p, C.c_void_p]
R.ref_pnp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, f32, C.c_int, C.c_int, f32, f32, f32, f32, C.c_void_p]


def main():
    R.ref_config(0, 1)
    d = {}
    for tag, (n, seed, of) in {"a": (300, 0, 0.1), "b": (576, 1, 0.3), "c": (40, 2, 0.0)}.items():
        pr = make_pose_problem(n, seed, outlier_frac=of)
        K32 = pr["K"].astype(np.float32)
        T = np.zeros(12)
        o = np.zeros(n, np.uint8)
        ok = R.ref_p3p_lmeds(P(pr["bv"]), P(pr["X"]), n, 100, 3.0, K32[0], K32[1], P(T), P(o))
        d.update({f"{tag}_bv": pr["bv"], f"{tag}_X": pr["X"], f"{tag}_uv": pr["uv"], f"{tag}_K": K32, f"{tag}_pose0": pr["pose0"],
                  f"{tag}_p3p_ok": ok, f"{tag}_p3p_T": T, f"{tag}_p3p_outlier": o})
        for rob, l2 in ((1, 1), (1, 0), (0, 0)):
            p = pr["pose0"].copy()
            oo = np.zeros(n, np.uint8)
            k = R.ref_pnp(P(pr["uv"]), P(pr["X"]), n, P(p), 5, 5.9915, rob, l2, K32[0], K32[1], K32[2], K32[3], P(oo))
            d.update({f"{tag}_pnp{rob}{l2}_ok": k, f"{tag}_pnp{rob}{l2}_pose": p, f"{tag}_pnp{rob}{l2}_outlier": oo})
        print(tag, "p3p ok", ok, "outliers", int(o.sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose.npz"), **d)


if __name__ == "__main__":
    main()
