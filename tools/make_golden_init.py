#!/usr/bin/env python3
"""Dump map-initialisation golden vectors from the REFERENCE ITSELF: AlvaAR's MultiViewGeometry::compute5ptEssentialMatrix /
triangulate (src/slam/src/multi_view_geometry.cpp compiled unmodified with the vendored OpenGV 1.0 into
oracle/_ref/libalva_ref.so; sampler seed pinned through doRandom = false).  tests/golden/init.npz is committed.

Per problem: the RANSAC-only model (optimize = false), the refined model (optimize = true), the outlier set, and the refined
model for the same input perturbed by +-1 ulp -- the reference's refinement (Eigen LM on forward differences of a 1 - cos cost,
ftol = xtol = 10 eps) is noise-limited, and the spread of its OWN output under a 1-ulp input change is the honest tolerance
of that stage (tests/test_oracle_init.py)."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
f32 = C.c_float
R.ref_essential_5pt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, f32, C.c_int, f32, f32, C.c_void_p, C.c_void_p]

CASES = {"a": dict(n=192, seed=1, noise_px=0.3, outlier_frac=0.1), "b": dict(n=576, seed=2, noise_px=0.5, outlier_frac=0.25),
         "c": dict(n=60, seed=5, noise_px=0.2, outlier_frac=0.05, baseline=0.2), "d": dict(n=1296, seed=7, w=1920, h=1080, noise_px=0.4, outlier_frac=0.15),
         "e": dict(n=12, seed=9, noise_px=0.1, outlier_frac=0.0)}


def run(b1, b2, K, opt):
    n = len(b1)
    Rt, o = np.zeros(12), np.zeros(n, np.uint8)
    ok = R.ref_essential_5pt(P(b1), P(b2), n, 100, 3.0, opt, K[0], K[1], P(Rt), P(o))
    return ok, Rt, o


def main():
    R.ref_config(0, 1)
    d = {}
    rng = np.random.default_rng(123)
    for tag, kw in CASES.items():
        pr = synth.make_twoview_problem(**kw)
        K32 = pr["K"].astype(np.float32)
        ok0, Rt0, o0 = run(pr["bv1"], pr["bv2"], K32, 0)
        ok1, Rt1, o1 = run(pr["bv1"], pr["bv2"], K32, 1)
        b1 = np.ascontiguousarray(pr["bv1"] * (1 + rng.choice([-1, 0, 1], pr["bv1"].shape) * 2.2e-16))
        b2 = np.ascontiguousarray(pr["bv2"] * (1 + rng.choice([-1, 0, 1], pr["bv2"].shape) * 2.2e-16))
        ok2, Rt2, o2 = run(b1, b2, K32, 1)
        assert ok0 == ok1 == ok2 and (o0 == o1).all() and (o0 == o2).all()
        d.update({f"{tag}_bv1": pr["bv1"], f"{tag}_bv2": pr["bv2"], f"{tag}_K": K32, f"{tag}_ok": ok0, f"{tag}_ransac_Rt": Rt0,
                  f"{tag}_refined_Rt": Rt1, f"{tag}_refined_Rt_ulp": Rt2, f"{tag}_outlier": o0})
        print(tag, "ok", ok0, "outliers", int(o0.sum()), "refined spread under 1 ulp: dR %.1e" % np.abs(Rt1.reshape(3, 4)[:, :3] - Rt2.reshape(3, 4)[:, :3]).max())
    # triangulation: MultiViewGeometry::triangulate on the noisy correspondences of case a under its true relative pose
    pr = synth.make_twoview_problem(**CASES["a"])
    q = synth._quat_from_R(pr["R12"])
    Tlr = np.concatenate([pr["t12"], q])
    out = np.zeros((len(pr["bv1"]), 3))
    R.ref_triangulate(P(Tlr), P(pr["bv1"]), P(pr["bv2"]), len(out), P(out))
    d.update({"tri_Tlr": Tlr, "tri_points": out})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "init.npz"), **d)


if __name__ == "__main__":
    main()
