"""Shared helpers of the map-initialisation tests (5-point essential matrix + triangulation)."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import P, ROOT

f32 = C.c_float
TAGS = ["a", "b", "c", "d", "e"]


def ransac_threshold(K):
    """compute5ptEssentialMatrix's inlier threshold in the reference's float arithmetic (multi_view_geometry.cpp:274-278)."""
    focal = np.float32((np.float32(K[0]) + np.float32(K[1])) / 2.)
    return 2.0 * (1.0 - float(np.cos(np.arctan(np.float32(3.0) / focal, dtype=np.float32), dtype=np.float32)))


def orc_essential(oracle, bv1, bv2, K, opt, seed=12345, max_iter=100, err=3.0):
    n = len(bv1)
    oracle.orc_essential_5pt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, f32, C.c_int, f32, f32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    Rt, o, info = np.zeros(12), np.zeros(n, np.uint8), np.zeros(4)
    ok = oracle.orc_essential_5pt(P(np.ascontiguousarray(bv1)), P(np.ascontiguousarray(bv2)), n, max_iter, err, opt, K[0], K[1], seed, P(Rt), P(o), P(info))
    return ok, Rt, o, info


def ref_essential(ref, bv1, bv2, K, opt, max_iter=100, err=3.0):
    n = len(bv1)
    ref.ref_essential_5pt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, f32, C.c_int, f32, f32, C.c_void_p, C.c_void_p]
    Rt, o = np.zeros(12), np.zeros(n, np.uint8)
    ok = ref.ref_essential_5pt(P(np.ascontiguousarray(bv1)), P(np.ascontiguousarray(bv2)), n, max_iter, err, opt, K[0], K[1], P(Rt), P(o))
    return ok, Rt, o


def host_core():
    """alvaar_b200/csrc/init_core.h (the arithmetic the CUDA kernels run) compiled for the host -- test infrastructure."""
    so = os.path.join(ROOT, "tests", "_build", "libinit_core_host.so")
    srcs = [os.path.join(ROOT, "tests", "host", "init_core_host.cpp"), os.path.join(ROOT, "alvaar_b200", "csrc", "init_core.h"),
            os.path.join(ROOT, "alvaar_b200", "csrc", "lmdif_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.host_essential_5pt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def pose_error(Rt_a, Rt_b):
    """(max |dR|, max |d t/|t||): the translation's length is a gauge freedom of the refinement (the caller normalises it)."""
    A, B = np.asarray(Rt_a).reshape(3, 4), np.asarray(Rt_b).reshape(3, 4)
    ta, tb = A[:, 3] / np.linalg.norm(A[:, 3]), B[:, 3] / np.linalg.norm(B[:, 3])
    return float(np.abs(A[:, :3] - B[:, :3]).max()), float(np.abs(ta - tb).max())


def refine_cost(Rt, bv1, bv2, inl):
    """the cost optimize_nonlinear minimises: sum over the inliers of (e1 + e2)^2 (relative_pose/methods.cpp:1097-1149)."""
    A = np.asarray(Rt).reshape(3, 4)
    R, t = A[:, :3], A[:, 3]
    f1, f2 = bv1[inl], bv2[inl]
    f2u = f2 @ R.T
    b0, b1 = f1 @ t, f2u @ t
    a00 = (f1 * f1).sum(1); a10 = (f1 * f2u).sum(1); a11 = -(f2u * f2u).sum(1)
    det = a00 * a11 + a10 * a10
    l0 = (a11 * b0 + a10 * b1) / det
    l1 = (-a10 * b0 + a00 * b1) / det
    p = (l0[:, None] * f1 + t + l1[:, None] * f2u) / 2
    r2 = (p - t) @ R
    e1 = 1 - (f1 * p).sum(1) / np.linalg.norm(p, axis=1)
    e2 = 1 - (f2 * r2).sum(1) / np.linalg.norm(r2, axis=1)
    return float(((e1 + e2) ** 2).sum())
