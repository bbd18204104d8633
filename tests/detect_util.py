"""Shared helpers of the detector tests (oracle plumbing)."""
import ctypes as C

import numpy as np

from conftest import P


def oracle_detect(oracle, img, cs, cur, roi, q0=0.001, cap=4096):
    h, w = img.shape
    oracle.orc_detect_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
    out = np.zeros((cap, 2), np.float32)
    oi = np.zeros((cap, 2), np.int32)
    q = C.c_double(q0)
    cur = np.ascontiguousarray(cur, np.float32)
    n = oracle.orc_detect_points(P(np.ascontiguousarray(img)), w, h, cs, P(cur), len(cur), P(np.ascontiguousarray(roi, np.int32)),
                                 C.byref(q), P(out), P(oi), cap)
    return out[:n].copy(), oi[:n].copy(), q.value


def random_cur(w, h, n, seed):
    rng = np.random.default_rng(seed)
    if n == 0:
        return np.zeros((0, 2), np.float32)
    return np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)
