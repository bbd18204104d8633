"""Shared helpers of the KLT tests: pyramids built with the CPU oracle (test infrastructure) and ctypes plumbing."""
import ctypes as C

import numpy as np

from conftest import P


def build_pyramid(oracle, gray, levels):
    """levels + 1 tightly packed u8 levels and their Scharr derivative images (oracle: orc_pyrdown / orc_scharr)."""
    imgs, ders = [np.ascontiguousarray(gray)], []
    for k in range(levels + 1):
        if k > 0:
            hh, ww = imgs[-1].shape
            nxt = np.empty(((hh + 1) // 2, (ww + 1) // 2), np.uint8)
            oracle.orc_pyrdown(P(imgs[-1]), ww, hh, P(nxt))
            imgs.append(nxt)
        hh, ww = imgs[k].shape
        d = np.empty((hh, ww, 2), np.int16)
        oracle.orc_scharr(P(imgs[k]), ww, hh, P(d))
        ders.append(d)
    return imgs, ders


def table(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def oracle_fb_klt(oracle, pa, da, pb, db, w, h, levels, pts, priors, err_val=30.0, fb=0.5, win=9):
    n = len(pts)
    q = np.ascontiguousarray(priors, np.float32).copy()
    good = np.zeros(n, np.uint8)
    oracle.orc_fb_klt.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                   C.c_int, C.c_int, C.c_double]
    oracle.orc_fb_klt(table(pa), table(da), table(pb), table(db), w, h, len(pa) - 1, levels, win, err_val, fb,
                      P(np.ascontiguousarray(pts, np.float32)), P(q), P(good), n, 30, 0.01)
    return q, good


def oracle_klt_lk(oracle, pa, da, pb, w, h, levels, pts, nxt, use_initial=1, max_count=30, eps=0.01, win=9):
    n = len(pts)
    q = np.ascontiguousarray(nxt, np.float32).copy()
    st = np.zeros(n, np.uint8)
    er = np.zeros(n, np.float32)
    oracle.orc_klt_lk.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_double, C.c_int,
                                                                                                     C.c_double]
    oracle.orc_klt_lk(table(pa), table(da), table(pb), w, h, min(levels, len(pa) - 1), P(np.ascontiguousarray(pts, np.float32)),
                      P(q), P(st), P(er), n, win, max_count, eps, use_initial, 1e-4)
    return q, st, er


def klt_points(w, h, n, seed, sigma=2.0):
    """Seeded previous positions (some outside the image on purpose) and noisy priors."""
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-5, w + 5, n), rng.uniform(-5, h + 5, n)], 1).astype(np.float32)
    pri = (pts + rng.normal(0, sigma, (n, 2))).astype(np.float32)
    return pts, pri
