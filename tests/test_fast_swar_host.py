"""CPU test: host emulation of the front end's FAST-9 candidate pre-test (tests/host/fast_swar_host.cpp): the antipodal-sharing
variant in alvaar_b200/csrc/fast_swar.h (single source for device and host; opt-in on the GPU) equals the baseline bit-sliced
formulation on every lane / pixel that can carry a valid candidate, and both equal the scalar definition of the pre-test."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import P, ROOT

GP = 144


def lib():
    so = os.path.join(ROOT, "tests", "_build", "libfast_swar_host.so")
    srcs = [os.path.join(ROOT, "tests", "host", "fast_swar_host.cpp"), os.path.join(ROOT, "alvaar_b200", "csrc", "fast_swar.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.fsw_warp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.fsw_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    return L


def tile(rng, kind):
    if kind == "noise":
        g = rng.integers(0, 256, (14, GP), dtype=np.uint8)
    elif kind == "smooth":   # low-pass noise: corner density like a real image
        g = rng.integers(0, 256, (14 + 6, GP + 6)).astype(np.float32)
        k = np.ones(5) / 5
        g = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, g)
        g = np.apply_along_axis(lambda c: np.convolve(c, k, "same"), 0, g)
        g = ((g - g.min()) / (g.max() - g.min()) * 255)[3:17, 3:GP + 3].astype(np.uint8)
    else:                    # blobs: saturated plateaus, long equal runs, extreme contrasts
        g = np.where(rng.random((14, GP)) < 0.5, 0, 255).astype(np.uint8)
        g[:, ::7] = 128
    return np.ascontiguousarray(g)


@pytest.mark.parametrize("kind", ["noise", "smooth", "blobs"])
@pytest.mark.parametrize("thr", [7, 20, 60, 127, 128, 200])
def test_antipodal_equals_baseline_and_scalar(kind, thr):
    L = lib()
    rng = np.random.default_rng(hash((kind, thr)) % 2**32)
    first_word = 1   # as the kernel: lane l owns the word at index cbw + l with cbw >= 1 (a word to the left exists)
    ncand = 0
    for _ in range(20):
        g = tile(rng, kind)
        anti = np.zeros(32, np.uint32); base = np.zeros(32, np.uint32)
        L.fsw_warp(P(g), first_word, thr, P(anti), P(base))
        for lane in range(32):
            for j in range(4):
                x = 4 * (first_word + lane) + j
                for i in range(8):
                    bit = 1 << (8 * j + i)
                    want = L.fsw_scalar(P(g), x, 3 + i, thr)
                    assert bool(base[lane] & bit) == bool(want), ("baseline", kind, thr, lane, j, i)
                    # lanes 0 and 31 stand for the tile's 4-pixel halo words: the kernel only reads byte 3 of lane 0 and byte 0 of
                    # lane 31 (the 1-px NMS border); their other pixels have no neighbour lane to assemble from
                    if (lane == 0 and j < 3) or (lane == 31 and j > 0):
                        continue
                    assert bool(anti[lane] & bit) == bool(want), ("antipodal", kind, thr, lane, j, i)
                    ncand += want
    assert ncand > 0 or thr >= 127
