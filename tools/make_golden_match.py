#!/usr/bin/env python3
"""Dump matchToMap golden vectors from the REFERENCE ITSELF: AlvaAR's own Mapper::matchToMap (src/slam/src/mapper.cpp:354-587,
compiled unmodified) running on its own Frame / MapPoint / MapManager objects rebuilt from the flat synthetic map of
alvaar_b200.synth.make_match_problem (oracle/ref_system.cpp::ref_match_to_map).  Stored: the local map's iteration order the
reference used (its unordered_set order decides ties) and the resulting keypoint -> map point pairs.  The problems themselves
are regenerated from their seeds (a checksum guards the generator).  tests/golden/match.npz is committed."""
import ctypes as C
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from alvaar_b200 import synth  # noqa: E402
from match_util import reference_match  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))


def digest(p):
    h = hashlib.sha256()
    for k in sorted(p):
        if isinstance(p[k], np.ndarray):
            h.update(np.ascontiguousarray(p[k]).tobytes())
    return h.hexdigest()


def main():
    R.ref_config(0, 1)
    d = {}
    for seed in range(4):
        p = synth.make_match_problem(seed, n_frame_kp=150 + 20 * seed, n_local=350 + 50 * seed)
        d[f"s{seed}_sha"] = digest(p)
        for nk in (100, 10):
            order, m = reference_match(R, p, nk)
            d[f"s{seed}_{nk}_order"] = order
            d[f"s{seed}_{nk}_kp"] = np.array(sorted(m), np.int32)
            d[f"s{seed}_{nk}_mp"] = np.array([m[k] for k in sorted(m)], np.int32)
            print(seed, nk, "matches", len(m))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "match.npz"), **d)


if __name__ == "__main__":
    main()
