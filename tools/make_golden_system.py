#!/usr/bin/env python3
"""Dump System-level golden vectors from the REFERENCE ITSELF: AlvaAR's own `System` (every src/slam/src/*.cpp except the
Emscripten binding, compiled unmodified into oracle/_ref/libalva_ref.so; pins: sampler seed 12345, injected time stamps
k * 33.333 ms, 1 thread -- oracle/ref_system.cpp).  Per frame: status, pose, and the 2-D keypoints (track ids, pixel
positions, truncated getFramePoints coordinates).  The input frames are alvaar_b200.synth.make_frames(seed) -- their SHA-256
is stored so that a change of the generator is noticed.  tests/golden/system.npz is committed."""
import ctypes as C
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
R.ref_system_create.restype = C.c_void_p
R.ref_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 8
R.ref_system_find_camera_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
R.ref_system_get_frame_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
R.ref_system_destroy.argtypes = [C.c_void_p]
R.ref_system_get_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]


def main():
    R.ref_config(0, 1)
    w, h, nf, seed = 640, 480, 20, 7
    K = synth.intrinsics(w, h)
    frames, _ = synth.make_frames(nf, w, h, seed=seed, rgba=True)
    d = {"w": w, "h": h, "nframes": nf, "seed": seed, "K": np.array(K), "sha256": hashlib.sha256(frames.tobytes()).hexdigest()}
    s = R.ref_system_create(w, h, K[0], K[1], K[2], K[3], 0, 0, 0, 0)
    status = []
    for k in range(nf):
        pose = np.zeros(16, np.float32)
        st = R.ref_system_find_camera_pose(s, P(np.ascontiguousarray(frames[k])), k * 33.333, P(pose))
        xy = np.zeros((4096, 2), np.int32); ids = np.zeros(4096, np.int32); px = np.zeros((4096, 2), np.float32)
        n = R.ref_system_get_frame_points(s, P(xy), P(ids), P(px), 4096)
        o = np.argsort(ids[:n])
        d[f"f{k}_ids"], d[f"f{k}_px"], d[f"f{k}_xy"], d[f"f{k}_pose"] = ids[:n][o], px[:n][o], xy[:n][o], pose
        if k in (0, 5):   # descriptors are computed at keyframe creation and carried by the tracked keypoints
            desc = np.zeros((4096, 32), np.uint8); has = np.zeros(4096, np.uint8)
            R.ref_system_get_descriptors(s, P(desc), P(has), 4096)
            d[f"f{k}_desc"], d[f"f{k}_has_desc"] = desc[:n][o], has[:n][o]
        status.append(st)
        print(k, "status", st, "2-D keypoints", n)
    d["status"] = np.array(status, np.int32)
    R.ref_system_destroy(s)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "system.npz"), **d)


if __name__ == "__main__":
    main()
