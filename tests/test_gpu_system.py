"""GPU test (B200): the System facade (alva_system_*) -- the reference's public API surface (system.hpp:28-38)."""
import ctypes as C

import numpy as np
import pytest

from conftest import P
from alvaar_b200 import synth, lib

pytestmark = pytest.mark.gpu


def test_system_api(oracle):
    L = lib()
    L.alva_system_create.restype = C.c_void_p
    for f in ("alva_system_destroy", "alva_system_reset"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.alva_system_configure.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 8
    L.alva_system_find_camera_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.alva_system_find_camera_pose_imu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.alva_system_get_frame_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.alva_system_find_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.alva_system_num_matched.argtypes = [C.c_void_p]
    w, h = 640, 480
    fx, fy, cx, cy = synth.intrinsics(w, h)
    s = C.c_void_p(L.alva_system_create(0))
    pose = np.zeros(16, np.float32)
    frames, _ = synth.make_frames(3, w, h, seed=3)
    assert L.alva_system_find_camera_pose(s, P(frames[0]), P(pose)) == -4          # not configured -> ALVA_E_STATE
    assert L.alva_system_configure(s, w, h, fx, fy, cx, cy, 0, 0, 0, 0) == 0
    nkp = []
    for f in range(3):
        st = L.alva_system_find_camera_pose(s, P(frames[f]), P(pose))
        assert st == 3 and (pose == np.eye(4, dtype=np.float32).ravel()).all()      # honest: not initialised, identity
        xy = np.zeros((4096, 2), np.int32)
        n = L.alva_system_get_frame_points(s, P(xy), 4096)
        nkp.append(n)
        # the frame's points are exactly FAST + retainBest(max keypoints of the 40-px grid) of the oracle
        gray = np.empty((h, w), np.uint8)
        oracle.orc_gray(P(frames[f]), w, h, P(gray))
        xs = np.zeros((w * h // 4, 3), np.int32)
        m = oracle.orc_fast9(P(gray), w, h, 20, 1, P(xs), len(xs))
        k = xs[:m]
        k = np.ascontiguousarray(k[(k[:, 0] >= 31) & (k[:, 0] < w - 31) & (k[:, 1] >= 31) & (k[:, 1] < h - 31)])
        thr = oracle.orc_retain_best_threshold(P(k), len(k), 16 * 12)
        want = k[k[:, 2] >= thr][:, :2]
        assert n == len(want) and (xy[:n] == want).all()
        if f > 0:
            assert L.alva_system_num_matched(s) > 20          # consecutive frames overlap: many features re-found
    out = np.zeros(16, np.float32)
    assert L.alva_system_find_plane(s, P(out), 50) == 0
    imu = np.array([1.0, 0, 0, 0, 0], np.float64)
    assert L.alva_system_find_camera_pose_imu(s, P(frames[0]), P(imu), P(pose)) == 1
    assert np.allclose(pose, np.eye(4, dtype=np.float32).ravel())
    assert L.alva_system_reset(s) == 0
    L.alva_system_destroy(s)
