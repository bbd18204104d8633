#!/usr/bin/env python3
"""Development aid: run the reference's own System (oracle/_ref) and the host-side state machine of alva's System over the CPU
oracle backend (tests/_build/libsystem_cpu.so, TEST ONLY) on the same synthetic sequence and print where they part."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def main(nf=40, w=640, h=480, seed=7):
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
    S = C.CDLL(os.path.join(ROOT, "tests", "_build", "libsystem_cpu.so"))
    R.ref_config(0, 1)
    R.ref_system_create.restype = C.c_void_p
    R.ref_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 8
    R.ref_system_find_camera_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    R.ref_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    R.ref_system_info8.argtypes = [C.c_void_p, C.c_void_p]
    S.cpu_system_create.restype = C.c_void_p
    S.cpu_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
    S.cpu_system_process.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    S.cpu_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int]
    S.cpu_system_info.argtypes = [C.c_void_p, C.c_void_p]
    K = synth.intrinsics(w, h)
    frames, _ = synth.make_frames(nf, w, h, seed=seed, rgba=True)
    r = R.ref_system_create(w, h, K[0], K[1], K[2], K[3], 0, 0, 0, 0)
    s = S.cpu_system_create(w, h, K[0], K[1], K[2], K[3])
    if os.environ.get("REF_INIT"):
        S.cpu_system_set_essential_hook.argtypes = [C.c_void_p, C.c_void_p]
        S.cpu_system_set_essential_hook(s, C.cast(R.ref_essential_5pt, C.c_void_p))
    cap = 4096
    for k in range(nf):
        f = np.ascontiguousarray(frames[k])
        pose = np.zeros(16, np.float32)
        st_r = R.ref_system_find_camera_pose(r, P(f), k * 33.333, P(pose))
        T_s = np.zeros(7)
        st_s = S.cpu_system_process(s, P(f), k * 33.333, P(T_s))
        ids_r = np.zeros(cap, np.int32); px_r = np.zeros((cap, 2), np.float32); d3_r = np.zeros(cap, np.uint8); w_r = np.zeros((cap, 3)); T_r = np.zeros(7)
        n_r = R.ref_system_keypoints(r, P(ids_r), P(px_r), P(d3_r), P(w_r), cap, P(T_r))
        ids_s = np.zeros(cap, np.int32); px_s = np.zeros((cap, 2), np.float32); d3_s = np.zeros(cap, np.uint8); w_s = np.zeros((cap, 3))
        n_s = S.cpu_system_keypoints(s, P(ids_s), P(px_s), P(d3_s), P(w_s), cap)
        i_r = np.zeros(8, np.int32); i_s = np.zeros(8, np.int32)
        R.ref_system_info8(r, P(i_r)); S.cpu_system_info(s, P(i_s))
        same_order = n_r == n_s and (ids_r[:n_r] == ids_s[:n_s]).all()
        same_set = set(ids_r[:n_r]) == set(ids_s[:n_s])
        dpx = dw = float("nan")
        if same_order and n_r:
            dpx = float(np.abs(px_r[:n_r] - px_s[:n_s]).max())
            dw = float(np.abs(w_r[:n_r] - w_s[:n_s]).max())
            same3 = (d3_r[:n_r] == d3_s[:n_s]).all()
        else:
            same3 = False
        dq = min(np.abs(T_r[3:] - T_s[3:]).max(), np.abs(T_r[3:] + T_s[3:]).max())
        print(f"f{k:02d} st {st_r}/{st_s} n {n_r}/{n_s} order {int(same_order)} set {int(same_set)} 3d {int(same3)} "
              f"dpx {dpx:.2e} dwpt {dw:.2e} dt {np.abs(T_r[:3] - T_s[:3]).max():.2e} dq {dq:.2e} info r{i_r.tolist()} s{i_s.tolist()}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
