"""Shared helpers of the System tests: golden trace access and the CPU-oracle build of the host-side state machine."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from conftest import P, ROOT, golden
from alvaar_b200 import synth

CAP = 4096


def frames_and_golden():
    g = golden("system")
    w, h, nf = int(g["w"]), int(g["h"]), int(g["nframes"])
    frames, _ = synth.make_frames(nf, w, h, seed=int(g["seed"]), rgba=True)
    assert hashlib.sha256(frames.tobytes()).hexdigest() == str(g["sha256"]), "synthetic frame generator changed: re-dump the golden"
    return g, frames


def frame_slice(g, pre, k):
    a, b = int(g[pre + "start"][k]), int(g[pre + "start"][k + 1])
    return g[pre + "ids"][a:b], g[pre + "px"][a:b], g[pre + "is3d"][a:b], g[pre + "wpt"][a:b]


def cpu_system_lib():
    """alvaar_b200/csrc/system_core.h over the CPU oracle backend (tests/host/system_cpu_backend.cpp) -- test infrastructure."""
    so = os.path.join(ROOT, "tests", "_build", "libsystem_cpu.so")
    orc = os.path.join(ROOT, "oracle", "_build", "libalva_oracle.so")
    srcs = [os.path.join(ROOT, "tests", "host", "system_cpu_backend.cpp"), os.path.join(ROOT, "alvaar_b200", "csrc", "system_core.h"), orc]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, srcs[0], orc,
                               "-Wl,-rpath," + os.path.dirname(orc)])
    S = C.CDLL(so)
    S.cpu_system_create.restype = C.c_void_p
    S.cpu_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
    S.cpu_system_process.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    S.cpu_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int]
    S.cpu_system_info.argtypes = [C.c_void_p, C.c_void_p]
    S.cpu_system_set_essential_hook.argtypes = [C.c_void_p, C.c_void_p]
    S.cpu_system_destroy.argtypes = [C.c_void_p]
    return S


def quat_dist(a, b):
    return float(min(np.abs(a - b).max(), np.abs(a + b).max()))
