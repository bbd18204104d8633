#!/usr/bin/env python3
"""Dump System-level golden vectors (tests/golden/system.npz, committed).

`ref_*`: the REFERENCE ITSELF -- AlvaAR's own `System` (every src/slam/src/*.cpp except the Emscripten binding, compiled
unmodified into oracle/_ref/libalva_ref.so; pins: sampler seed 12345, injected time stamps k * 33.333 ms, 1 thread --
oracle/ref_system.cpp).  Per frame: status, pose (float[16] as the API returns it, and [t, q] in double), frame / keyframe
counters, and every keypoint in the iteration order of Frame::mapKeypoints_ (track id, pixel position, 3-D flag, world point).
The first `n_desc` frames' keyframe descriptors are stored for frame 0.

`cpu_*`: alva's own host-side System state machine (alvaar_b200/csrc/system_core.h) run over the CPU oracle backend
(tests/host/system_cpu_backend.cpp) on the same frames -- the trajectory the GPU build must reproduce tightly (same arithmetic,
same initialisation); its distance to `ref_*` after the initialisation is bounded by the reference's own noise-limited 5-point
refinement (tests/test_oracle_init.py).

The input frames are alvaar_b200.synth.make_frames(seed); their SHA-256 is stored so that a change of the generator is noticed."""
import ctypes as C
import hashlib
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
CAP = 4096


def build_cpu_system():
    so = os.path.join(ROOT, "tests", "_build", "libsystem_cpu.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                           os.path.join(ROOT, "tests", "host", "system_cpu_backend.cpp"), os.path.join(ROOT, "oracle", "_build", "libalva_oracle.so"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")])
    S = C.CDLL(so)
    S.cpu_system_create.restype = C.c_void_p
    S.cpu_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
    S.cpu_system_process.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    S.cpu_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int]
    S.cpu_system_info.argtypes = [C.c_void_p, C.c_void_p]
    S.cpu_system_set_essential_hook.argtypes = [C.c_void_p, C.c_void_p]
    S.cpu_system_destroy.argtypes = [C.c_void_p]
    return S


class Trace:
    def __init__(self):
        self.status, self.T, self.info, self.start, self.ids, self.px, self.is3d, self.wpt = [], [], [], [0], [], [], [], []

    def add(self, st, T, info, ids, px, is3d, wpt):
        self.status.append(st); self.T.append(T.copy()); self.info.append(info.copy())
        self.ids.append(ids.copy()); self.px.append(px.copy()); self.is3d.append(is3d.copy()); self.wpt.append(wpt.copy())
        self.start.append(self.start[-1] + len(ids))

    def dump(self, pre):
        return {pre + "status": np.array(self.status, np.int32), pre + "Twc": np.array(self.T), pre + "info": np.array(self.info, np.int32),
                pre + "start": np.array(self.start, np.int32), pre + "ids": np.concatenate(self.ids), pre + "px": np.concatenate(self.px),
                pre + "is3d": np.concatenate(self.is3d), pre + "wpt": np.concatenate(self.wpt)}


def run_cpu(S, frames, K, hook=None):
    w, h = frames.shape[2], frames.shape[1]
    s = S.cpu_system_create(w, h, K[0], K[1], K[2], K[3])
    if hook is not None:
        S.cpu_system_set_essential_hook(s, hook)
    tr = Trace()
    for k in range(len(frames)):
        T = np.zeros(7)
        st = S.cpu_system_process(s, P(np.ascontiguousarray(frames[k])), k * 33.333, P(T))
        ids = np.zeros(CAP, np.int32); px = np.zeros((CAP, 2), np.float32); d3 = np.zeros(CAP, np.uint8); wp = np.zeros((CAP, 3)); info = np.zeros(8, np.int32)
        n = S.cpu_system_keypoints(s, P(ids), P(px), P(d3), P(wp), CAP)
        S.cpu_system_info(s, P(info))
        tr.add(st, T, info, ids[:n], px[:n], d3[:n], wp[:n])
    S.cpu_system_destroy(s)
    return tr


def trace_only(libpath, out):
    """the reference System trace of another build of the reference (status, Twc, keypoint ids per frame) -> out (npz)"""
    R = C.CDLL(libpath)
    R.ref_system_create.restype = C.c_void_p
    R.ref_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 8
    R.ref_system_find_camera_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    R.ref_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    R.ref_system_destroy.argtypes = [C.c_void_p]
    R.ref_config(0, 1)
    R.ref_config_time_caps(1)
    w, h, nf, seed = 640, 480, 100, 7
    K = synth.intrinsics(w, h)
    frames, _ = synth.make_frames(nf, w, h, seed=seed, rgba=True)
    s = R.ref_system_create(w, h, K[0], K[1], K[2], K[3], 0, 0, 0, 0)
    status, Ts, ids_all, start = [], [], [], [0]
    for k in range(nf):
        pose = np.zeros(16, np.float32)
        st = R.ref_system_find_camera_pose(s, P(np.ascontiguousarray(frames[k])), k * 33.333, P(pose))
        ids = np.zeros(CAP, np.int32); px = np.zeros((CAP, 2), np.float32); d3 = np.zeros(CAP, np.uint8); wp = np.zeros((CAP, 3)); T = np.zeros(7)
        n = R.ref_system_keypoints(s, P(ids), P(px), P(d3), P(wp), CAP, P(T))
        status.append(st); Ts.append(T.copy()); ids_all.append(ids[:n].copy()); start.append(start[-1] + n)
    R.ref_system_destroy(s)
    np.savez(out, status=np.array(status, np.int32), Twc=np.array(Ts), ids=np.concatenate(ids_all), start=np.array(start, np.int32))


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--trace":
        return trace_only(sys.argv[2], sys.argv[3])
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
    R.ref_system_create.restype = C.c_void_p
    R.ref_system_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 8
    R.ref_system_find_camera_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    R.ref_system_get_frame_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    R.ref_system_destroy.argtypes = [C.c_void_p]
    R.ref_system_get_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    R.ref_system_keypoints.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    R.ref_system_info8.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_config(0, 1)
    R.ref_config_time_caps(1)   # the three wall-clock caps of the Ceres solves lifted (oracle/build_ref.sh): the golden must not depend on host load
    w, h, nf, seed = 640, 480, 100, 7
    K = synth.intrinsics(w, h)
    frames, _ = synth.make_frames(nf, w, h, seed=seed, rgba=True)
    d = {"w": w, "h": h, "nframes": nf, "seed": seed, "K": np.array(K), "sha256": hashlib.sha256(frames.tobytes()).hexdigest()}
    s = R.ref_system_create(w, h, K[0], K[1], K[2], K[3], 0, 0, 0, 0)
    tr = Trace()
    pose16, xys = [], []
    for k in range(nf):
        pose = np.zeros(16, np.float32)
        st = R.ref_system_find_camera_pose(s, P(np.ascontiguousarray(frames[k])), k * 33.333, P(pose))
        ids = np.zeros(CAP, np.int32); px = np.zeros((CAP, 2), np.float32); d3 = np.zeros(CAP, np.uint8); wp = np.zeros((CAP, 3)); T = np.zeros(7); info = np.zeros(8, np.int32)
        n = R.ref_system_keypoints(s, P(ids), P(px), P(d3), P(wp), CAP, P(T))
        R.ref_system_info8(s, P(info))
        tr.add(st, T, info, ids[:n], px[:n], d3[:n], wp[:n])
        pose16.append(pose)
        xy = np.zeros((CAP, 2), np.int32); i2 = np.zeros(CAP, np.int32); p2 = np.zeros((CAP, 2), np.float32)
        m = R.ref_system_get_frame_points(s, P(xy), P(i2), P(p2), CAP)   # System::getFramePoints: the 2-D keypoints, truncated unpx
        xys.append(xy[:m].copy())
        if k == 0:   # descriptors are computed at keyframe creation and carried by the tracked keypoints
            desc = np.zeros((CAP, 32), np.uint8); has = np.zeros(CAP, np.uint8)
            R.ref_system_get_descriptors(s, P(desc), P(has), CAP)
            d["f0_desc"], d["f0_has_desc"] = desc[:m], has[:m]
        print(k, "status", st, "keypoints", n, "3-D", int(d3[:n].sum()), "keyframe", info[1])
    R.ref_system_destroy(s)
    d.update(tr.dump("ref_"))
    # The reference's OWN sensitivity on this trace: the same run with one intrinsic moved by one or two ulps (16 runs; every bearing vector
    # then changes in its last bit).  Its 5-point refinement is noise-limited (tests/test_oracle_init.py), so the trajectory
    # the reference itself produces moves by `ref_spread_*` -- no independent implementation can be asked to sit closer to
    # `ref_Twc` than the reference sits to itself.  Stored per frame: max over the perturbed runs of |dt|_inf and of the
    # quaternion distance; runs whose discrete state (status, keypoint ids) leaves the base run's are dropped from that frame on.
    base_T = np.array(tr.T)
    spread_t, spread_q, nruns = np.zeros(nf), np.zeros(nf), np.zeros(nf, np.int32)
    for which in range(4):
        for sgn in (+1, -1, +2, -2):
            Kp = list(K)
            for _ in range(abs(sgn)):
                Kp[which] = float(np.nextafter(Kp[which], Kp[which] + sgn))
            s2 = R.ref_system_create(w, h, Kp[0], Kp[1], Kp[2], Kp[3], 0, 0, 0, 0)
            alive = True
            for k in range(nf):
                pose = np.zeros(16, np.float32)
                st = R.ref_system_find_camera_pose(s2, P(np.ascontiguousarray(frames[k])), k * 33.333, P(pose))
                ids = np.zeros(CAP, np.int32); px = np.zeros((CAP, 2), np.float32); d3 = np.zeros(CAP, np.uint8); wp = np.zeros((CAP, 3)); T = np.zeros(7)
                n = R.ref_system_keypoints(s2, P(ids), P(px), P(d3), P(wp), CAP, P(T))
                alive = alive and st == tr.status[k] and n == len(tr.ids[k]) and (ids[:n] == tr.ids[k]).all()
                if not alive:
                    break
                spread_t[k] = max(spread_t[k], float(np.abs(T[:3] - base_T[k, :3]).max()))
                qd = 1.0 - abs(float(np.dot(T[3:], base_T[k, 3:])))
                spread_q[k] = max(spread_q[k], float(np.sqrt(max(2.0 * qd, 0.0))))
                nruns[k] += 1
            R.ref_system_destroy(s2)
    # ... and under a rebuild: the SAME reference sources compiled with FMA contraction (oracle/build_ref.sh, ALVA_REF_VARIANT=fma:
    # -O2 -mfma -ffp-contract=fast for AlvaAR's sources and OpenGV -- gcc's default on aarch64, -march=native on x86).  Two
    # legitimate builds of the reference; how far apart they end up is the floor for any third implementation.
    alt = os.path.join(ROOT, "oracle", "_ref", "libalva_ref_fma.so")
    build_t, build_q, build_alive = np.zeros(nf), np.zeros(nf), np.zeros(nf, np.uint8)
    if os.path.exists(alt):
        # in its own process: two builds of the same C++ symbols cannot share one
        tmp = os.path.join(ROOT, "tests", "_build", "alt_trace.npz")
        os.makedirs(os.path.dirname(tmp), exist_ok=True)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--trace", alt, tmp])
        t2 = np.load(tmp)
        alive = True
        for k in range(nf):
            a0, a1 = int(t2["start"][k]), int(t2["start"][k + 1])
            alive = alive and int(t2["status"][k]) == tr.status[k] and (a1 - a0) == len(tr.ids[k]) and (t2["ids"][a0:a1] == tr.ids[k]).all()
            if not alive:
                break
            T = t2["Twc"][k]
            build_t[k] = float(np.abs(T[:3] - base_T[k, :3]).max())
            build_q[k] = float(np.sqrt(max(2.0 * (1.0 - abs(float(np.dot(T[3:], base_T[k, 3:])))), 0.0)))
            build_alive[k] = 1
        print("reference (-O2) vs reference (-O2 -mfma -ffp-contract=fast): max |dt|", float(build_t.max()), "max |dq|", float(build_q.max()),
              "same discrete state for", int(build_alive.sum()), "of", nf, "frames")
    else:
        print("oracle/_ref/libalva_ref_fma.so not built (ALVA_REF_VARIANT=fma bash oracle/build_ref.sh): ref_build_* left at zero")
    d["ref_build_t"], d["ref_build_q"], d["ref_build_alive"] = build_t, build_q, build_alive
    d["ref_spread_t"], d["ref_spread_q"], d["ref_spread_runs"] = spread_t, spread_q, nruns
    print("reference vs itself under a 1-ulp change of one intrinsic: max |dt|", float(spread_t.max()), "max |dq|", float(spread_q.max()),
          "runs alive at the last frame:", int(nruns[-1]), "of 16")
    d["ref_pose16"] = np.array(pose16)
    d["ref_xy_start"] = np.cumsum([0] + [len(x) for x in xys]).astype(np.int32)
    d["ref_xy"] = np.concatenate(xys)
    kfid = d["ref_info"][:, 1]
    d["first_ba_frame"] = int(np.argmax(kfid >= 2)) if (kfid >= 2).any() else nf   # Optimizer::localBA runs from keyframe id 2 on
    S = build_cpu_system()
    own = run_cpu(S, frames, K)
    d.update(own.dump("cpu_"))
    # sanity of what is committed: lockstep with the reference given the reference's own initialisation stage; its result
    # (what MultiViewGeometry::compute5ptEssentialMatrix returned inside the reference run) is recorded for the GPU test
    rec = {}
    HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p)
    R.ref_essential_5pt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]

    def hook(b1, b2, n, it, err, opt, fx, fy, Rt, outl):
        ok = R.ref_essential_5pt(b1, b2, n, it, err, opt, fx, fy, Rt, outl)
        if ok and "Rt" not in rec:
            rec["Rt"] = np.ctypeslib.as_array(C.cast(Rt, C.POINTER(C.c_double)), (12,)).copy()
            rec["outlier"] = np.ctypeslib.as_array(C.cast(outl, C.POINTER(C.c_uint8)), (n,)).copy()
        return ok
    cb = HOOK(hook)
    hooked = run_cpu(S, frames, K, C.cast(cb, C.c_void_p))
    d["ref_init_Rt"], d["ref_init_outlier"] = rec["Rt"], rec["outlier"]
    for k in range(nf):
        assert hooked.status[k] == tr.status[k] and (hooked.ids[k] == tr.ids[k]).all() and (hooked.px[k].view(np.uint32) == tr.px[k].view(np.uint32)).all()
        assert np.abs(hooked.T[k] - tr.T[k]).max() < 1e-9
    print("lockstep with the reference (its own initialisation plugged in) over all", nf, "frames; own initialisation: max |dt|",
          float(np.abs(np.array(own.T)[:, :3] - np.array(tr.T)[:, :3]).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "system.npz"), **d)


if __name__ == "__main__":
    main()
