#!/usr/bin/env python3
"""Dump golden vectors from the REFERENCE ITSELF (vendored OpenCV 4.5.5 / Ceres 2.0 built from
/root/reference by oracle/build_ref.sh -> oracle/_ref/libalva_ref.so) into tests/golden/*.npz.

Run in the build container only (the reference tree is not on the GPU box); the .npz files are
committed.  OpenCV's CPU dispatch is pinned to the SSE baseline (ref_config(0, 1)): that is the
arithmetic of the shipped WASM simd128 build (no FMA); the AVX2/FMA variant of the one
fusion-sensitive stage (ORB's float blur) is dumped as well (`blur_fma`).
"""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alvaar_b200 import synth  # noqa: E402

R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libalva_ref.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def main():
    R.ref_config(0, 1)
    buf = C.create_string_buffer(1024)
    R.ref_info(buf, 1024)
    info = buf.value.decode()
    print(info)

    # ---- gray
    rgba = synth.random_rgba(64, 48, 1, seed=11)[0]
    g = np.empty((48, 64), np.uint8)
    R.ref_gray(P(rgba), 64, 48, P(g))
    np.savez_compressed(os.path.join(OUT, "gray.npz"), rgba=rgba, gray=g, info=info)

    # ---- pyramid (odd sizes on purpose) through buildOpticalFlowPyramid(win 9, maxLevel 3)
    w, h = 161, 91
    img = synth.crop(w, h)
    lv = [np.zeros((h, w), np.uint8)]
    ww, hh = w, h
    for _ in range(3):
        ww, hh = (ww + 1) // 2, (hh + 1) // 2
        lv.append(np.zeros((hh, ww), np.uint8))
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in lv])
    got = R.ref_build_pyramid(P(img), w, h, 9, 3, ptrs, None)
    assert got == 3 and (lv[0] == img).all()
    np.savez_compressed(os.path.join(OUT, "pyramid.npz"), img=img, l1=lv[1], l2=lv[2], l3=lv[3], levels=got)

    # ---- FAST
    w, h = 320, 240
    img = synth.crop(w, h, 700, 500)
    d = {"img": img}
    for thr in (20, 7):
        for nms in (1, 0):
            xs = np.zeros((w * h, 3), np.int32)
            n = R.ref_fast(P(img), w, h, thr, nms, P(xs), w * h)
            d[f"kp_t{thr}_n{nms}"] = xs[:n].copy()
    np.savez_compressed(os.path.join(OUT, "fast.npz"), **d)

    # ---- ORB: blur (both dispatches), describe at points (constant -1 deg and given angles), detect
    blur = np.empty((h, w), np.uint8)
    R.ref_orb_blur(P(img), w, h, P(blur))
    R.ref_config(1, 1)
    blur_fma = np.empty((h, w), np.uint8)
    R.ref_orb_blur(P(img), w, h, P(blur_fma))
    R.ref_config(0, 1)
    rng = np.random.default_rng(3)
    n = 600
    pts = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1).astype(np.float32)
    pts[:100] = np.floor(pts[:100]) + 0.5
    pts[100:200] = np.floor(pts[100:200])
    desc = np.zeros((n, 32), np.uint8)
    kept = np.zeros(n, np.uint8)
    R.ref_orb_compute(P(img), w, h, P(pts), None, n, P(desc), P(kept))
    ang = rng.uniform(0, 360, n).astype(np.float32)
    desc_a = np.zeros((n, 32), np.uint8)
    kept_a = np.zeros(n, np.uint8)
    R.ref_orb_compute(P(img), w, h, P(pts), P(ang), n, P(desc_a), P(kept_a))
    kp = np.zeros((2000, 5), np.float32)
    dd = np.zeros((2000, 32), np.uint8)
    nd = R.ref_orb_detect(P(img), w, h, 300, 20, P(kp), P(dd), 2000)
    k7 = np.zeros(7, np.float32)
    R.ref_gaussian_kernel(7, C.c_double(2.0), P(k7))
    np.savez_compressed(os.path.join(OUT, "orb.npz"), img=img, blur=blur, blur_fma=blur_fma, pts=pts, desc=desc,
                        kept=kept, angles=ang, desc_angles=desc_a, kept_angles=kept_a, det_kp=kp[:nd], det_desc=dd[:nd],
                        gauss7=k7)

    # ---- Hamming 2-NN with exact duplicates (tie rule)
    q, t = synth.make_descriptors(200, 500, seed=7)
    t[400:450] = t[0:50]
    q[150:160] = t[0:10]
    out = np.zeros((200, 4), np.int32)
    R.ref_knn2(P(q), 200, P(t), 500, P(out))
    np.savez_compressed(os.path.join(OUT, "knn.npz"), q=q, t=t, out=out)
    # ---- local BA: ceres::Solve (SPARSE_SCHUR, LM, <= 5 it, Huber) on AlvaAR's anchored-inverse-depth functor
    pb = synth.make_ba_problem(8, 300, 3, seed=7)
    poses, invd = pb["poses"].copy(), pb["invd"].copy()
    summary, costs = np.zeros(8), np.zeros(64)
    R.ref_ba_solve(P(pb["calib"]), P(poses), P(pb["pose_const"]), len(poses), P(invd), P(pb["anch_kf"]), P(pb["anch_uv"]),
                   len(invd), P(pb["obs_kf"]), P(pb["obs_lm"]), P(pb["obs_uv"]), len(pb["obs_kf"]),
                   C.c_double(pb["huber"]), 5, P(summary), P(costs))
    np.savez_compressed(os.path.join(OUT, "ba.npz"), poses_out=poses, invd_out=invd, summary=summary, costs=costs, **pb)
    golden_ba_local()
    golden_scharr()
    print("golden vectors written to", OUT)


def golden_ba_local():
    """Optimizer::localBA steps 2-4 with Ceres itself (ref_ba_local): solve, remove outliers, re-solve, flag."""
    pb = synth.make_ba_problem(12, 800, 5, seed=3)
    poses, invd = pb["poses"].copy(), pb["invd"].copy()
    summary, flags = np.zeros(10), np.zeros(len(pb["obs_kf"]), np.int32)
    R.ref_ba_local.restype = C.c_int
    nbad = R.ref_ba_local(P(pb["calib"]), P(poses), P(pb["pose_const"]), len(poses), P(invd), P(pb["anch_kf"]), P(pb["anch_uv"]),
                          len(invd), P(pb["obs_kf"]), P(pb["obs_lm"]), P(pb["obs_uv"]), len(pb["obs_kf"]),
                          C.c_double(pb["huber"]), C.c_double(5.9915), 5, P(flags), P(summary))
    assert nbad == (flags == 1).sum() and nbad > 0
    np.savez_compressed(os.path.join(OUT, "ba_local.npz"), poses_out=poses, invd_out=invd, summary=summary, flags=flags, **pb)


def golden_scharr():
    """Derivative pyramid of cv::buildOpticalFlowPyramid(win 9, maxLevel 3, withDerivatives = true)."""
    w, h = 161, 91
    img = synth.crop(w, h, 300, 200)
    ws, hs = [w], [h]
    for _ in range(3):
        ws.append((ws[-1] + 1) // 2)
        hs.append((hs[-1] + 1) // 2)
    lv = [np.zeros((hs[k], ws[k]), np.uint8) for k in range(4)]
    dv = [np.zeros((hs[k], ws[k], 2), np.int16) for k in range(4)]
    LP = (C.c_void_p * 4)(*[a.ctypes.data for a in lv])
    DP = (C.c_void_p * 4)(*[a.ctypes.data for a in dv])
    got = R.ref_build_pyramid(P(img), w, h, 9, 3, LP, DP)
    assert got == 3
    np.savez_compressed(os.path.join(OUT, "scharr.npz"), img=img, **{f"l{k}": lv[k] for k in range(4)},
                        **{f"d{k}": dv[k] for k in range(4)})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ba_local":
        golden_ba_local()
    elif len(sys.argv) > 1 and sys.argv[1] == "scharr":
        golden_scharr()
    else:
        main()
