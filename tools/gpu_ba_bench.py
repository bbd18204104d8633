"""Micro-benchmark of alva_k_ba_solve (13 problems of 20 KF x 3000 landmarks x 9000 residual observations): gather-form
Schur (default) vs the FP64 tensor-core SYRK option.  CUDA-event timing, inputs reset on the device each repetition."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alvaar_b200
from alvaar_b200 import synth, lib
NP = 13
pb = synth.make_ba_problem(20, 3000, 4, seed=42)
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)  # handle 0 -> legacy default stream
st = lambda k: torch.from_numpy(np.stack([pb[k]] * NP)).cuda()
calib, poses0, const, invd0 = st("calib"), st("poses"), st("pose_const"), st("invd")
akf, auv, okf, olm, ouv = st("anch_kf"), st("anch_uv"), st("obs_kf"), st("obs_lm"), st("obs_uv")
summ = torch.zeros((NP, 8), dtype=torch.float64, device="cuda")
L = lib()
for name, opt in (("gather", 0), ("dense_dmma", 1), ("gather", 0)):
    L.alva_set_option(b"ba_dense_schur", opt)
    ts = []
    for rep in range(6):
        poses, invd = poses0.clone(), invd0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.ba_solve(NP, 20, 3000, len(pb["obs_kf"]), calib, poses, const, invd, akf, auv, okf, olm, ouv, pb["huber"], 5, summ)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    s = summ.cpu().numpy()[0]
    print(f"{name:12s} ms per solve batch (13 problems): {np.round(ts, 3)}  summary {s[:5]}")
L.alva_set_option(b"ba_dense_schur", 0)
