"""ms per pipeline step for several configurations (what does each part of the step cost once everything overlaps?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alvaar_b200
from alvaar_b200 import synth
from alvaar_b200.pipeline import Pipeline

W, H, B = 1280, 720, 64
dev = "cuda:0"
stream = torch.cuda.Stream()
ctx = alvaar_b200.Context(0, stream.cuda_stream)
frames, _ = synth.make_frames(B, W, H, seed=99)
_, mapd = synth.make_descriptors(8, 10000, seed=7)
ba = synth.make_ba_problem(20, 3000, 4, seed=42)
d_in = torch.from_numpy(frames).to(dev)


def run(name, kf, deriv, harris, mapsize=10000, overlap=1, qpw=4):
    ctx.L.alva_set_option(b"pipeline_ba_overlap", overlap)
    ctx.L.alva_set_option(b"knn_qpw", qpw)
    flags = alvaar_b200.ORB_IC_ANGLE | (alvaar_b200.ORB_HARRIS if harris else 0)
    pipe = Pipeline(ctx, W, H, B, fast_thr=20, nfeatures=1000, orb_flags=flags, map_size=mapsize, kf_interval=kf, ba_nkf=20,
                    ba_nlm=3000, ba_nobs=len(ba["obs_kf"]), ba_max_iter=5, ba_huber=ba["huber"], derivatives=deriv)
    if mapsize:
        pipe.set_map(mapd[:mapsize])
    for s in range(pipe.nprob):
        pipe.set_ba(s, ba)
    with torch.cuda.stream(stream):
        for _ in range(3):
            pipe.step_dev(d_in)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            pipe.step_dev(d_in)
        e1.record(stream)
        torch.cuda.synchronize()
    print(f"{name:46s} {e0.elapsed_time(e1) / 20:7.3f} ms/step")
    pipe.close()


run("full (BA overlapped, derivs, Harris)", 5, True, True)
run("full, knn qpw 8", 5, True, True, qpw=8)
run("full, BA serial", 5, True, True, overlap=0)
run("no BA", 0, True, True)
run("no BA, no derivatives", 0, False, True)
run("no BA, no derivatives, FAST-score selection", 0, False, False)
run("no BA, no derivs, no Harris, no matching", 0, False, False, mapsize=0)
run("BA only-ish (no matching, no derivs)", 5, False, False, mapsize=0)
