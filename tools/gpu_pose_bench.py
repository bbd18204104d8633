"""Times alva_k_p3p_lmeds + alva_k_pnp on the GPU box: B problems (one per frame) of N 3-D <-> 2-D correspondences."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import alvaar_b200
from pose_util import make_pose_problem

ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
for B, N in ((64, 1000), (64, 2000), (1, 1000)):
    prs = [make_pose_problem(N, i, outlier_frac=0.1) for i in range(min(B, 8))]
    bv = torch.from_numpy(np.stack([prs[i % len(prs)]["bv"] for i in range(B)])).cuda()
    X = torch.from_numpy(np.stack([prs[i % len(prs)]["X"] for i in range(B)])).cuda()
    uv = torch.from_numpy(np.stack([prs[i % len(prs)]["uv"] for i in range(B)])).cuda()
    K = torch.from_numpy(np.tile(prs[0]["K"], (B, 1))).cuda()
    pose0 = torch.from_numpy(np.stack([prs[i % len(prs)]["pose0"] for i in range(B)])).cuda()
    T = torch.zeros((B, 12), dtype=torch.float64, device="cuda"); out = torch.zeros((B, N), dtype=torch.uint8, device="cuda")
    info = torch.zeros((B, 4), dtype=torch.float64, device="cuda"); summ = torch.zeros((B, 12), dtype=torch.float64, device="cuda")
    def timeit(fn):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts[1:]) * 1e3
    t1 = timeit(lambda: ctx.p3p_lmeds(B, N, bv, X, None, T, out, info, fx=float(prs[0]["K"][0]), fy=float(prs[0]["K"][1])))
    def pnp():
        p = pose0.clone(); ctx.pnp(B, N, K, uv, X, None, p, out, summ, 2.4477, 5.9915)
    t2 = timeit(pnp)
    print(f"B={B} N={N}: p3p_lmeds {t1:.1f} us  pnp {t2:.1f} us  (ok={int(info[:,0].sum())}/{B}, pnp ok={int(summ[:,10].sum())})")
