"""Times alva_k_match_to_map on the GPU box: a 720p keyframe (576 keypoints, 30 keyframes) against a 5760-point local map."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alvaar_b200
from alvaar_b200 import synth

ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
p = synth.make_match_problem(5, w=1280, h=720, n_kf=30, n_frame_kp=576, n_local=5760, dup_frac=0.3)
idx = {int(i): k for k, i in enumerate(p["mp_id"])}
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
kp_mp = d(np.array([idx[int(i)] for i in p["kp_id"]], np.int32)); local = d(np.array([idx[int(i)] for i in p["local_ids"]], np.int32))
args = [d(p[k]) for k in ("cur_T",)] + [kp_mp, d(p["kp_px"])]
T, px, kfT, wpt, is3d, os_, okf, opx, ds, desc = (d(p[k]) for k in ("cur_T", "kp_px", "kf_T", "mp_wpt", "mp_is3d", "obs_start", "obs_kf", "obs_px", "desc_start", "desc"))
out = torch.zeros(576, dtype=torch.int32, device="cuda"); dist = torch.zeros(576, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
ts = []
for _ in range(8):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.match_to_map(1280, 720, 40, [float(v) for v in p["K"]], T, kp_mp, px, 200, kfT, wpt, is3d, os_, okf, opx, ds, desc, local, out, dist, cnt); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"match_to_map 576 kps x 5760 local map points x 30 KFs: {min(ts[1:])*1e3:.1f} us, {int(cnt.item())} matches")
