"""Times alva_k_klt_fb on the GPU box: B frame pairs of 1280x720, 1000 keypoints each, priors 2 px off (CUDA events)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alvaar_b200
from alvaar_b200 import synth

B, N, w, h = int(os.environ.get("KLT_B", 32)), 1000, 1280, 720
fr, _ = synth.make_frames(2, w, h, seed=3, rgba=True)
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
def pyramid(rgba):
    d = torch.from_numpy(rgba).cuda()[None].repeat(B, 1, 1, 1).contiguous()
    ws, hs = [w], [h]
    for _ in range(3): ws.append((ws[-1] + 1) // 2); hs.append((hs[-1] + 1) // 2)
    lv = [torch.zeros((B, hs[k], ws[k]), dtype=torch.uint8, device="cuda") for k in range(4)]
    keys = torch.zeros((B, 65536), dtype=torch.int32, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    ctx.frontend(d, w, h, B, lv[0], lv[1], lv[2], lv[3], 20, keys, cnt, 65536, False)
    de = [torch.zeros((B, hs[k], ws[k], 2), dtype=torch.int16, device="cuda") for k in range(4)]
    for k in range(4): ctx.scharr(lv[k], de[k], ws[k], hs[k], B)
    return lv, de
pa, da = pyramid(fr[0]); pb, db = pyramid(fr[1])
rng = np.random.default_rng(0)
pts = np.stack([rng.uniform(20, w - 20, (B, N)), rng.uniform(20, h - 20, (B, N))], -1).astype(np.float32)
pri0 = torch.from_numpy((pts + rng.normal(0, 2, pts.shape)).astype(np.float32)).cuda()
d_pts = torch.from_numpy(pts).cuda(); good = torch.zeros((B, N), dtype=torch.uint8, device="cuda")
for levels in (1, 3):
    ts = []
    for it in range(6):
        pri = pri0.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.klt_fb(pa, da, pb, db, w, h, B, levels, d_pts, pri, N, good); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"klt_fb levels={levels}: {B}x{N} points  {min(ts[1:])*1e3:.1f} us  ({B*N/min(ts[1:])/1e3:.2f} Mpts/s)  good={int(good.sum())}")
