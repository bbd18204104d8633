"""Timing of the loop-closure detector on the bench's shape (13 keyframe blocks per step, n_max 1536, ~1000 live descriptors, world 2 on
one GPU: the remote rank shows the same scene two frames later): pack and detect, CUDA events; run under ncu for the per-kernel list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import alvaar_b200
from alvaar_b200 import synth
from alvaar_b200.loopclosure import LoopClosure, block_bytes

DEV = "cuda:0"
W, H, NMAX, K, NF = 1280, 720, 1536, 13, 15
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
frames, _ = synth.make_frames(NF, W, H, seed=7, rgba=True)
rgba = torch.from_numpy(np.ascontiguousarray(frames)).to(DEV)
gray = torch.zeros((NF, H, W), dtype=torch.uint8, device=DEV)
ctx.gray(rgba, gray, W, H, NF)
kp = torch.zeros((NF, NMAX, 4), dtype=torch.float32, device=DEV)
desc = torch.zeros((NF, NMAX, 32), dtype=torch.uint8, device=DEV)
cnt = torch.zeros(NF, dtype=torch.int32, device=DEV)
ctx.orb_detect(gray, W, H, NF, 1000, 20, 2, kp, desc, cnt, NMAX)
torch.cuda.synchronize()
cnt = torch.clamp(cnt, max=NMAX)
pts = kp[:, :, :2].contiguous()
print("live descriptors per frame:", cnt.cpu().numpy()[:5], "...")
K4 = synth.intrinsics(W, H)
dets = [LoopClosure(ctx, NMAX, K, 2, r, K4) for r in range(2)]
bb = block_bytes(NMAX)
gathered = torch.zeros(2 * K * bb, dtype=torch.uint8, device=DEV)
idx0 = torch.arange(0, K, dtype=torch.int32, device=DEV)
idx1 = idx0 + 2
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for rep in range(4):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    dets[0].pack(desc, pts, cnt, idx0, gathered[:K * bb])
    dets[1].pack(desc, pts, cnt, idx1, gathered[K * bb:])
    e1.record()
    dets[0].detect(gathered)
    e2.record()
    torch.cuda.synchronize()
    events = dets[0].poll(wait=True)
    s = dets[0].last_scores()
    print(f"rep {rep}: pack x2 {e0.elapsed_time(e1) * 1e3:7.1f} us   detect {e1.elapsed_time(e2) * 1e3:8.1f} us   events {len(events)}  "
          f"matches {s[:3, 1, 0]}  ok {s[:3, 1, 1]}  inliers {s[:3, 1, 2]}")
