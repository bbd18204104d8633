"""Times alva_k_detect_grid on the GPU box: B keyframes 1280x720, cell 40, half the cells occupied."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alvaar_b200
from alvaar_b200 import synth

w, h, cs = 1280, 720, 40
fr, _ = synth.make_frames(4, w, h, seed=3, rgba=False)
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
for B in [int(x) for x in os.environ.get("DET_B", "13,64,1").split(",")]:
    imgs = torch.from_numpy(np.ascontiguousarray(fr)).cuda().repeat((B + 3) // 4, 1, 1)[:B].contiguous()
    rng = np.random.default_rng(0)
    ncur = 250
    cur = torch.from_numpy(np.stack([rng.uniform(0, w - 1, (B, ncur)), rng.uniform(0, h - 1, (B, ncur))], -1).astype(np.float32)).cuda()
    nc = torch.full((B,), ncur, dtype=torch.int32, device="cuda")
    out = torch.zeros((B, 1024, 2), dtype=torch.float32, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    ts = []
    for it in range(6):
        q = torch.full((B,), 0.001, dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.detect_grid(imgs, w, h, B, cs, cur, nc, ncur, [20, 20, w - 40, h - 40], q, out, None, cnt, 1024); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"detect_grid B={B}: {min(ts[1:])*1e3:.1f} us  ({int(cnt.float().mean())} corners/frame)")
