"""A/B timing of the fused front-end launch (64 x 1280x720 RGBA, one launch) over its kernel variants:
round-1 kernel, frontend_tile_kernel_v2, and v2 with the TMA L2 prefetch of a later frame's tile.  CUDA events, median of 8 after 2 warm-ups."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import alvaar_b200
from alvaar_b200 import synth

ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
n, w, h = 64, 1280, 720
fr, _ = synth.make_frames(4, w, h)
d = torch.from_numpy(np.ascontiguousarray(np.tile(fr, (16, 1, 1, 1)))).cuda()
l0 = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
l1 = torch.zeros((n, h // 2, w // 2), dtype=torch.uint8, device="cuda")
keys = torch.zeros((n, 32768), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
ALG = 4838400 * n
for name, var, anti, pf in (("round-1", 0, 0, 0), ("v2", 2, 0, 0), ("v2 + L2 prefetch", 2, 0, 1)) * 3:
    ctx.L.alva_set_option(b"frontend_variant", var)
    ctx.L.alva_set_option(b"frontend_antipodal", anti)
    ctx.L.alva_set_option(b"frontend_prefetch", pf)
    ts = []
    for _ in range(10):
        cnt.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.frontend(d, w, h, n, l0, l1, None, None, 20, keys, cnt, 32768, False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts[2:]))
    print(f"{name:22s} median {us:7.1f} us  -> {ALG / us / 1e3:7.1f} GB/s algorithmic  corners {int(cnt.sum())}")
ctx.L.alva_set_option(b"frontend_variant", 2)
ctx.L.alva_set_option(b"frontend_antipodal", 0)
ctx.L.alva_set_option(b"frontend_prefetch", 0)
