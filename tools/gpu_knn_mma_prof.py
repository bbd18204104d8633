"""The bench's matcher problem (64 x 1536 query slots, ~1137 live, 10 000-descriptor map) a few times, for ncu:
  ncu --metrics gpu__time_duration.sum ... python tools/gpu_knn_mma_prof.py [knn_mma option] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import alvaar_b200

opt = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda:0"
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
assert ctx.L.alva_set_option(b"knn_mma", opt) == 0
assert ctx.L.alva_set_option(b"knn_mma_kind", kind) == 0
rng = np.random.default_rng(5)
nb, qcap, live, nt = 64, 1536, 1137, 10000
q = rng.integers(0, 256, (nb * qcap, 32), dtype=np.uint8)
t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
q[::3] = t[rng.integers(0, nt, len(q[::3]))] ^ (rng.random((len(q[::3]), 32)) < 0.05).astype(np.uint8)
dq, dt = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
counts = torch.from_numpy(rng.integers(live - 100, live + 100, nb).astype(np.int32)).to(dev)
out = torch.zeros((nb * qcap, 4), dtype=torch.int32, device=dev)
for _ in range(rep):
    ctx.hamming_knn2_batch(dq, counts, nb, qcap, dt, nt, out)
torch.cuda.synchronize()
print("done", int((out[:, 0] >= 0).sum()))
