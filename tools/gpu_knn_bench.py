"""Stand-alone timing of the Hamming 2-NN kernel on the bench's shape (64 x 1536 query slots, 1000 live, 10 000 map)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alvaar_b200
from alvaar_b200 import synth

dev = "cuda:0"
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
if len(sys.argv) > 2:
    assert ctx.L.alva_set_option(b"knn_qpw", int(sys.argv[2])) == 0
nb, qcap, live, nt = 64, 1536, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 10000
rng = np.random.default_rng(1)
q = rng.integers(0, 256, (nb * qcap, 32), dtype=np.uint8)
t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
q[::3] = t[rng.integers(0, nt, len(q[::3]))] ^ (rng.random((len(q[::3]), 32)) < 0.05).astype(np.uint8)   # planted near matches
dq, dt = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
counts = torch.full((nb,), live, dtype=torch.int32, device=dev)
out = torch.zeros((nb * qcap, 4), dtype=torch.int32, device=dev)
for _ in range(3):
    ctx.hamming_knn2_batch(dq, counts, nb, qcap, dt, nt, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.hamming_knn2_batch(dq, counts, nb, qcap, dt, nt, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"knn2_batch qpw={sys.argv[2] if len(sys.argv) > 2 else 8} live={live}: {ms*1000:.1f} us  -> {nb*live*nt/ms/1e6:.1f} G dist/s")
