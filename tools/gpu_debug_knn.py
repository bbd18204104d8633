import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alvaar_b200
from alvaar_b200 import synth
O = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "libalva_oracle.so"))
P = lambda a: a.ctypes.data_as(C.c_void_p)
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
for nq, nt in [(1000, 10000), (1000, 1000), (1000, 1024), (1000, 999), (64, 1000), (1000, 2000), (7, 1), (1, 2), (33, 1025)]:
    q, t = synth.make_descriptors(nq, nt, seed=nq + nt, planted=0.3 if nt >= nq else 0.0)
    if nt > 100:
        t[nt // 2:nt // 2 + 20] = t[:20]
    want = np.zeros((nq, 4), np.int32)
    O.orc_knn2(P(q), nq, P(t), nt, P(want))
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    out = torch.full((nq, 4), -7, dtype=torch.int32, device="cuda")
    ctx.hamming_knn2(dq, nq, dt, nt, out)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    bad = np.nonzero((o != want).any(1))[0]
    print(nq, nt, "mismatch rows:", len(bad), "first:", bad[:5], "gpu", o[bad[:3]].tolist(), "want", want[bad[:3]].tolist())
