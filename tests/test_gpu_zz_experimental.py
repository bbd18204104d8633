"""GPU test (B200), deliberately the LAST gpu test file in collection order: the experimental front-end instantiation
(alva_set_option("frontend_antipodal", 1)) against the default kernel.  Kept apart so that nothing else shares a process state
with an experimental kernel before it has been seen on a GPU."""
import numpy as np
import pytest
import torch

from alvaar_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_frontend_antipodal_variant_matches_default(gpu_ctx):
    """The experimental pre-test variant (alva_set_option("frontend_antipodal", 1); fast_swar.h -- 8 of the 16 ring flag words
    assembled from their antipodal partners) must return exactly the default kernel's corners (which the tests above pin to the
    reference).  Its logic is checked on the CPU by host emulation (tests/test_fast_swar_host.py); it has not been validated on a
    GPU yet, so a mismatch is reported as XFAIL instead of failing the suite -- the default path does not depend on it."""
    n, w, h = 4, 1280, 720
    frames, _ = synth.make_frames(n, w, h)
    d_in = dev(frames)
    cap = 32768
    res = []
    try:
        for opt, thr in ((0, 20), (1, 20), (0, 150), (1, 150), (0, 7), (1, 7)):
            gpu_ctx.L.alva_set_option(b"frontend_antipodal", opt)
            keys = torch.zeros((n, cap), dtype=torch.int32, device=DEV)
            counts = torch.zeros(n, dtype=torch.int32, device=DEV)
            l0 = torch.zeros((n, h, w), dtype=torch.uint8, device=DEV)
            l1 = torch.zeros((n, h // 2, w // 2), dtype=torch.uint8, device=DEV)
            gpu_ctx.frontend(d_in, w, h, n, l0, l1, None, None, thr, keys, counts, cap, True)
            torch.cuda.synchronize()
            res.append((keys.cpu().numpy(), counts.cpu().numpy(), l0.cpu().numpy(), l1.cpu().numpy()))
    finally:
        gpu_ctx.L.alva_set_option(b"frontend_antipodal", 0)
    for a, b in ((res[0], res[1]), (res[2], res[3]), (res[4], res[5])):
        same = (a[1] == b[1]).all() and all((a[0][f, :a[1][f]] == b[0][f, :b[1][f]]).all() for f in range(n)) and (a[2] == b[2]).all() and (a[3] == b[3]).all()
        if not same:
            pytest.xfail("experimental antipodal pre-test differs from the default kernel on this GPU")
    assert res[0][1].min() > 5000
