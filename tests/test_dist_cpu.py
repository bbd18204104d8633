"""CPU test of the N > 1 host-side path with world_size 2 on the gloo backend (no GPU): stream sharding covers every
stream once, timing is the max over ranks, throughput is the sum of units over that time, and the keyframe-descriptor
all-gather returns the concatenation of the per-rank blocks."""
import os
import socket
import subprocess
import sys

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_gloo(tmp_path):
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), worker, str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1", GLOO_SOCKET_IFNAME="lo")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    res = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert sorted(list(res[0]["ids"]) + list(res[1]["ids"])) == list(range(8))        # every stream exactly once
    assert float(res[0]["tmax"]) == float(res[1]["tmax"]) == 15.0                     # max over ranks
    assert abs(float(res[0]["thr"]) - (64 * 8) / 15e-3) < 1e-6                        # all units / slowest rank
    for r_ in res:
        assert (r_["gd"][0] == res[0]["desc"]).all() and (r_["gd"][1] == res[1]["desc"]).all()   # gather == concat
        assert (r_["gc"] == np.array([[16, 3], [15, 4]])).all()


def test_keyframe_block_wire_format_constants():
    """the Python side of the loop-closure exchange and the C header agree on the keyframe block (no GPU: symbol + constants)"""
    import ctypes as C
    import re
    from alvaar_b200 import lib
    from alvaar_b200.loopclosure import HEADER_BYTES, MAGIC, VERSION, LcConfig, LcEvent, block_bytes
    L = lib()
    L.alva_lc_block_bytes.restype = C.c_size_t
    for n in (8, 1024, 1536):
        assert int(L.alva_lc_block_bytes(n)) == block_bytes(n) == 64 + 40 * n
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "alva_b200.h")).read()
    assert int(re.search(r"#define ALVA_LC_MAGIC\s+(0x[0-9A-Fa-f]+)", hdr).group(1), 16) == MAGIC == int.from_bytes(b"ALKF", "little")
    assert int(re.search(r"#define ALVA_LC_VERSION\s+(\d+)", hdr).group(1)) == VERSION
    assert int(re.search(r"#define ALVA_LC_HEADER_BYTES\s+(\d+)", hdr).group(1)) == HEADER_BYTES
    assert C.sizeof(LcConfig) == 13 * 4 and C.sizeof(LcEvent) == 6 * 4 + 12 * 8
