"""Helper for tests/test_dist_cpu.py: one gloo rank.  Launched by torch.distributed.run."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alvaar_b200 import dist as ad  # noqa: E402


def main(outdir):
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = ad.stream_ids(8, rank, world)
    tmax = ad.reduce_max(10.0 + 5.0 * rank)
    thr = ad.aggregate_throughput(64 * len(ids), 10.0 + 5.0 * rank)
    rng = np.random.default_rng(rank)
    desc = torch.from_numpy(rng.integers(0, 256, (2, 16, 32), dtype=np.uint8))
    counts = torch.tensor([16 - rank, 3 + rank], dtype=torch.int32)
    gd, gc = ad.gather_keyframe_descriptors(desc, counts)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), ids=np.array(ids), tmax=tmax, thr=thr, gd=gd.numpy(), gc=gc.numpy(),
             desc=desc.numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
