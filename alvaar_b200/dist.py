"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).  The hot path shards by camera stream:
rank r owns streams r, r + world, ... and there is no data-path collective; the only exchanges are the timing
reduction of bench.py and the optional keyframe-descriptor all-gather for cross-stream loop closure (the
reference has no multi-stream mode, so this step has no reference behaviour to match)."""
import torch
import torch.distributed as dist


def stream_ids(n_streams, rank, world):
    """Streams owned by `rank` (round-robin so that any world size covers every stream exactly once)."""
    return list(range(rank, n_streams, world))


def reduce_max(value, device="cpu"):
    """max over ranks of a python float (timing: a step ends when the slowest rank ends)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, ms_per_rank, device="cpu"):
    """whole-job throughput = units all ranks processed / max-over-ranks time."""
    t = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item()) / (reduce_max(ms_per_rank, device) * 1e-3)


def gather_keyframe_descriptors(desc, counts):
    """All-gather of each rank's fixed-size keyframe descriptor block: desc [K, cap, 32] uint8, counts [K] int32 ->
    ([world, K, cap, 32], [world, K]).  Fixed shapes so the exchange can live in a CUDA graph."""
    world = dist.get_world_size()
    out_d = torch.empty((world * desc.shape[0],) + tuple(desc.shape[1:]), dtype=desc.dtype, device=desc.device)
    out_c = torch.empty((world * counts.shape[0],) + tuple(counts.shape[1:]), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_d, desc.contiguous())       # concatenation along dim 0 (gloo and nccl)
    dist.all_gather_into_tensor(out_c, counts.contiguous())
    return out_d.view((world,) + tuple(desc.shape)), out_c.view((world,) + tuple(counts.shape))
