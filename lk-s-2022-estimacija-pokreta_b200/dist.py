"""Multi-GPU plumbing.  The hot path shards by (pair, direction): image pairs and the two directions of a pair are
independent (README.md:40), so ranks never exchange data; `torch.distributed` is used only for the barrier and the
max-over-ranks timing the benchmark contract asks for (NCCL on GPUs, gloo in the CPU tests)."""
import os


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_units(n_units, rank, world_size):
    """Pair indices this rank processes: round robin, so every rank gets floor or ceil of n/world units.
    A pair's two directions share both DAISY descriptor sets and meet in the consistency check, so the unit is the
    pair."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, n_units, world_size))


def init(backend=None, device=None):
    import torch
    import torch.distributed as dist
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, ws, local


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a Python float over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
