"""Frame-parallel sharding for multi-GPU runs (one process per GPU).

Frames of a stereo video are independent (no temporal state anywhere in
adcensus_stm, d_io.cu:7-238), so a stream is partitioned across ranks with no
data-path collective; torch.distributed is used only for the start/stop
barriers and for reducing the timing (max over ranks) and the frame counts.
"""
import os


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def frame_shard(num_frames, rank, world_size):
    """Contiguous block of frame indices owned by `rank`: sizes differ by at most one,
    every frame is owned exactly once, order inside a shard is stream order."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_frames, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def init_process_group(backend=None):
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1; returns True if initialised."""
    import torch
    import torch.distributed as dist
    rank, world, _ = dist_env()
    if world <= 1:
        return False
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return True


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_scalar(value, op="max", device=None):
    """All-reduce one float over the ranks (max or sum); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(local_units, local_seconds, device=None):
    """Whole-job throughput: units processed by all ranks / the slowest rank's time."""
    total = reduce_scalar(local_units, "sum", device)
    slowest = reduce_scalar(local_seconds, "max", device)
    return total / slowest if slowest > 0 else 0.0, total, slowest
