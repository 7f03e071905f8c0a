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


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of sysfs cpulist files)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_cpus(device_index, sysfs="/sys/bus/pci/devices"):
    """One process per GPU: run this process on the CPUs next to its GPU (the PCI device's local_cpulist), so that the
    pinned frame buffers it allocates afterwards (first touch) and the copies into them stay on that socket.  Returns
    the CPU list, or None when the topology cannot be read or the binding is refused (nothing changes then)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(os.path.join(sysfs, addr, "local_cpulist")) as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001  (no sysfs entry, no permission, no such attribute: keep the default placement)
        return None
