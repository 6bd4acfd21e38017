"""
Multi-GPU plumbing for the CaVE hot path: instances are independent, so a batch shards
by instance — one process per GPU, contiguous slices, NO data-path collective.  The only collectives
are the ones a data-parallel trainer needs anyway: the predictor's gradient all-reduce (DDP) and,
optionally, one scalar all-reduce to log the global loss.

The reference has no distributed code (its only parallelism is a pathos process pool over instances,
src/cave.py:259); these helpers are what replaces that pool on a multi-GPU box.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def instance_shard(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of a global batch for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_mean_loss(local_loss_sum: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Mean loss over the global batch from per-rank sums: one 2-element all-reduce."""
    t = torch.stack([local_loss_sum.detach().to(torch.float64).reshape(()),
                     torch.tensor(float(local_count), dtype=torch.float64, device=local_loss_sum.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0] / t[1].clamp(min=1.0)


def sharded_grad_scale(local_count: int, global_count: int, world: int) -> float:
    """DDP averages gradients over ranks.  With reduction='mean' on every rank's shard the averaged
    gradient equals the global-batch mean gradient when shards are equal; for ragged shards multiply the
    local loss by this factor first: (local_count / global_count) * world."""
    return float(local_count) * world / float(max(global_count, 1))


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (one process per GPU), so that the pinned
    host buffers it allocates afterwards are local to the GPU's PCIe root and eight ranks do not all stream through one
    socket's memory controllers.  Best effort: returns what it found / did; never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        info["pci"], info["numa_node"] = bus, node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"], info["cpus"] = True, len(allowed)
    except Exception as e:      # no sysfs / NVML in this container: leave the affinity alone
        info["error"] = f"{type(e).__name__}: {e}"
    return info
