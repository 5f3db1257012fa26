"""Multi-GPU: one process per GPU, contiguous clip/frame ranges, NO collective on the math path; one gather of the
features at the end (replaces the reference's single-process nn.DataParallel, src/preprocess_resnet_features.py
:214-217, which re-broadcasts all parameters on every forward)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa(local_rank: int) -> Optional[List[int]]:
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE it allocates pinned host buffers
    (first-touch then places them on the GPU's NUMA node).  With 8 ranks feeding 8 GPUs from host memory, buffers on
    the wrong socket halve the H2D rate.  Best effort: returns the CPU list, or None when NVML / affinity is not
    available (containers with a restricted cpuset keep their current mask)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_rank
        if vis:
            try:
                phys = int(vis.split(",")[local_rank])
            except (ValueError, IndexError):
                phys = local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: ceil(n / world) items per rank, the last ranks may get fewer/none."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_counts(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank row blocks [n_r, ...] (contiguous ranges from shard_range) to rank `dst` in rank order.

    Uses one padded `gather` over NCCL (GPU tensors) or gloo (CPU tensors).  Returns the [n_total, ...] tensor on
    `dst`, None elsewhere.  With world size 1 it is the identity."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    counts = all_counts(n_total, world)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
