"""Clip sharding for the multi-GPU front end: every clip is independent (its statistics are per
clip), so ranks get disjoint sets of clips and never communicate."""
from __future__ import annotations

import numpy as np


def shard_clips(lengths, world_size: int) -> list[np.ndarray]:
    """Balance clips over ranks by audio length (longest-processing-time greedy).

    Returns one sorted index array per rank; every clip appears exactly once.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, np.int64)
    buckets: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        buckets[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def unshard(parts: list[np.ndarray], shards: list[np.ndarray]) -> np.ndarray:
    """Inverse of shard_clips for per-clip results stacked along axis 0."""
    n = sum(len(s) for s in shards)
    first = next(p for p in parts if len(p))
    out = np.empty((n,) + first.shape[1:], first.dtype)
    for p, s in zip(parts, shards):
        out[s] = p
    return out


def bind_to_gpu_cpus(device_index: int) -> dict:
    """Pin the calling process to the CPUs NVML reports as closest to the GPU (its NUMA node), so that the pinned
    staging buffers allocated afterwards sit on that node and H2D/D2H copies do not cross the socket interconnect.
    One process per GPU: call it before the first pinned allocation.  Returns what was done (for logs); never
    raises — without NVML or affinity support the process is left as it is."""
    import os
    info = {"device": int(device_index), "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
        info["cpus"] = len(allowed)
    except Exception as e:                                    # noqa: BLE001 (best effort by design)
        info["error"] = f"{type(e).__name__}: {e}"
    return info
