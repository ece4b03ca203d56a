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
