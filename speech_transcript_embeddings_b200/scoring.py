"""Cosine scoring beyond the reference's pairwise call: the N x M matrix and its multi-GPU form.

The reference only ever scores pairs (R/processor.py:148-159, R/inference.py:121); BASELINE.json's
north_star adds the N x M matrix, assembled across ranks with one NCCL all-gather of the embedding
shards (the only exchange step on the whole path).  Each rank computes its [N/W, M] stripe.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist

from . import ops


def cosine_matrix(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """S = normalize(a) @ normalize(b).T on the current CUDA device, float32 [N, M]."""
    return ops.cosine_nxm(a.contiguous(), b.contiguous(), always_normalize=True)


def shard_rows(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, stop) of the rows rank owns; shards differ by at most one row."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate every rank's [n_r, D] shard (ragged n_r allowed) in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    padded = local
    if local.shape[0] < n_max:
        padded = torch.cat([local, local.new_zeros(n_max - local.shape[0], local.shape[1])])
    gathered = torch.empty((world * n_max, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(c == n_max for c in counts):
        return gathered
    return torch.cat([gathered[r * n_max:r * n_max + c] for r, c in enumerate(counts)])


def sharded_cosine_matrix(a_local: torch.Tensor, b_local: torch.Tensor, group=None,
                          _score: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] | None = None) -> torch.Tensor:
    """This rank's stripe S[rows of a_local, all M] of the global cosine matrix.

    ``b_local`` shards are all-gathered (NCCL over NVLink on GPUs); ``a_local`` stays local.
    ``_score`` is a test hook for the CPU (gloo) plumbing tests; the product path is ``cosine_matrix``.
    """
    b_all = all_gather_rows(b_local, group)
    return (_score or cosine_matrix)(a_local, b_all)
