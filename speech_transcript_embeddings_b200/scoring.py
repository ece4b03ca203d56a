"""Cosine scoring beyond the reference's pairwise call: the N x M matrix and its multi-GPU form.

The reference only ever scores pairs (R/processor.py:148-159, R/inference.py:121); BASELINE.json's
north_star adds the N x M matrix, assembled across ranks: the text-embedding shards are the only thing exchanged
on the whole path and each rank computes its [N/W, M] stripe.  Two implementations of the exchange:

  GatheredScorer           the product path on GPUs: the all-gather is fused into the kernels over NVLink peer
                           memory (stx_cosine_nxm_gathered: P2P stores from the split kernel, per-source flags
                           acquired by the tcgen05 GEMM's TMA producer)
  sharded_cosine_matrix    NCCL all-gather followed by the local N x M kernel (the baseline the fused path is
                           measured against; also the CPU/gloo plumbing test hook)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import torch
import torch.distributed as dist

from . import _lib, ops


def cosine_matrix(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """S = normalize(a) @ normalize(b).T on the current CUDA device, float32 [N, M]."""
    return ops.cosine_nxm(a.contiguous(), b.contiguous(), always_normalize=True)


def shard_rows(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, stop) of the rows rank owns; shards differ by at most one row."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, group=None, counts=None) -> torch.Tensor:
    """Concatenate every rank's [n_r, D] shard (ragged n_r allowed) in rank order.  ``counts`` (rows per rank),
    when the caller knows them, saves the small all-gather and its host synchronisation."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if counts is None:
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
                          _score: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] | None = None,
                          counts=None) -> torch.Tensor:
    """This rank's stripe S[rows of a_local, all M] of the global cosine matrix.

    ``b_local`` shards are all-gathered (NCCL over NVLink on GPUs); ``a_local`` stays local.
    ``_score`` is a test hook for the CPU (gloo) plumbing tests; the product path is ``cosine_matrix``.
    """
    b_all = all_gather_rows(b_local, group, counts)
    return (_score or cosine_matrix)(a_local, b_all)


class GatheredScorer:
    """Fused all-gather + N x M cosine scoring over NVLink peer memory (one instance per process group and shape).

    Holds the rank's symmetric buffer (``torch.distributed._symmetric_memory``: mapped into every peer of the
    group), the peers' addresses and the call epoch.  ``scorer(a_local, b_local)`` returns this rank's stripe
    ``S[rows of a_local, all M]``; rows are always L2-normalised.  Shard sizes may differ (``m_cap`` = the largest).
    """

    def __init__(self, m_cap: int, D: int, group=None, device=None, multicast: bool | None = None):
        import torch.distributed._symmetric_memory as symm_mem

        if not dist.is_initialized():
            raise _lib.StxError("GatheredScorer needs an initialised NCCL process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.m_cap, self.D = int(m_cap), int(D)
        lib = _lib.load()
        ws, symm = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.stx_cosine_gather_sizes(0, self.m_cap, self.world, self.D, C.byref(ws), C.byref(symm)),
                   "stx_cosine_gather_sizes")
        self.symm_bytes = int(symm.value)
        self.buf = symm_mem.empty(self.symm_bytes // 4, dtype=torch.float32, device=self.device)
        self.buf.zero_()                       # flags start at 0; epochs start at 1
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peers = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        # Push through the NVSwitch multicast address (one multimem.st reaches every peer) or with unicast P2P stores.
        # Measured on B200s (cfg5): 2 GPUs 0.124 ms multicast vs 0.107 unicast; 8 GPUs 0.078 vs 0.088 (NCCL all-gather
        # + GEMM: 0.121 / 0.084).  Default (None): multicast from 4 ranks up, when the fabric offers it.
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        want = (self.world >= 4) if multicast is None else bool(multicast)
        self.multicast_ptr = mc if (want and mc != 0) else None
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)            # every rank's buffer is zeroed before anyone pushes into it

    def counts(self, m_local: int) -> list[int]:
        """Rows per rank (one small all-gather; callers with fixed shards can pass ``counts`` to __call__)."""
        t = torch.tensor([m_local], dtype=torch.int64, device=self.device)
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [int(x.item()) for x in out]

    def __call__(self, a_local: torch.Tensor, b_local: torch.Tensor, counts=None, out: torch.Tensor | None = None):
        lib = _lib.load()
        ops._require_cuda(a_local, "a_local", torch.float32)
        ops._require_cuda(b_local, "b_local", torch.float32)
        if a_local.dim() != 2 or b_local.dim() != 2 or a_local.shape[1] != self.D or b_local.shape[1] != self.D:
            raise ValueError(f"expected [n, {self.D}] and [m, {self.D}]")
        counts = list(counts) if counts is not None else self.counts(b_local.shape[0])
        if len(counts) != self.world or counts[self.rank] != b_local.shape[0] or max(counts) > self.m_cap:
            raise ValueError("counts do not describe the shards (or exceed m_cap)")
        n, M = a_local.shape[0], sum(counts)
        if out is None:
            out = torch.empty((n, M), dtype=torch.float32, device=self.device)
        ws_b, symm = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.stx_cosine_gather_sizes(n, self.m_cap, self.world, self.D, C.byref(ws_b), C.byref(symm)),
                   "stx_cosine_gather_sizes")
        ws = torch.empty(int(ws_b.value), dtype=torch.uint8, device=self.device)
        h_counts = (C.c_int32 * self.world)(*counts)
        self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(lib.stx_cosine_nxm_gathered(a_local.data_ptr(), b_local.data_ptr(), n, self.D, self.world,
                                                   self.rank, h_counts, self.m_cap, self.peers, self.multicast_ptr,
                                                   self.epoch,
                                                   out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   ops._stream_ptr(self.device)), "stx_cosine_nxm_gathered")
        # no barrier here: the symmetric buffer holds two sets used by epoch parity (see include/stx_b200.h)
        return out
