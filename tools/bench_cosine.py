"""cfg5: inference-style scoring, N = M = 4096 embedding pairs, all-gathered N x M cosine matrix.

    python tools/bench_cosine.py                       # 1 GPU: full 4096 x 4096 matrix
    torchrun --nproc-per-node 8 tools/bench_cosine.py  # 8 GPUs: [512, D] shards, all-gather fused into the kernels
                                                       # over NVLink peer memory, [512, 4096] stripes

Prints one JSON line per D in {768, 1024}: time (CUDA events, max over ranks), TFLOP/s, GB/s, max-abs error vs
a float64 product on a sample of rows.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from speech_transcript_embeddings_b200 import _lib, scoring, synth  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = M = 4096
iters, warm = 20, 5
for D in (768, 1024):
    a, b = synth.embedding_pairs(N, D, seed=0)
    lo, hi = scoring.shard_rows(N, world, rank)
    a_loc = torch.from_numpy(a[lo:hi]).to(dev)
    b_loc = torch.from_numpy(b[lo:hi]).to(dev)

    counts = [scoring.shard_rows(M, world, r)[1] - scoring.shard_rows(M, world, r)[0] for r in range(world)]
    scorer = scoring.GatheredScorer(max(counts), D, device=dev) if world > 1 else None

    def step():
        if world > 1:
            return scorer(a_loc, b_loc, counts=counts)          # all-gather fused into the kernels over NVLink
        return scoring.cosine_matrix(a_loc, b_loc)

    for _ in range(warm):
        S = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        S = step()
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    ms_nccl = None
    if world > 1:                                               # the baseline: NCCL all-gather, then the local kernel
        for _ in range(warm):
            scoring.sharded_cosine_matrix(a_loc, b_loc, counts=counts)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            scoring.sharded_cosine_matrix(a_loc, b_loc, counts=counts)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_nccl = float(t.item())
    a64, b64 = a[lo:lo + 64].astype(np.float64), b.astype(np.float64)       # float64 check of 64 rows (not the oracle: tools/
    ref = (a64 / np.maximum(np.linalg.norm(a64, axis=1, keepdims=True), 1e-12)) @ \
          (b64 / np.maximum(np.linalg.norm(b64, axis=1, keepdims=True), 1e-12)).T   # never imports oracle/)
    err = float(np.abs(S[:64].cpu().numpy() - ref).max())
    if rank == 0:
        flop = 2.0 * N * M * D
        byts = 4.0 * (N * D + M * D + N * M)
        print(json.dumps({"workload": f"cfg5 cosine N=M={N} D={D}", "n_gpus": world, "ms": ms,
                          "tflops": flop / ms / 1e9, "gbs": byts / ms / 1e6, "max_abs_err_vs_f64": err,
                          "kernel_launches_per_call": launches / iters, "ms_nccl_allgather_then_gemm": ms_nccl,
                          "scores_per_s": N * M / (ms * 1e-3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
