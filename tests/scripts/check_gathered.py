"""Multi-GPU check + timing of the fused all-gather + N x M cosine path (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/scripts/check_gathered.py

Every rank checks its stripe against the float64 oracle (checker only) and against the NCCL all-gather path, for
equal and ragged shards; rank 0 prints one JSON line with both timings (CUDA events, max over ranks).
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import cosine as OC  # noqa: E402  (checker)
from speech_transcript_embeddings_b200 import scoring, synth  # noqa: E402

world = int(os.environ["WORLD_SIZE"])
rank = int(os.environ["RANK"])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
results = {}
for name, N, M, D, ragged in (("cfg5", 4096, 4096, 768, False), ("ragged", 1000, 1003, 100, True), ("cfg5_d1024", 4096, 4096, 1024, False)):
    a, b = synth.embedding_pairs(max(N, M), D, seed=3)
    a, b = a[:N] * np.float32(1.3), b[:M]
    if ragged:
        cuts = np.linspace(0, M, world + 1).astype(int)
        cuts[1:-1] += 7 * (np.arange(1, world) % 2)             # unequal shards
    else:
        cuts = np.array([scoring.shard_rows(M, world, r)[0] for r in range(world)] + [M])
    lo_a, hi_a = scoring.shard_rows(N, world, rank)
    a_loc = torch.from_numpy(a[lo_a:hi_a]).to(dev)
    b_loc = torch.from_numpy(b[cuts[rank]:cuts[rank + 1]]).to(dev)
    counts = [int(cuts[r + 1] - cuts[r]) for r in range(world)]
    scorer = scoring.GatheredScorer(max(counts), D, device=dev, multicast=False)
    mc = scoring.GatheredScorer(max(counts), D, device=dev, multicast=True)
    S = scorer(a_loc, b_loc, counts=counts)
    assert torch.equal(mc(a_loc, b_loc, counts=counts), S)      # multicast and unicast pushes give the same bits
    S2 = scorer(a_loc, b_loc)                                    # second call: epoch 2, counts via all-gather
    ref = OC.matrix_f64(a[lo_a:hi_a][:96], b)
    err = float(np.abs(S[:96].cpu().numpy() - ref).max())
    nccl = scoring.sharded_cosine_matrix(a_loc, b_loc)
    diff = float((S - nccl).abs().max().item())
    same = bool(torch.equal(S, S2))
    assert err <= 1e-5 and diff <= 5e-6 and same, (name, rank, err, diff, same)

    def timed(f, iters=30, warm=5):
        for _ in range(warm):
            f()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = torch.empty((hi_a - lo_a, M), dtype=torch.float32, device=dev)
    ms_fused = timed(lambda: scorer(a_loc, b_loc, counts=counts, out=out))
    ms_mc = timed(lambda: mc(a_loc, b_loc, counts=counts, out=out))
    ms_nccl = timed(lambda: scoring.sharded_cosine_matrix(a_loc, b_loc, counts=counts))
    results[name] = {"N": N, "M": M, "D": D, "max_abs_err_vs_f64": err, "max_abs_diff_vs_nccl_path": diff,
                     "ms_fused_unicast": ms_fused, "ms_fused_multicast": ms_mc if mc.multicast_ptr is not None else None,
                     "ms_nccl_allgather_then_gemm": ms_nccl,
                     "scores_per_s_fused": N * M / (min(ms_fused, ms_mc) * 1e-3)}
    del scorer, mc
# ---- empty shards, back to back: fewer rows than ranks on both sides, new inputs every call, no barrier in between.  A rank
# with no local rows (no GEMM) and slots with no rows (never read by a GEMM) must still take part in the flag protocol, or a
# fast rank's push of call e + 2 lands on planes a slow peer is still reading for call e.
D = 96
N, M = world - 1, max(1, world // 2)                            # the last rank has no a rows; half the ranks have no b rows
lo_a, hi_a = scoring.shard_rows(N, world, rank)
lo_b, hi_b = scoring.shard_rows(M, world, rank)
counts = [scoring.shard_rows(M, world, r)[1] - scoring.shard_rows(M, world, r)[0] for r in range(world)]
scorer = scoring.GatheredScorer(max(max(counts), 1), D, device=dev)
calls = []
for it in range(6):
    a, b = synth.embedding_pairs(max(N, M), D, seed=100 + it)
    a_loc = torch.from_numpy(a[lo_a:hi_a]).to(dev)
    b_loc = torch.from_numpy(b[:M][lo_b:hi_b]).to(dev)
    calls.append((a, b, scorer(a_loc, b_loc, counts=counts)))   # back to back: nothing synchronises between the calls
    if rank % 2 == it % 2:
        torch.cuda._sleep(2_000_000)                            # skew the ranks (about a millisecond)
worst_empty = 0.0
for a, b, S in calls:
    assert tuple(S.shape) == (hi_a - lo_a, M)
    if hi_a > lo_a:
        worst_empty = max(worst_empty, float(np.abs(S.cpu().numpy() - OC.matrix_f64(a[lo_a:hi_a], b[:M])).max()))
assert worst_empty <= 1e-5, (rank, worst_empty)
results["empty_shards"] = {"N": N, "M": M, "D": D, "calls": len(calls), "max_abs_err_vs_f64": worst_empty}
del scorer
if rank == 0:
    print(json.dumps({"n_gpus": world, "results": results}), flush=True)
dist.destroy_process_group()
