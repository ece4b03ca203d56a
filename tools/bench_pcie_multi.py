"""Concurrent H2D / D2H rates of all ranks of one box, with and without binding each rank to its GPU's CPUs
(NVML affinity) before the pinned allocation.  Run under torchrun.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_pcie_multi.py
"""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import sharding  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 * 480000


def rates(tag):
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    h2 = torch.empty(n // 2, dtype=torch.float32, pin_memory=True)
    d2 = torch.empty(n // 2, dtype=torch.float32, device=dev)
    out = {}
    for name, f, nbytes in (("h2d", lambda: d.copy_(h, non_blocking=True), n * 4),
                            ("d2h", lambda: h2.copy_(d2, non_blocking=True), n * 2)):
        for _ in range(3):
            f()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        t = torch.tensor([nbytes / dt / 1e9], dtype=torch.float64, device=dev)
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        out[name] = {"min_GBps": round(float(lo.item()), 1), "sum_GBps": round(float(sm.item()), 1)}
    return {tag: out}


res = {"world": world, "cpus": os.cpu_count(), "affinity_before": len(os.sched_getaffinity(0))}
res.update(rates("unbound"))
info = sharding.bind_to_gpu_cpus(local)
res["bind"] = info if rank == 0 else None
res.update(rates("bound"))
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
