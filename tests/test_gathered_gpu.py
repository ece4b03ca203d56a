"""GPU, >= 2 devices: the fused all-gather + N x M cosine path (stx_cosine_nxm_gathered) under torchrun.

Skipped on single-GPU boxes; `gpurun --gpus 2 -- python -m pytest tests/test_gathered_gpu.py -m gpu` runs it.  The
worker (tests/scripts/check_gathered.py) asserts on every rank: <= 1e-5 vs the float64 oracle, <= 5e-6 vs the NCCL
all-gather path, equal and ragged shards, repeated calls (epochs)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_fused_gather_matches_oracle_and_nccl_path():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4 if n < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "scripts" / "check_gathered.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == world and set(res["results"]) == {"cfg5", "ragged", "cfg5_d1024", "empty_shards"}
    for v in res["results"].values():
        assert v["max_abs_err_vs_f64"] <= 1e-5
