"""CPU, world_size 2, gloo: the only exchange step of the path (all-gather of embedding shards for
N x M scoring) and the clip sharding of the front end."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cosine as OC
from speech_transcript_embeddings_b200 import scoring, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = synth.embedding_pairs(37, 64, seed=3)          # ragged shards: 19 + 18 rows
        lo, hi = scoring.shard_rows(37, world, rank)
        a_loc, b_loc = torch.from_numpy(a[lo:hi]), torch.from_numpy(b[lo:hi])
        gathered = scoring.all_gather_rows(b_loc)
        score = lambda x, y: torch.from_numpy(OC.matrix_f64(x.numpy(), y.numpy()))
        stripe = scoring.sharded_cosine_matrix(a_loc, b_loc, _score=score)
        q.put((rank, lo, hi, gathered.numpy(), stripe.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_cosine_matrix_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = synth.embedding_pairs(37, 64, seed=3)
    full = OC.matrix_f64(a, b)
    S = np.empty_like(full)
    for rank, lo, hi, gathered, stripe in results:
        assert np.array_equal(gathered, b)                      # rank order, ragged shards
        S[lo:hi] = stripe
    assert np.abs(S - full).max() < 1e-15


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 8):
            spans = [scoring.shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
