"""Multi-GPU host logic on CPU: world_size-2 gloo run of the frame sharding + the single all-gather
(SURVEY.md §8e).  The data path has no collective; only results are gathered."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trackdlo_b200 import sharding


def test_shard_range_covers_all_frames():
    for F in (0, 1, 7, 64, 4096):
        for G in (1, 2, 3, 8):
            seen = []
            for r in range(G):
                lo, hi = sharding.shard_range(F, r, G)
                assert 0 <= lo <= hi <= F
                seen += list(range(lo, hi))
            assert seen == list(range(F))


def _worker(rank, world, port, F, Nn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, rank, world)
    g = torch.Generator().manual_seed(1234)
    Yall = torch.randn(F, Nn, 3, generator=g, dtype=torch.float64)
    s2 = torch.arange(F, dtype=torch.float64) * 1e-5
    it = torch.arange(F, dtype=torch.int32) % 50
    st = torch.arange(F, dtype=torch.int32) % 3
    Y, S, I, T = sharding.all_gather_results(Yall[lo:hi], s2[lo:hi], it[lo:hi], st[lo:hi], F)
    ok = torch.equal(Y, Yall) and torch.equal(S, s2) and torch.equal(I, it) and torch.equal(T, st)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 7, 5, out), nprocs=2, join=True)
    assert out[0] and out[1]
