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
    it = torch.stack([torch.arange(F, dtype=torch.int32) % 50, torch.arange(F, dtype=torch.int32) % 7], dim=1)
    st = torch.arange(F, dtype=torch.int32) % 3
    # what the kernel's epilogue writes for this rank's frames (tdlo_track_batch::packed_results)
    packed = sharding.alloc_packed(F, Nn, world, "cpu")
    n = hi - lo
    packed[:n, :3 * Nn] = Yall[lo:hi].reshape(n, 3 * Nn)
    packed[:n, 3 * Nn] = s2[lo:hi]
    packed[:n, 3 * Nn + 1:3 * Nn + 3] = it[lo:hi].to(torch.float64)
    packed[:n, 3 * Nn + 3] = st[lo:hi].to(torch.float64)
    Y, S, I, T = sharding.unpack(sharding.all_gather_packed(packed), F, Nn)
    ok = torch.equal(Y, Yall) and torch.equal(S, s2) and torch.equal(I, it) and torch.equal(T, st)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 8, 5, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_record_width_matches_header():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "trackdlo_b200.h")).read()
    assert re.search(r"\[n_frames\]\[3\*n_nodes \+ 4\]", hdr)
    assert sharding.record_width(50) == 154
