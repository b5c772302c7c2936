"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding of the
batch and the single sum all-reduce of the packed integer counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, T, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from challenge_b200 import dist as D
    from oracle import metrics as M
    rng = np.random.default_rng(123)          # same global tensors on every rank
    yt = (rng.random((B, T, 3)) < 0.3).astype(np.float32)
    yt = np.repeat(yt[:, ::8], 8, axis=1)[:, :T]
    yp = np.clip(yt + rng.normal(0, 0.35, yt.shape), 0, 1).astype(np.float32)
    lo, hi = D.shard_range(B, world, rank)
    nt, npd, co = M.er_parts(yt[lo:hi], yp[lo:hi])          # this rank's slice only
    triples = torch.tensor(np.stack([nt, npd, co], 1), dtype=torch.int32)
    tpfpfn = torch.tensor(M.f1_counts(yt[lo:hi], yp[lo:hi]), dtype=torch.int64)
    buf = D.allreduce_counts(D.pack_counts(tpfpfn, triples, B, lo))
    g_tpfpfn, g_triples = D.unpack_counts(buf)
    # the reduced result equals the single-process result on the whole batch
    nt, npd, co = M.er_parts(yt, yp)
    ok = np.array_equal(g_triples.numpy(), np.stack([nt, npd, co], 1))
    ok &= tuple(g_tpfpfn.tolist()) == M.f1_counts(yt, yp)
    ok &= np.array_equal(D.er_from_triples(g_triples), M.er_from_parts(nt, npd, co), equal_nan=True)
    ok &= np.isclose(D.f1_from_counts(*g_tpfpfn.tolist()), M.f1_from_counts(*M.f1_counts(yt, yp)))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_and_allreduce_counts_world2():
    from challenge_b200.dist import shard_range
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(8192, 8, 7) == (7168, 8192)
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 11, 96, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
