"""N > 1 path on CPU: two gloo ranks shard a batch, each computes its shard (the oracle stands in for the GPU engine,
which needs a B200), results are gathered with the same helper bench.py uses, and must equal the unsharded result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT, oracle_smul, random_scalars


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libecc_b200.sharding import gather_results, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = random_scalars("SECP256R1", n, tag=90)
    lo, hi = shard_bounds(n, rank, world)
    out, st = oracle_smul("SECP256R1", sc[lo:hi], nthreads=2)
    full_out = gather_results(torch.from_numpy(out).view(-1), n, 64)
    full_st = gather_results(torch.from_numpy(st.view(np.uint8)).view(-1), n, 1)
    if rank == 0:
        q.put((full_out.numpy().copy(), full_st.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_out, got_st = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, wst = oracle_smul("SECP256R1", random_scalars("SECP256R1", n, tag=90), nthreads=4)
    assert (got_out.reshape(n, 64) == want).all()
    assert (got_st.view(np.int8) == wst).all()


def test_two_rank_even_shards():
    _run(64)


def test_two_rank_ragged_shards():
    _run(37)


def test_shard_bounds_cover_everything():
    from libecc_b200.sharding import shard_bounds
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 5):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
