"""CPU, world_size 2, gloo: the N>1 host path (sharding + the single gather of [B_local,S] relevance)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, S, q):
    sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from lxt_b200 import dist as ldist
    r, w, _ = ldist.init_from_env("gloo")
    lo, hi = ldist.shard_range(n_total, r, w)
    # stand-in for the engine: relevance[i, s] = 1000*i + s, so order and raggedness are checkable
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * 1000 + torch.arange(S, dtype=torch.float32)[None, :]
    full = ldist.gather_relevance(local, n_total, w)
    q.put((r, full))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_total", [4, 5])
def test_gather_relevance_world2_gloo(n_total):
    S, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:                  # a port the kernel just handed out is free (a pid-derived one can collide)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, S, q)) for r in range(world)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    expect = torch.arange(n_total, dtype=torch.float32)[:, None] * 1000 + torch.arange(S, dtype=torch.float32)[None, :]
    for r in range(world):
        assert torch.equal(got[r], expect)
