"""N > 1 plumbing on CPU: contiguous sharding of the parameter sets and the single final gather, world_size 2,
gloo backend (the GPU path uses the same code with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relagn_b200.batch import shard_bounds, evaluate_sharded


class _StubEvaluator:
    """Stands in for BatchEvaluator on CPU: the 'spectrum' of a set is a deterministic function of its params."""

    nE = 8

    def evaluate_device(self, params, rel=True, out=None, attrs=None):
        e = torch.arange(8, dtype=torch.float64)
        v = params[:, :1] * 1000 + params[:, 2:3] + e[None, :]
        if out is not None:
            out.copy_(v)
            return out, None
        return v, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allp = torch.arange(P * 15, dtype=torch.float64).reshape(P, 15)
    lo, hi = shard_bounds(P, world, rank)
    out, gathered = evaluate_sharded(_StubEvaluator(), allp[lo:hi].contiguous(), gather_dst=0)
    if rank == 0:
        full = torch.cat(gathered, dim=0)
        ref, _ = _StubEvaluator().evaluate_device(allp)
        q.put(bool(torch.equal(full, ref)))
    else:
        assert gathered is None
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for P in (0, 1, 7, 8, 1000001):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(P, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == P
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_piece_edges_cover_the_shard():
    from relagn_b200.batch import piece_edges
    for P in (1, 5000, 12288, 12289, 20000, 125000, 1000000):
        for pieces in (None, 1, 4):
            e = piece_edges(P, pieces)
            assert e[0] == 0 and e[-1] == P and all(b >= a for a, b in zip(e, e[1:]))
        e = piece_edges(P)
        sizes = [b - a for a, b in zip(e, e[1:])]
        assert max(sizes) <= 65536 and sizes[-1] <= 12288            # the exposed transfer is the small last piece
        assert all(n % 8192 == 0 for n in sizes[:-1])


# 80000: four shrinking pieces per rank (40000 sets each), gathered piece by piece.  11 and 24577: shards of different
# size (6 + 5; 12289 + 12288 sets -- two pieces on rank 0 where rank 1 alone would plan one): padded to the largest
# shard with the same piece plan on every rank, trimmed on the destination
@pytest.mark.parametrize("P", [10, 11, 24577, 80000])
def test_two_rank_gather_gloo(P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
