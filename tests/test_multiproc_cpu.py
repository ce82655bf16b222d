"""CPU, world_size 2, gloo: the host side of the multi-GPU path (row sharding + the one all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from batman_b200.uq import allreduce_totals, shard_rows
from oracle import functions as ofun, philox, saltelli


def test_shards_partition_the_base_rows():
    for size in (0, 1, 2, 7, 1000, 10 ** 8 + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_rows(size, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == size
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert all(hi >= lo for lo, hi in edges)
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1


def _sums(Y, N, p):
    """The raw Saltelli sums of one shard (what each rank's partial slots add up to), shift 0."""
    yA, yB = Y[0:N, 0], Y[N:2 * N, 0]
    out = [yA.sum(), yB.sum(), (yA * yA).sum(), (yB * yB).sum(), (yA * yB).sum()]
    for i in range(p):
        yE = Y[(2 + i) * N:(3 + i) * N, 0]
        out += [yE.sum(), (yE * yE).sum(), (yB * yE).sum(), (yA * yE).sum()]
    return np.array(out)


def _worker(rank, world, port, N, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_rows(N, world, rank)
    dists = [('uniform', -np.pi, np.pi)] * 3
    A, B = philox.base_samples(123456, lo, hi, dists)          # generator is a function of the absolute row
    Y = ofun.ishigami(saltelli.design_from_base(A, B, second_order=False))
    t = torch.from_numpy(_sums(Y, hi - lo, 3))
    allreduce_totals(t)
    if rank == 0:
        q.put(t.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process():
    N = 1001
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, B = philox.base_samples(123456, 0, N, [('uniform', -np.pi, np.pi)] * 3)
    want = _sums(ofun.ishigami(saltelli.design_from_base(A, B, second_order=False)), N, 3)
    np.testing.assert_allclose(got, want, rtol=1e-13)
