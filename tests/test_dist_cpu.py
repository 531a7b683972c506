"""Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo processes (no GPU): the halo plan of the node-range
partition and the flat gradient all-reduce of the graph-batch sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnn_b200
from gnn_b200 import dist_graph


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ret = mp.get_context('spawn').Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _halo_case(rank, world, local):
    n, deg = 40, 3
    rng = np.random.default_rng(5)
    dst = np.repeat(np.arange(n), deg)
    src = (dst + rng.integers(-3, 4, dst.shape[0])) % n if local else rng.integers(0, n, dst.shape[0])
    bounds = dist_graph.partition_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = (dst >= lo) & (dst < hi)
    plan = dist_graph.HaloPlan(torch.as_tensor(src[mine]), bounds, rank, world)
    truth = torch.arange(n * 4, dtype=torch.float32).view(n, 4)          # row r holds 4r .. 4r+3
    x = torch.full((n, 4), -1.0)
    x[lo:hi] = truth[lo:hi]                                               # own rows valid, the rest garbage
    plan.exchange(x)
    needed = np.unique(src[mine])
    ok = bool(torch.equal(x[torch.as_tensor(needed)], truth[torch.as_tensor(needed)]))
    flag = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    return ok, plan.use_allgather, int(flag), plan.bytes_received_per_exchange(16)


def _halo_local(rank, world): return _halo_case(rank, world, True)


def _halo_uniform(rank, world): return _halo_case(rank, world, False)


def test_halo_plan_local_sources_uses_all_to_all():
    res = _run(_halo_local)
    for rank in (0, 1):
        ok, allgather, flag, nbytes = res[rank]
        assert ok and not allgather and flag == 1
        assert 0 < nbytes < 20 * 16          # only the boundary rows travel


def test_halo_plan_uniform_sources_uses_all_gather():
    res = _run(_halo_uniform)
    for rank in (0, 1):
        ok, allgather, flag, nbytes = res[rank]
        assert ok and allgather and flag == 1 and nbytes == 20 * 16


def _grad_allreduce(rank, world):
    grads = [torch.full((3, 2), float(rank + 1)), torch.full((2,), 10.0 * (rank + 1))]
    out = dist_graph.allreduce_gradients(grads)
    return [o.tolist() for o in out]


def test_flat_gradient_allreduce():
    res = _run(_grad_allreduce)
    for rank in (0, 1):
        assert res[rank][0] == [[3.0, 3.0]] * 3 and res[rank][1] == [30.0, 30.0]


def test_partition_bounds_and_sharding():
    assert dist_graph.partition_bounds(10, 4) == [0, 3, 6, 9, 10]
    assert dist_graph.partition_bounds(1_000_000, 8)[1] == 125_000
    graphs = list(range(10))
    parts = [dist_graph.shard_graphs(graphs, r, 3) for r in range(3)]
    assert sum(parts, []) == graphs and max(map(len, parts)) - min(map(len, parts)) <= 1
