"""N > 1 host logic on CPU: two gloo ranks shard a work list, MAX-reduce a timing, gather result maps."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cds_mvsnet_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_worklist(n_items, rank, world)
        # each rank "computes" a map per work item: value = item id
        local = torch.stack([torch.full((2, 3), float(i)) for i in mine]) if mine else torch.zeros(0, 2, 3)
        full = parallel.gather_maps(local, n_items)
        t = parallel.max_over_ranks(10.0 + rank, torch.device("cpu"))
        ret[rank] = (mine, full[:, 0, 0].tolist(), t)
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_gather():
    world, n_items = 2, 5
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert sorted(r0[0] + r1[0]) == list(range(n_items))          # every item exactly once
    assert not set(r0[0]) & set(r1[0])
    assert r0[1] == r1[1] == [float(i) for i in range(n_items)]    # gathered back in work-list order
    assert r0[2] == r1[2] == 11.0                                  # slowest rank's time


def test_shard_worklist_edges():
    assert parallel.shard_worklist(0, 0, 4) == []
    assert parallel.shard_worklist(3, 3, 4) == []
    assert [len(parallel.shard_worklist(10, r, 4)) for r in range(4)] == [3, 3, 2, 2]


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                   # replicas: identical parameters on every rank
        a, b, frozen = torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2), requires_grad=False)
        a.grad = torch.full((3, 4), float(rank + 1))           # rank-dependent gradients
        if rank == 0:
            b.grad = torch.arange(5.0)                         # b has no gradient on rank 1
        n = parallel.all_reduce_gradients([a, b, frozen])
        ret[rank] = (n, a.grad.tolist(), b.grad.tolist(), frozen.grad is None)
    finally:
        dist.destroy_process_group()


def test_two_ranks_average_gradients():
    """The one exchange step of data-parallel training: a flat all-reduce of the gradients."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0 == r1
    assert r0[0] == 17 and r0[3]
    assert r0[1] == [[1.5] * 4] * 3                            # mean of 1 and 2
    assert r0[2] == [0.0, 0.5, 1.0, 1.5, 2.0]                  # mean of arange(5) and the missing gradient (zeros)


def test_average_gradients_single_process_is_a_noop():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert parallel.all_reduce_gradients([p]) == 3 and p.grad.tolist() == [2.0, 2.0, 2.0]
