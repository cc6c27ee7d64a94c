"""world_size-2 checks of the multi-GPU plumbing on CPU (gloo): sharding, the final gather and the
merge back into global trajectory order.  The GPU path uses the same functions with NCCL tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numba_integrators_b200 import distributed as nd


def test_shard_indices_cover_everything():
    for N, world in ((10, 2), (11, 4), (5, 8), (1000, 8)):
        for inter in (False, True):
            idx = np.concatenate([nd.shard_indices(N, r, world, inter) for r in range(world)])
            assert sorted(idx.tolist()) == list(range(N))
    assert nd.shard_indices(10, 1, 2, interleaved=True).tolist() == [1, 3, 5, 7, 9]
    assert nd.shard_indices(10, 1, 2).tolist() == [5, 6, 7, 8, 9]


def test_merge_is_inverse_of_shard():
    y = np.arange(33 * 3, dtype=np.float64).reshape(33, 3)
    for inter in (False, True):
        parts = [nd.shard({'y': y}, r, 4, inter)[0]['y'] for r in range(4)]
        assert np.array_equal(nd.merge_shards(parts, 33, 4, inter), y)


def _worker(rank, world, port, interleaved, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        N = 37                                     # ragged: shards of unequal length
        y = np.arange(N * 3, dtype=np.float64).reshape(N, 3)
        acc = np.arange(N, dtype=np.int32) * 7
        mine, idx = nd.shard({'y': y, 'n_accepted': acc}, rank, world, interleaved)
        got = nd.all_gather_arrays(mine)
        full = {k: nd.merge_shards([p.numpy() for p in parts], N, world, interleaved) for k, parts in got.items()}
        ok = np.array_equal(full['y'], y) and np.array_equal(full['n_accepted'], acc)
        total = torch.tensor([float(mine['n_accepted'].sum())], dtype=torch.float64)
        dist.all_reduce(total)                     # the bench's whole-job aggregate
        ok = ok and total.item() == float(acc.sum())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('interleaved', [False, True])
def test_gather_world_size_2_gloo(interleaved):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, interleaved, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
