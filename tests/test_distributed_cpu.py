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


def _records(N, n, Q):
    """Synthetic terminal records in the layout of ni_ensemble_pack_terminal (include/ni_b200.h)."""
    from numba_integrators_b200._solver import terminal_dtype
    rec = np.zeros(N, dtype=terminal_dtype(n, Q))
    rec['t'] = np.arange(N) * 0.5
    rec['y'] = np.arange(N * n, dtype=np.float64).reshape(N, n)
    if Q:
        rec['aux'] = -np.arange(N * Q, dtype=np.float64).reshape(N, Q)
    rec['n_accepted'] = np.arange(N) * 7
    rec['n_rejected'] = np.arange(N) % 5
    rec['status'] = 1
    return rec


def test_terminal_record_layout():
    from numba_integrators_b200._solver import terminal_dtype
    dt = terminal_dtype(3, 1)  # Lorenz-63 through Advanced: the 52-byte record of SURVEY.md section 8e
    assert dt.itemsize == 52 and dt.names == ('t', 'y', 'aux', 'n_accepted', 'n_rejected', 'status')
    assert [dt.fields[k][1] for k in dt.names] == [0, 8, 32, 40, 44, 48]
    assert terminal_dtype(2, 0).itemsize == 36 and 'aux' not in terminal_dtype(2, 0).names


def test_merge_records_is_inverse_of_sharding():
    N, world = 37, 4
    rec = _records(N, 3, 1)
    rows = nd.shard_rows(N, world)
    for inter in (False, True):
        slots = np.zeros((world, rows), dtype=rec.dtype)
        for r in range(world):
            idx = nd.shard_indices(N, r, world, inter)
            slots[r, :len(idx)] = rec[idx]
        full = nd.merge_records(slots, N, world, inter)
        assert np.array_equal(full['y'], rec['y']) and np.array_equal(full['auxiliary'], rec['aux'])
        assert np.array_equal(full['n_accepted'], rec['n_accepted']) and np.array_equal(full['t'], rec['t'])


def _worker(rank, world, port, interleaved, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        N = 37                                     # ragged: shards of unequal length, equal-sized slots
        rec = _records(N, 3, 1)
        idx = nd.shard_indices(N, rank, world, interleaved)
        slots = nd.all_gather_records(rec[idx], N)  # ONE collective, no size exchange
        full = nd.merge_records(slots, N, world, interleaved)
        ok = all(np.array_equal(full[k], rec[s]) for k, s in (('t', 't'), ('y', 'y'), ('auxiliary', 'aux'),
                                                             ('n_accepted', 'n_accepted'), ('n_rejected', 'n_rejected'),
                                                             ('status', 'status')))
        total = torch.tensor([float(rec['n_accepted'][idx].sum())], dtype=torch.float64)
        dist.all_reduce(total)                     # the bench's whole-job aggregate
        ok = ok and total.item() == float(rec['n_accepted'].sum())
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


def test_merge_of_padded_sink_records():
    """Rows of the in-kernel terminal sink (ni_ensemble_set_terminal_sink / FusedGather) are the packed record padded to
    16-byte vectors: the host-side merge must read them through the padded dtype and ignore the padding."""
    from numba_integrators_b200._solver import terminal_dtype
    n, Q, N, world = 3, 1, 11, 3
    dt = terminal_dtype(n, Q)
    assert dt.itemsize == 52
    padded = np.dtype({'names': list(dt.names), 'formats': [dt.fields[k][0] for k in dt.names],
                       'offsets': [dt.fields[k][1] for k in dt.names], 'itemsize': 64})
    full = _records(N, n, Q)
    rows = nd.shard_rows(N, world)
    for inter in (False, True):
        raw = np.full((world, rows * 64), 0xEE, dtype=np.uint8)       # padding bytes hold garbage
        for r in range(world):
            idx = nd.shard_indices(N, r, world, inter)
            view = raw[r].view(padded)
            for name in dt.names:
                view[name][:len(idx)] = full[name][idx]
        out = nd.merge_records(raw.view(padded).reshape(world, rows), N, world, inter)
        assert np.array_equal(out['y'], full['y']) and np.array_equal(out['t'], full['t'])
        assert np.array_equal(out['auxiliary'], full['aux']) and np.array_equal(out['n_accepted'], full['n_accepted'])
        assert np.array_equal(out['status'], full['status'])
