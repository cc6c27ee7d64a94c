"""Multi-GPU plumbing: one process per GPU (torchrun), trajectories sharded by index.

The integration itself needs no communication -- trajectories are independent.
The only collective is the final gather of terminal states and statistics, done with
`torch.distributed.all_gather` (NCCL over NVLink on GPUs; gloo in the CPU tests) directly on the
ensemble's device buffers (zero-copy views through __cuda_array_interface__).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib

FIELDS = ('t', 'y', 'auxiliary', 'n_accepted', 'n_rejected', 'status')


def rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def shard_indices(N, rank, world, interleaved=False):
    """Global trajectory indices owned by `rank`: contiguous blocks for random ensembles, `i mod world`
    for sorted parameter sweeps (balances cost, BASELINE config C4)."""
    if interleaved:
        return np.arange(rank, N, world)
    per = (N + world - 1) // world
    return np.arange(min(rank * per, N), min((rank + 1) * per, N))


def shard(arrays, rank, world, interleaved=False):
    """Slice every (N, ...) array of `arrays` (dict or sequence) down to this rank's trajectories."""
    first = next(iter(arrays.values())) if isinstance(arrays, dict) else arrays[0]
    idx = shard_indices(len(first), rank, world, interleaved)
    if isinstance(arrays, dict):
        return {k: (None if v is None else np.ascontiguousarray(v[idx])) for k, v in arrays.items()}, idx
    return [None if v is None else np.ascontiguousarray(v[idx]) for v in arrays], idx


class _DeviceView:
    """Minimal __cuda_array_interface__ carrier for a raw device pointer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


def device_tensors(ens):
    """Zero-copy torch views of the ensemble's terminal-state buffers (device layout)."""
    import torch
    dev = torch.device('cuda', ens.device)
    N, n, Q = ens.N, ens.n, ens.Q
    small = ens.kind == 0
    shapes = {
        't': (_lib.F_T, (N,), '<f8'),
        'y': (_lib.F_Y, (n, N) if small else (N, n), '<f8'),
        'n_accepted': (_lib.F_N_ACCEPTED, (N,), '<i4'),
        'n_rejected': (_lib.F_N_REJECTED, (N,), '<i4'),
        'status': (_lib.F_STATUS, (N,), '<i4'),
    }
    if Q:
        shapes['auxiliary'] = (_lib.F_AUX, (Q, N) if small else (N, Q), '<f8')
    out = {}
    for name, (field, shape, ts) in shapes.items():
        out[name] = torch.as_tensor(_DeviceView(ens.device_pointer(field), shape, ts), device=dev)
    return out, small


def all_gather_arrays(local: dict, group=None):
    """all_gather a dict of equally-shaped-per-rank tensors/arrays; returns {name: [rank0, rank1, ...]}.
    Shards may have different lengths along axis 0 (padded to the maximum for the collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = {}
    for name, v in local.items():
        tns = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
        n0 = torch.tensor([tns.shape[0]], dtype=torch.int64, device=tns.device)
        sizes = [torch.zeros_like(n0) for _ in range(world)]
        dist.all_gather(sizes, n0, group=group)
        sizes = [int(s.item()) for s in sizes]
        m = max(sizes)
        if tns.shape[0] < m:
            pad = torch.zeros((m - tns.shape[0],) + tuple(tns.shape[1:]), dtype=tns.dtype, device=tns.device)
            tns = torch.cat([tns, pad], dim=0)
        bufs = [torch.empty_like(tns) for _ in range(world)]
        dist.all_gather(bufs, tns.contiguous(), group=group)
        out[name] = [b[:s] for b, s in zip(bufs, sizes)]
    return out


def merge_shards(parts, N, world, interleaved=False):
    """Inverse of shard_indices: list of per-rank (N_r, ...) arrays -> one (N, ...) array in global order."""
    first = parts[0]
    out = np.empty((N,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r, p in enumerate(parts):
        out[shard_indices(N, r, world, interleaved)] = p
    return out


def gather(ens, N_global=None, interleaved=False, group=None):
    """Final gather of terminal t, y, auxiliary and the step statistics of a sharded ensemble.

    Every rank receives the full arrays in global trajectory order (numpy, host).  On GPUs the
    collective runs on the ensemble's own device buffers over NCCL; one D2H copy follows."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ens.sync()
    tens, small = device_tensors(ens)
    # trajectory axis first for the collective
    local = {k: (v.t().contiguous() if small and v.dim() == 2 else v) for k, v in tens.items()}
    got = all_gather_arrays(local, group)
    sizes = [len(p) for p in got['t']]
    N = N_global if N_global is not None else sum(sizes)
    return {k: merge_shards([p.cpu().numpy() for p in parts], N, world, interleaved) for k, parts in got.items()}
