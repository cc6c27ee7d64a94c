"""Multi-GPU plumbing: one process per GPU (torchrun), trajectories sharded by index.

The integration itself needs no communication -- trajectories are independent.
The only collective is the final gather of terminal states and statistics: one packed record per
trajectory (`ni_ensemble_pack_terminal`), one `all_gather_into_tensor` of equal-sized slots over NCCL
(NVLink / NVSwitch); the CPU tests run the same record layout and merge logic over gloo.
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


def _slot_bytes(rows, dtype):
    """Bytes of one rank's slot in the gather buffer: the records, padded to 16 bytes (vector stores over NVLink)."""
    return (rows * dtype.itemsize + 15) // 16 * 16


def shard_rows(N, world):
    """Rows of the gather buffer per rank: the largest shard (equal-sized slots, no size exchange: every rank
    derives every shard's length from N and the sharding rule)."""
    return (N + world - 1) // world


class TerminalGather:
    """The one collective of the path: terminal records of every shard to every rank (or to rank 0).

    Each rank packs its N_r trajectories into rows of `terminal_dtype` (t, y, aux, n_accepted, n_rejected, status:
    52 bytes for Lorenz-63) with ONE kernel (ni_ensemble_pack_terminal) straight into its slot of a preallocated
    (world, rows, record) buffer; ONE `all_gather_into_tensor` (or `gather` to rank 0) moves them; shard lengths come
    from `shard_indices`, so there is no size exchange, no host synchronisation and no per-field collective.
    Buffers are allocated once and reused by every `run()`."""

    def __init__(self, ens, N_global, interleaved=False, group=None, to_rank0=False):
        import torch
        import torch.distributed as dist
        self.ens, self.N, self.interleaved, self.group, self.to_rank0 = ens, int(N_global), interleaved, group, to_rank0
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dtype = ens.terminal_dtype()
        self.rows = shard_rows(self.N, self.world)
        own = len(shard_indices(self.N, self.rank, self.world, interleaved))
        if own != ens.N:
            raise ValueError(f'rank {self.rank} holds {ens.N} trajectories, the sharding rule gives it {own} of {self.N}')
        dev = torch.device('cuda', ens.device)
        nbytes = self.slot_bytes = _slot_bytes(self.rows, self.dtype)
        # own slot is part of the receive buffer: the pack kernel writes where the collective reads (in place)
        self.recv = torch.zeros((self.world, nbytes), dtype=torch.uint8, device=dev)
        self.send = self.recv[self.rank]

    def run(self):
        """Pack + collective, asynchronous on the ensemble's stream / NCCL's stream; returns the device buffer
        (world, rows * record bytes) -- on every rank, or meaningful on rank 0 only with `to_rank0`."""
        import torch.distributed as dist
        self.ens.pack_terminal(self.send.data_ptr())
        if self.to_rank0:
            dist.gather(self.send, list(self.recv.unbind(0)) if self.rank == 0 else None, dst=0, group=self.group)
        else:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        return self.recv

    def host(self):
        """The gathered records as numpy arrays in GLOBAL trajectory order (one D2H copy)."""
        used = self.rows * self.dtype.itemsize
        recs = np.ascontiguousarray(self.recv.cpu().numpy()[:, :used]).view(self.dtype).reshape(self.world, self.rows)
        return merge_records(recs, self.N, self.world, self.interleaved)


class PeerGather(TerminalGather):
    """The same gather without a library collective: ONE kernel packs the terminal records and stores every word
    straight into this rank's slot of EVERY rank's receive buffer -- peer memory over NVLink / NVSwitch
    (torch.distributed._symmetric_memory maps the buffers of the other processes at rendezvous) -- so the transfer
    overlaps the packing and nothing is staged; a symmetric-memory barrier on either side orders it against the peers'
    reads.  Falls back to the NCCL collective of `TerminalGather` where symmetric memory is not available."""

    def __init__(self, ens, N_global, interleaved=False, group=None):
        import torch
        import torch.distributed as dist
        self.ens, self.N, self.interleaved, self.to_rank0 = ens, int(N_global), interleaved, False
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.dtype = self._record_dtype(ens)
        self.rows = shard_rows(self.N, self.world)
        own = len(shard_indices(self.N, self.rank, self.world, interleaved))
        if own != ens.N:
            raise ValueError(f'rank {self.rank} holds {ens.N} trajectories, the sharding rule gives it {own} of {self.N}')
        nbytes = self.slot_bytes = _slot_bytes(self.rows, self.dtype)
        dev = torch.device('cuda', ens.device)
        self.hdl = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.recv = symm_mem.empty((self.world, nbytes), dtype=torch.uint8, device=dev)
            self.recv.zero_()
            self.hdl = symm_mem.rendezvous(self.recv, self.group)
            # where this rank's rows live in every rank's receive buffer
            self.dst = [int(self.hdl.buffer_ptrs[r]) + self.rank * nbytes for r in range(self.world)]
        except Exception as exc:  # no symmetric memory on this system / backend: the NCCL collective
            self.fallback_reason = f'{type(exc).__name__}: {exc}'
            self.recv = torch.zeros((self.world, nbytes), dtype=torch.uint8, device=dev)
        self.send = self.recv[self.rank]

    def _record_dtype(self, ens):
        return ens.terminal_dtype()

    @property
    def impl(self):
        return 'peer stores over NVLink (symmetric memory), fused with the packing kernel' if self.hdl is not None \
            else 'NCCL all_gather_into_tensor'

    def run(self):
        if self.hdl is None:
            return TerminalGather.run(self)
        self.hdl.barrier()                       # every rank is done reading the previous contents
        self.ens.pack_terminal_peers(self.dst)   # (the ensemble runs on torch's current stream, like the barriers)
        self.hdl.barrier()                       # every rank's stores have landed
        return self.recv


class FusedGather(PeerGather):
    """The gather fused into the INTEGRATION: the run kernel itself stores the terminal record of every trajectory
    that finishes -- 64 bytes for Lorenz-63, four 16-byte vectors -- into this rank's slot of every rank's receive
    buffer (ni_ensemble_set_terminal_sink; peer memory over NVLink / NVSwitch, mapped by symmetric memory).  The
    transfer is spread over the whole run and hides behind the arithmetic; what is left of the collective is the
    barrier after the run.  Usage: `g = FusedGather(ens, N); g.begin(); ens.run_async(); buf = g.finish()`.
    Thread-per-trajectory kernels only.  Without symmetric memory the sink is this rank's own slot and `finish` runs
    the NCCL collective."""

    def __init__(self, ens, N_global, interleaved=False, group=None):
        # CTA-per-trajectory systems have no in-kernel sink (their run kernel sits at the register limit; a 32 KB row per
        # trajectory is also not worth hiding): `finish` packs and stores to the peers after the run, like PeerGather
        self.in_kernel = ens.kind == 0
        PeerGather.__init__(self, ens, N_global, interleaved, group)
        if self.in_kernel:
            ens.set_terminal_sink(self.dst if self.hdl is not None else [self.send.data_ptr()], 0)

    def _record_dtype(self, ens):
        return ens.terminal_sink_dtype() if ens.kind == 0 else ens.terminal_dtype()

    @property
    def impl(self):
        if not self.in_kernel:
            return PeerGather.impl.fget(self) + ' (CTA-per-trajectory kernels: after the run)'
        return 'peer stores over NVLink from the run kernel\'s retire path (symmetric memory): gather fused into the ' \
               'integration' if self.hdl is not None else 'run kernel writes the local slot, NCCL all_gather_into_tensor'

    def begin(self):
        """Before the run: every rank is done reading the previous contents of the receive buffers."""
        if self.hdl is not None:
            self.hdl.barrier()

    def finish(self):
        """After the run (same stream): every rank's records have landed."""
        import torch.distributed as dist
        if not self.in_kernel:
            if self.hdl is not None:
                self.ens.pack_terminal_peers(self.dst)
            else:
                self.ens.pack_terminal(self.send.data_ptr())
        if self.hdl is not None:
            self.hdl.barrier()
        else:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        return self.recv

    def run(self):
        raise RuntimeError('FusedGather has no separate gather step: begin(), run the ensemble, finish()')

    def close(self):
        if self.in_kernel:
            self.ens.set_terminal_sink([])


def merge_records(recs, N, world, interleaved=False):
    """(world, rows) structured terminal records -> dict of (N, ...) arrays in global trajectory order."""
    out = None
    for r in range(world):
        idx = shard_indices(N, r, world, interleaved)
        part = recs[r, :len(idx)]
        if out is None:
            out = {name: np.empty((N,) + part[name].shape[1:], dtype=part[name].dtype) for name in part.dtype.names}
        for name in part.dtype.names:
            out[name][idx] = part[name]
    if 'aux' in out:
        out['auxiliary'] = out.pop('aux')
    return out


def all_gather_records(local, N, group=None):
    """Backend-neutral form of the same collective for host arrays (the gloo tests): `local` is this rank's
    structured record array; returns the (world, rows) array of every rank's records."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = shard_rows(N, world)
    send = np.zeros(rows, dtype=local.dtype)
    send[:len(local)] = local
    s = torch.from_numpy(send.view(np.uint8))
    recv = torch.empty((world, s.numel()), dtype=torch.uint8)
    try:
        dist.all_gather_into_tensor(recv.view(-1), s, group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the flat form: same data through the list form
        dist.all_gather(list(recv.unbind(0)), s, group=group)
    return recv.numpy().view(local.dtype).reshape(world, rows)


def merge_shards(parts, N, world, interleaved=False):
    """Inverse of shard_indices: list of per-rank (N_r, ...) arrays -> one (N, ...) array in global order."""
    first = parts[0]
    out = np.empty((N,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r, p in enumerate(parts):
        out[shard_indices(N, r, world, interleaved)] = p
    return out


def gather(ens, N_global=None, interleaved=False, group=None, peer=False):
    """Final gather of terminal t, y, auxiliary and the step statistics of a sharded ensemble.

    Every rank receives the full arrays in global trajectory order (numpy, host): one pack kernel, one NCCL
    all-gather of equal-sized record slots (or, `peer=True`, one kernel that packs and stores to every GPU's buffer
    over NVLink: PeerGather), one D2H copy."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if N_global is None:  # the only case that needs an exchange: the caller did not say how many trajectories exist
        n = torch.tensor([ens.N], dtype=torch.int64, device=torch.device('cuda', ens.device))
        dist.all_reduce(n, group=group)
        N_global = int(n.item())
    g = PeerGather(ens, N_global, interleaved, group) if peer else TerminalGather(ens, N_global, interleaved, group)
    ens.sync()
    if peer:
        import torch.cuda
        ens.set_stream(torch.cuda.current_stream().cuda_stream)  # the symmetric-memory barriers run on torch's stream
    g.run()
    torch.cuda.synchronize()
    return g.host()
