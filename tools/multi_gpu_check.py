"""Multi-GPU correctness check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank integrates its shard (contiguous blocks for C3, interleaved for the sorted C4 sweep); the NCCL gather, the
peer-store gather and the gather fused into the run kernel must each give every rank the full result in global order,
bit-identical to rank 0 integrating everything alone."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numba_integrators_b200 as ni  # noqa: E402
from numba_integrators_b200 import distributed as nd, workloads as wl  # noqa: E402

rank, world = nd.rank_world()
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for name, w, interleaved in (('c3', wl.c3_lorenz(N=200_003, t_bound=1.0), False),
                             ('c4', wl.c4_vanderpol(N=50_001, tol=1e-8, t_bound=5.0), True),
                             ('c5', wl.c5_heat(N=world * 9 + 3, n=300, t_bound=1.0), False)):
    arrays, idx = nd.shard({'y0': w.y0, 'params': w.params}, rank, world, interleaved)
    f = ni.BuiltinRHS(w.rhs)
    if name == 'c5':  # CTA-per-trajectory layout (the fused form packs after the run: no in-kernel sink there)
        mk = lambda y0, par: ni.RK45(ni.rhs.heat1d(par[:, 0]), w.t0, y0, w.t_bound, rtol=w.rtol, atol=w.atol)
        ens = mk(arrays['y0'], arrays['params'])
    elif w.advanced:
        ens = ni.Advanced(f.P, f.Q, ni.RK45)(f, w.t0, arrays['y0'], arrays['params'], w.t_bound, rtol=w.rtol, atol=w.atol)
    else:
        ens = ni.RK23(f(*arrays['params'].T), w.t0, arrays['y0'], w.t_bound, rtol=w.rtol, atol=w.atol)
    ens.run()
    full = nd.gather(ens, N_global=w.N, interleaved=interleaved)
    full_p = nd.gather(ens, N_global=w.N, interleaved=interleaved, peer=True)  # fused pack + peer stores over NVLink
    same_p = all(np.array_equal(full[k], full_p[k]) for k in full)
    # the gather fused into the integration: a fresh ensemble whose run kernel stores the terminal records to every
    # GPU's receive buffer while it integrates (FusedGather); what follows the run is a barrier
    if name == 'c5':
        ens2 = mk(arrays['y0'], arrays['params'])
    elif w.advanced:
        ens2 = ni.Advanced(f.P, f.Q, ni.RK45)(f, w.t0, arrays['y0'], arrays['params'], w.t_bound, rtol=w.rtol, atol=w.atol)
    else:
        ens2 = ni.RK23(f(*arrays['params'].T), w.t0, arrays['y0'], w.t_bound, rtol=w.rtol, atol=w.atol)
    ens2.sync()
    ens2.set_stream(torch.cuda.current_stream().cuda_stream)
    fg = nd.FusedGather(ens2, w.N, interleaved)
    fg.begin()
    ens2.run_async()
    fg.finish()
    torch.cuda.synchronize()
    full_f = fg.host()
    same_p = same_p and all(np.array_equal(full[k], full_f[k]) for k in full)
    fused_impl = fg.impl
    if rank == 0:
        if name == 'c5':
            ref = mk(w.y0, w.params)
        elif w.advanced:
            ref = ni.Advanced(f.P, f.Q, ni.RK45)(f, w.t0, w.y0, w.params, w.t_bound, rtol=w.rtol, atol=w.atol)
        else:
            ref = ni.RK23(f(*w.params.T), w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol)
        ref.run()
        same = (np.array_equal(full['y'], ref.y) and np.array_equal(full['t'], ref.t)
                and np.array_equal(full['n_accepted'], ref.n_accepted) and np.array_equal(full['n_rejected'], ref.n_rejected)
                and np.array_equal(full['status'], ref.status))
        if w.advanced:
            same = same and np.array_equal(full['auxiliary'], ref.auxiliary)
        same = same and same_p
        print(f'{name}: world={world} N={w.N} shard sizes ~{len(idx)} interleaved={interleaved} '
              f'gather (NCCL, peer-store and fused [{fused_impl}]) == single-GPU run: {same}; accepted {int(full["n_accepted"].sum())}')
        ok = ok and same
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
