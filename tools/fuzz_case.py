"""Replay seeds of tests/test_gpu_fuzz.py with both drivers and list the trajectories that differ from the oracle:

    python tools/fuzz_case.py 67 152
"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numba_integrators_b200 as ni
from oracle import oracle
import test_gpu_fuzz as F
for seed in [int(a) for a in sys.argv[1:]] or [0]:
    cfg = F.make_case(seed)
    print(seed, {k: (v if not isinstance(v, np.ndarray) or v.size < 6 else f'array{v.shape}') for k, v in cfg.items()})
    for driver in ('step', 'run'):
        s = F.build(cfg)
        if driver == 'run':
            s.run()
        else:
            while ni.step(s): pass
        m = oracle.RK45 if cfg['method'] == 'RK45' else oracle.RK23
        N = cfg['N']
        t0 = np.broadcast_to(np.asarray(cfg['t0'], float), (N,)); tb = np.broadcast_to(np.asarray(cfg['t_bound'], float), (N,))
        y, t, acc, rej = s.y, s.t, s.n_accepted, s.n_rejected
        shown = 0; bad = 0
        for i in range(N):
            o = oracle.OracleSolver(cfg['name'], m, t0[i], cfg['y0'][i], tb[i], params=None if cfg['params'] is None else cfg['params'][i], max_step=cfg['max_step'], rtol=cfg['rtol'], atol=cfg['atol'], first_step=cfg['first_step'], advanced=cfg['advanced'])
            while o.step(): pass
            rel = np.max(np.abs(y[i] - o.y)) / np.max(np.abs(o.y))
            if not (t[i] == o.t and acc[i] == o.n_accepted and rej[i] == o.n_rejected and rel <= 1e-11):
                bad += 1
                if shown < 5:
                    shown += 1
                    print('   ', driver, 'i', i, 't', t[i], o.t, 'acc', acc[i], o.n_accepted, 'rej', rej[i], o.n_rejected, 'rel', rel, 'status', s.status[i], o.status, 'mu', cfg['params'][i])
        print('  driver', driver, 'bad', bad, 'of', N)
